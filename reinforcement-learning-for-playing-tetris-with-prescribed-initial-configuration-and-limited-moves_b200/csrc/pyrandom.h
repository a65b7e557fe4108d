// pyrandom.h -- bit-exact restatement of the parts of CPython's `random` module the reference's generators use
// (host code).  Shared by carve_gen.cpp (game/tetris.py:226-352) and forward_gen.cpp (game/tetris_algo_main).
#pragma once
#include <cstdint>

namespace tplgen {

// ---------------------------------------------------------------------------------------------------------------
// CPython's random module: MT19937, seed(int) via init_by_array, randint/shuffle via _randbelow_with_getrandbits
// ---------------------------------------------------------------------------------------------------------------
struct PyRandom {
    uint32_t mt[624]; int idx;
    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void seed(uint64_t a) {                                   // random.seed(int >= 0)
        uint32_t key[2] = {(uint32_t)a, (uint32_t)(a >> 32)};
        const int klen = key[1] ? 2 : 1;
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = 624 > klen ? 624 : klen; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
            if (++j >= klen) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            if (++i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
        idx = 624;
    }
    uint32_t next32() {
        if (idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7FFFFFFFu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9D2C5680u; y ^= (y << 15) & 0xEFC60000u; y ^= y >> 18;
        return y;
    }
    uint32_t randbelow(uint32_t n) {                          // Random._randbelow_with_getrandbits
        int k = 0; for (uint32_t t = n; t; t >>= 1) ++k;       // n.bit_length()
        uint32_t r = next32() >> (32 - k);                     // getrandbits(k), k <= 32
        while (r >= n) r = next32() >> (32 - k);
        return r;
    }
    int randint(int a, int b) { return a + (int)randbelow((uint32_t)(b - a + 1)); }
    template <class T> void shuffle(T *x, int n) {            // random.shuffle
        for (int i = n - 1; i >= 1; --i) { const int j = (int)randbelow((uint32_t)(i + 1)); T t = x[i]; x[i] = x[j]; x[j] = t; }
    }
};


}  // namespace tplgen
