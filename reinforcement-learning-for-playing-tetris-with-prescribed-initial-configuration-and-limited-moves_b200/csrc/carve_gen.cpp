// carve_gen.cpp -- native prescribed-configuration generator (host code, g++): the supply side of the reset path.
//
// A from-scratch restatement of the reference's carving generator -- Tetris._generate_initial_config / carve /
// calculate_carve (game/tetris.py:226-352), CheckpointManager (:111-137) and RandomPieceGenerator (:64-108) -- on
// 20 x 10-bit bitrows, driven by a bit-exact restatement of CPython's `random` stream (MT19937 + randint/shuffle on
// getrandbits rejection sampling), so that
//     random.seed(s); Tetris(L, M, warm_reset=False, debug=True)          (the reference)
// and carve_generate(s, L, M, ...) produce the same board, the same M+1 pieces and the same recorded solution.
// The reference generates 0.3-19 configs/s/core in Python; this is the same algorithm at
// native speed, threaded over seeds, to fill device-resident config pools.  It is NOT on the GPU rollout path and is
// not a fallback for it: it only produces the (board, pieces) reset points the path consumes (:476-479, :447).
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "pyrandom.h"

namespace {

using tplgen::PyRandom;

// ---------------------------------------------------------------------------------------------------------------
// tetrominoes (game/tetris.py:23-57) as row masks top->bottom, bit j = shape column j
// ---------------------------------------------------------------------------------------------------------------
struct Shape { int h, w; uint8_t m[4]; int prof[4]; };
const uint8_t MASKS[7][4][4] = {
    {{0xF, 0, 0, 0}, {1, 1, 1, 1}, {0}, {0}},
    {{4, 7, 0, 0}, {3, 2, 2, 0}, {7, 1, 0, 0}, {1, 1, 3, 0}},
    {{1, 7, 0, 0}, {2, 2, 3, 0}, {7, 4, 0, 0}, {3, 1, 1, 0}},
    {{2, 7, 0, 0}, {2, 3, 2, 0}, {7, 2, 0, 0}, {1, 3, 1, 0}},
    {{6, 3, 0, 0}, {1, 3, 2, 0}, {0}, {0}},
    {{3, 6, 0, 0}, {2, 3, 1, 0}, {0}, {0}},
    {{3, 3, 0, 0}, {0}, {0}, {0}},
};
const int NROT[7] = {2, 4, 4, 4, 2, 2, 1};

Shape make_shape(int piece, int rot) {
    Shape s{}; const uint8_t *m = MASKS[piece][rot % NROT[piece]];                 // :60-61
    for (int i = 0; i < 4; ++i) {
        s.m[i] = m[i];
        if (m[i]) s.h = i + 1;
        for (int j = 0; j < 4; ++j) if ((m[i] >> j) & 1) { if (j + 1 > s.w) s.w = j + 1; }
    }
    for (int j = 0; j < s.w; ++j) for (int i = 0; i < s.h; ++i) if ((s.m[i] >> j) & 1) s.prof[j] = i;
    return s;
}

struct Board {
    uint16_t r[20];
    int top(int c) const { for (int i = 0; i < 20; ++i) if ((r[i] >> c) & 1) return i; return 20; }   // :429-431
};

// min_j(top[loc+j] - profile[j]) and the first j attaining it (np.argmin), :427-433
void drop_deltas(const Board &b, const Shape &s, int loc, int &dmin, int &jmin) {
    dmin = 1000; jmin = 0;
    for (int j = 0; j < s.w; ++j) {
        const int d = b.top(loc + j) - s.prof[j];
        if (d < dmin) { dmin = d; jmin = j; }
    }
}

// :313-352
bool calculate_carve(Board &b, int drop, int loc, const Shape &s, bool allow_partial) {
    if (drop + s.h > 20) return false;
    if (drop < 0) return false;                                // cannot occur for L <= 16; numpy would wrap
    if (!allow_partial)
        for (int i = 0; i < s.h; ++i)
            if (((b.r[drop + i] >> loc) & s.m[i]) != s.m[i]) return false;            // every shape cell must be filled
    uint16_t saved[4];
    for (int i = 0; i < s.h; ++i) { saved[i] = b.r[drop + i]; b.r[drop + i] &= (uint16_t)~(s.m[i] << loc); }
    int dmin, jmin; drop_deltas(b, s, loc, dmin, jmin);
    if (dmin - 1 != drop) { for (int i = 0; i < s.h; ++i) b.r[drop + i] = saved[i]; return false; }
    return true;
}

// :286-311
bool carve(Board &b, int piece, int rot, int loc, bool allow_partial) {
    const Shape s = make_shape(piece, rot);
    int dmin, jmin; drop_deltas(b, s, loc, dmin, jmin);
    int drop = dmin - 1 + s.prof[jmin] + 1;                    // pushed fully into the stack
    const int tries = allow_partial ? s.h : 1;
    for (int t = 0; t < tries; ++t) {
        if (calculate_carve(b, drop, loc, s, allow_partial)) return true;
        --drop;
    }
    return false;
}

struct Checkpoint { Board board; std::vector<int> pieces; std::vector<std::pair<int, int>> solution; };

// Tetris._generate_initial_config (:226-284).  pieces/solution are kept in play order (index 0 = first piece).
void generate(PyRandom &rng, int L, int M, uint16_t *rows_out, uint8_t *pieces_out, int *npieces_out,
              int8_t *solution_out /* [M][2] or null */, int *nsol_out) {
    Board b; std::memset(b.r, 0, sizeof(b.r));
    for (int i = 20 - L; i < 20; ++i) b.r[i] = 0x3FF;                                  // :228
    std::vector<int> bag;                                                              // RandomPieceGenerator.pieces
    std::vector<int> pieces;                                                           // play order
    std::vector<std::pair<int, int>> solution;
    std::vector<Checkpoint> cps;
    int attempts = 0, cp_uses = 0;                                                     // CheckpointManager (:111-119)
    while (__builtin_popcount(b.r[19]) > 8) {                                          // :234
        bool regenerated = false;
        if (bag.empty()) { bag = {0, 1, 2, 3, 4, 5, 6}; regenerated = true; }          // :71-76
        const int idx = rng.randint(0, (int)bag.size() - 1);                            // :85
        const int piece = bag[idx];
        if (regenerated) cps.push_back(Checkpoint{b, pieces, solution});               // :239-247
        const int rot = rng.randint(0, 3);                                              // :250
        const int w = make_shape(piece, rot).w;
        const int loc = rng.randint(0, 10 - w);                                         // :253
        if ((int)pieces.size() < M && carve(b, piece, rot, loc, pieces.empty())) {      // :257
            pieces.insert(pieces.begin(), piece);                                       // :258
            solution.insert(solution.begin(), {rot, loc});                              // :260
            bag.erase(bag.begin() + idx);                                               // :262
        } else {
            bool load = true;                                                           // 'len >= M or add_attempt()' (:268):
            if ((int)pieces.size() < M) load = ++attempts > 40;                         //  add_attempt only runs when len < M (:121-123)
            if (load) {
                attempts = 0;                                                           // load_checkpoint (:128-137)
                if (cps.size() > 1 && cp_uses > 10) { cps.pop_back(); cp_uses = 0; } else ++cp_uses;
                const Checkpoint &c = cps.back();
                b = c.board; pieces = c.pieces; solution = c.solution;                  // :275-276
                bag = {0, 1, 2, 3, 4, 5, 6};                                            // :278
            }
        }
    }
    // :281-284 pad with get_random_sequence(M - len + 1)  (:95-102)
    const int need = M - (int)pieces.size() + 1;
    int got = 0;
    while (got < need) {
        if (bag.empty()) bag = {0, 1, 2, 3, 4, 5, 6};
        for (int i = (int)bag.size() - 1; i >= 1; --i) {                                // random.shuffle
            const int j = (int)rng.randbelow((uint32_t)(i + 1));
            std::swap(bag[i], bag[j]);
        }
        const int take = need - got < 7 ? need - got : 7;
        for (int i = 0; i < take && i < (int)bag.size(); ++i) { pieces.push_back(bag[i]); ++got; }
        bag.clear();
    }
    for (int i = 0; i < 20; ++i) rows_out[i] = b.r[i];
    *npieces_out = (int)pieces.size();
    for (size_t i = 0; i < pieces.size(); ++i) pieces_out[i] = (uint8_t)pieces[i];
    *nsol_out = (int)solution.size();
    if (solution_out)
        for (size_t i = 0; i < solution.size() && (int)i < M; ++i) {
            solution_out[2 * i] = (int8_t)solution[i].first; solution_out[2 * i + 1] = (int8_t)solution[i].second;
        }
}

}  // namespace

extern "C" {

// Configs for seeds seed0 .. seed0+count-1 (each == random.seed(seed); Tetris(L, M, warm_reset=False, debug=True)).
// rows u16[count][20]; pieces u8[count][pieces_stride] (pieces_stride >= M+1); npieces u8[count];
// solutions i8[count][M][2] or NULL (-1 padded by the caller); nsol u8[count] or NULL.  Returns 0, or -1 on bad arguments.
int carve_generate(uint64_t seed0, int count, int L, int M, uint16_t *rows, uint8_t *pieces, int pieces_stride,
                   uint8_t *npieces, int8_t *solutions, uint8_t *nsol, int nthreads) {
    if (count < 0 || L < 1 || L > 16 || M < 1 || pieces_stride < M + 1 || !rows || !pieces || !npieces) return -1;
    if (nthreads < 1) nthreads = 1;
    auto work = [&](int t) {
        for (int k = t; k < count; k += nthreads) {
            int np = 0, ns = 0;
            PyRandom rng; rng.seed(seed0 + (uint64_t)k);
            generate(rng, L, M, rows + (size_t)k * 20, pieces + (size_t)k * pieces_stride, &np,
                     solutions ? solutions + (size_t)k * M * 2 : nullptr, &ns);
            npieces[k] = (uint8_t)np;
            if (nsol) nsol[k] = (uint8_t)ns;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    return 0;
}

// One config drawn from a caller-supplied MT19937 state (CPython's random.getstate()[1]: 624 words + index) which is
// advanced in place: lets the Python facade consume the GLOBAL `random` stream exactly like the reference's
// Tetris(L, M, warm_reset=False) does, so code that seeds `random` gets the reference's configs and stream position.
int carve_generate_from_state(uint32_t *mt625, int L, int M, uint16_t *rows, uint8_t *pieces, int pieces_stride, uint8_t *npieces,
                              int8_t *solution, uint8_t *nsol) {
    if (!mt625 || L < 1 || L > 16 || M < 1 || pieces_stride < M + 1 || !rows || !pieces || !npieces) return -1;
    PyRandom rng;
    std::memcpy(rng.mt, mt625, sizeof(rng.mt));
    rng.idx = (int)mt625[624];
    int np = 0, ns = 0;
    generate(rng, L, M, rows, pieces, &np, solution, &ns);
    *npieces = (uint8_t)np;
    if (nsol) *nsol = (uint8_t)ns;
    std::memcpy(mt625, rng.mt, sizeof(rng.mt));
    mt625[624] = (uint32_t)rng.idx;
    return 0;
}

// Tetris.carve (:286-311) on a bitrow board, in place.  Returns 1 if the piece was carved, 0 if not, -1 on bad arguments.
int carve_apply(uint16_t *rows, int piece, int rot, int loc, int allow_partial) {
    if (!rows || piece < 0 || piece > 6 || loc < 0) return -1;
    const Shape s = make_shape(piece, ((rot % 4) + 4) % 4);
    if (loc + s.w > 10) return -1;
    Board b; std::memcpy(b.r, rows, sizeof(b.r));
    const bool ok = carve(b, piece, ((rot % 4) + 4) % 4, loc, allow_partial != 0);
    std::memcpy(rows, b.r, sizeof(b.r));
    return ok ? 1 : 0;
}

// the first n outputs of random.seed(seed); [random.randint(0, hi) ...]  (lets the tests pin the RNG restatement)
void carve_pyrandom_randints(uint64_t seed, int hi, int n, int32_t *out) {
    PyRandom r; r.seed(seed);
    for (int i = 0; i < n; ++i) out[i] = r.randint(0, hi);
}

}  // extern "C"
