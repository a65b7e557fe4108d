// host_shim.h -- lets g++ compile the device headers (piclim_core.cuh / piclim_env.cuh) so the per-env
// device logic can be unit-tested on a CPU-only box.  TEST INFRASTRUCTURE ONLY: nothing in the product
// package includes or links this; the shipped library is built by nvcc for sm_100a and has no CPU path.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static const

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31u));
}
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t __vsadu4(uint32_t a, uint32_t b) {
    uint32_t s = 0;
    for (int i = 0; i < 4; ++i) { int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF; s += (uint32_t)(x > y ? x - y : y - x); }
    return s;
}
static inline uint32_t __sad(int a, int b, uint32_t c) { return c + (uint32_t)(a > b ? a - b : b - a); }
static inline uint32_t __viaddmax_s16x2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r = 0;
    for (int i = 0; i < 2; ++i) {
        const int16_t x = (int16_t)(a >> (16 * i)), y = (int16_t)(b >> (16 * i)), z = (int16_t)(c >> (16 * i));
        const int16_t sum = (int16_t)(x + y);
        r |= (uint32_t)(uint16_t)(sum > z ? sum : z) << (16 * i);
    }
    return r;
}
static inline int __viaddmax_s32(int a, int b, int c) { return a + b > c ? a + b : c; }
static inline uint32_t __reduce_max_sync(uint32_t, uint32_t v) { return v; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
