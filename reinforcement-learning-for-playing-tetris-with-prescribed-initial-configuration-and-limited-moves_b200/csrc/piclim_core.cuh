// piclim_core.cuh -- device-side core of the B200 Tetris-piclim hot path (sm_100a).
//
// Everything here restates, on a bit-column board, the semantics of the reference's
// game/tetris.py (citations are file:line in the reference tree):
//   tetromino table + get_tetromino   :23-57, :60-61
//   calculate_drop(_deltas)           :424-433
//   Tetris.move                       :354-422
// Board representation: 10 x uint32 "bit-columns"; bit b of col[c] is the cell in row (19 - b) of
// column c, so bit 0 is the floor row and the column height is simply 32 - clz(col[c]).  With the
// board stored column-wise, every quantity the move needs is one or two integer instructions per
// column: hard-drop = max over <= 4 columns of (height - piece bottom offset), placement = OR of a
// shifted 4-bit piece column, full rows = AND over the ten columns, line clear = deleting <= 4 bit
// positions from each column.  The boundary format (20 x uint16 bitrows, row 0 = top) is converted
// in tpl_pack / tpl_unpack only.
#pragma once
#include <stdint.h>
#ifdef TPL_HOST_EMUL
// tests/emul/ compiles this header with g++ to unit-test the device logic on a CPU-only box
// (test infrastructure; the product library is only ever built by nvcc for sm_100a).
#include "../../tests/emul/host_shim.h"
#else
#include <cuda_runtime.h>
#endif

namespace tpl {

constexpr int ROWS = 20;
constexpr int COLS = 10;
constexpr uint32_t COL_FULL = 0xFFFFFu;

enum : uint32_t { F_TOPOUT = 1, F_WIN = 2, F_LOSE = 4, F_ALIAS = 8, F_NOPIECE = 16 };
enum : uint32_t { S_RUNNING = 0, S_WON = 1, S_LOST = 2 };

// ---------------------------------------------------------------------------------------------
// A. orientation table.  Two uint4 per (piece, rot 0..3); rot >= n_rot repeats rot % n_rot (:61).
//   a.x: bits 0-15 cb  : 4-bit column images, nibble j = shape column j, bit i = cell i rows above
//                        the shape's bottom row
//        bits 16-18 w, bits 20-22 h, bits 24-25 n_rot-1, bit 28 = rot is an alias (rot >= n_rot)
//   a.y: bo bytes: byte j = rows between the shape's bottom row and the lowest cell of column j
//        (= h-1-profile[j] with profile the tuple at :25-55); 64 for j >= w so it never wins the max
//   a.z: to bytes: byte j = to_j = (height of the highest cell of column j above the bottom row) + 1, 0 for j >= w;
//        a covered column's new height is y + to_j (the piece always ends up on top of it), for all four columns at
//        once y * 0x01010101 + a.z
//   a.w: cover mask: byte j = 0xFF for j < w (selects the new height), 0 for the columns the shape does not cover
//   b.x: bo nibbles (4 bits per column, no sentinel) -- used by the column-aligned general path
//   b.y: (1 << h) - 1      b.z: 20 - h (top-out iff y > 20 - h)
//   b.w: the "distinct placements" (alias-free) numbering of this piece: bits 0-7 = index of placement (this rotation,
//        column 0) in the piece's run of distinct placements (rotations r < n_rot, columns c <= 10 - w, rotation-major),
//        0xFF when the rotation is an alias; bits 8-15 = length of the run (9 / 17 / 34)
// ---------------------------------------------------------------------------------------------
struct OrientEntry { uint32_t ax, ay, az, aw, bx, by, bz, bw; };

constexpr OrientEntry make_orient(int m0, int m1, int m2, int m3, int nrot, bool alias, int rot_base = 0xFF, int run_len = 0) {
    int m[4] = {m0, m1, m2, m3};
    int h = 0, w = 0;
    for (int i = 0; i < 4; ++i) {
        if (m[i]) h = i + 1;
        for (int j = 0; j < 4; ++j) if ((m[i] >> j) & 1) { if (j + 1 > w) w = j + 1; }
    }
    uint32_t cb = 0, bo = 0, bon = 0;
    uint32_t to4 = 0, cover = 0;
    for (int j = 0; j < 4; ++j) {
        if (j >= w) { bo |= 64u << (8 * j); continue; }
        int lowest = -1, highest = -1;
        for (int i = 0; i < h; ++i) if ((m[i] >> j) & 1) {
            lowest = i; if (highest < 0) highest = i;
            cb |= 1u << (4 * j + (h - 1 - i));
        }
        bo |= (uint32_t)(h - 1 - lowest) << (8 * j);
        bon |= (uint32_t)(h - 1 - lowest) << (4 * j);
        to4 |= (uint32_t)(h - highest) << (8 * j);
        cover |= 0xFFu << (8 * j);
    }
    return OrientEntry{cb | ((uint32_t)w << 16) | ((uint32_t)h << 20) | ((uint32_t)(nrot - 1) << 24) | (alias ? 1u << 28 : 0u),
                       bo, to4, cover,
                       bon, (1u << h) - 1u, (uint32_t)(20 - h), (uint32_t)(alias ? 0xFF : rot_base) | ((uint32_t)run_len << 8)};
}

// The afterstate enumeration reads the same facts once per rotation, already split into registers (no shift/mask work on
// the ALU pipe that bounds it): three uint4 per (piece, rot)
//   n: nb_j = -bo_j as int (-64 for j >= w)          c: cb_j, the 4-bit column images
//   m: to bytes | cover mask | (1 << h) - 1 | 20 - h
// w, the alias bit and n_rot are taken from the compact entry (a.x) where they are needed.
struct OrientWide { int32_t nb[4]; uint32_t cb[4]; uint32_t to4, cover, hm, thr; };

constexpr OrientWide make_wide(const OrientEntry &e) {
    return OrientWide{{-(int32_t)(e.ay & 0xFF), -(int32_t)((e.ay >> 8) & 0xFF), -(int32_t)((e.ay >> 16) & 0xFF), -(int32_t)(e.ay >> 24)},
                      {e.ax & 15u, (e.ax >> 4) & 15u, (e.ax >> 8) & 15u, (e.ax >> 12) & 15u},
                      e.az, e.aw, e.by, e.bz};
}

struct OrientTable { OrientEntry e[28]; OrientWide w[28]; };

// row masks top->bottom, bit j = shape column j, exactly the arrays at game/tetris.py:25-55
#define TPL_O(a, b, c, d, n, al, rb, len) make_orient(a, b, c, d, n, al, rb, len)
constexpr OrientTable make_table() {
    OrientTable t{{
    // I: 7 + 10 distinct placements
    TPL_O(0xF, 0, 0, 0, 2, false, 0, 17), TPL_O(1, 1, 1, 1, 2, false, 7, 17), TPL_O(0xF, 0, 0, 0, 2, true, 0xFF, 17), TPL_O(1, 1, 1, 1, 2, true, 0xFF, 17),
    // L: 8 + 9 + 8 + 9
    TPL_O(4, 7, 0, 0, 4, false, 0, 34), TPL_O(3, 2, 2, 0, 4, false, 8, 34), TPL_O(7, 1, 0, 0, 4, false, 17, 34), TPL_O(1, 1, 3, 0, 4, false, 25, 34),
    // J
    TPL_O(1, 7, 0, 0, 4, false, 0, 34), TPL_O(2, 2, 3, 0, 4, false, 8, 34), TPL_O(7, 4, 0, 0, 4, false, 17, 34), TPL_O(3, 1, 1, 0, 4, false, 25, 34),
    // T
    TPL_O(2, 7, 0, 0, 4, false, 0, 34), TPL_O(2, 3, 2, 0, 4, false, 8, 34), TPL_O(7, 2, 0, 0, 4, false, 17, 34), TPL_O(1, 3, 1, 0, 4, false, 25, 34),
    // S: 8 + 9
    TPL_O(6, 3, 0, 0, 2, false, 0, 17), TPL_O(1, 3, 2, 0, 2, false, 8, 17), TPL_O(6, 3, 0, 0, 2, true, 0xFF, 17), TPL_O(1, 3, 2, 0, 2, true, 0xFF, 17),
    // Z
    TPL_O(3, 6, 0, 0, 2, false, 0, 17), TPL_O(2, 3, 1, 0, 2, false, 8, 17), TPL_O(3, 6, 0, 0, 2, true, 0xFF, 17), TPL_O(2, 3, 1, 0, 2, true, 0xFF, 17),
    // O: 9
    TPL_O(3, 3, 0, 0, 1, false, 0, 9), TPL_O(3, 3, 0, 0, 1, true, 0xFF, 9), TPL_O(3, 3, 0, 0, 1, true, 0xFF, 9), TPL_O(3, 3, 0, 0, 1, true, 0xFF, 9),
    }, {}};
    for (int i = 0; i < 28; ++i) t.w[i] = make_wide(t.e[i]);
    return t;
}
#undef TPL_O
#ifdef TPL_HOST_EMUL
static const OrientTable c_orient = make_table();
#else
// In global memory, not __constant__: its only reader is the per-CTA copy into shared memory, one 16-byte chunk per thread, i.e.
// 32 different addresses per warp, which the constant cache serves one at a time (ncu: 1.9 % of the fused step's warp time sat on
// that copy); from global memory it is three coalesced 512-byte requests per CTA out of L2.
__device__ const OrientTable c_orient = make_table();
#endif

constexpr int TAB_COMPACT4 = 56;     // 28 entries x 2 uint4
constexpr int TAB_WORDS4 = 56 + 84;  // + 28 wide entries x 3 uint4

// `o` below is the first uint4 (a) of an entry, `ob` the second (b)
__device__ __forceinline__ int orient_w(const uint4 &o) { return (o.x >> 16) & 7; }
__device__ __forceinline__ int orient_h(const uint4 &o) { return (o.x >> 20) & 7; }
__device__ __forceinline__ int orient_nrot(const uint4 &o) { return ((o.x >> 24) & 3) + 1; }
__device__ __forceinline__ bool orient_alias(const uint4 &o) { return (o.x >> 28) & 1; }
// `ob.w`: see b.w above
__device__ __forceinline__ uint32_t orient_rot_base(const uint4 &ob) { return ob.w & 0xFFu; }
__device__ __forceinline__ uint32_t orient_run_len(const uint4 &ob) { return (ob.w >> 8) & 0xFFu; }
constexpr int DISTINCT_MAX = 34;     // most distinct placements any piece has (L, J, T)

// ---------------------------------------------------------------------------------------------
// env record in registers
// ---------------------------------------------------------------------------------------------
struct Env {
    uint32_t col[COLS];
    uint32_t q[4];          // piece queue, 3 bits per piece
    uint32_t lines, moves, state, head, npieces;
    uint32_t qblock;        // generated queues longer than 42 pieces: which 42-piece block of the episode's sequence q holds
};

__device__ __forceinline__ uint4 ldg_plain(const uint4 *p) { return *p; }

__device__ __forceinline__ void unpack_env(const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d, Env &e) {
    e.col[0] = a.x; e.col[1] = a.y; e.col[2] = a.z; e.col[3] = a.w;
    e.col[4] = b.x; e.col[5] = b.y; e.col[6] = b.z; e.col[7] = b.w;
    e.col[8] = c.x; e.col[9] = c.y; e.q[0] = c.z; e.q[1] = c.w;
    e.q[2] = d.x; e.q[3] = d.y;
    e.lines = d.z & 0xFFFFu; e.moves = d.z >> 16;
    e.state = d.w & 0xFFu; e.head = (d.w >> 8) & 0xFFu; e.npieces = (d.w >> 16) & 0xFFu; e.qblock = d.w >> 24;
}
__device__ __forceinline__ uint4 pack_meta(const Env &e) {
    return make_uint4(e.q[2], e.q[3], (e.lines & 0xFFFFu) | (e.moves << 16),
                      (e.state & 0xFFu) | ((e.head & 0xFFu) << 8) | ((e.npieces & 0xFFu) << 16) | (e.qblock << 24));
}

// state planes: chunk j of env i at st[j * stride + i]
__device__ __forceinline__ void load_env(const uint4 *st, int64_t stride, int64_t i, Env &e) {
    uint4 a = ldg_plain(st + i), b = ldg_plain(st + stride + i), c = ldg_plain(st + 2 * stride + i),
          d = ldg_plain(st + 3 * stride + i);
    unpack_env(a, b, c, d, e);
}
__device__ __forceinline__ void store_env(uint4 *st, int64_t stride, int64_t i, const Env &e) {
    st[i] = make_uint4(e.col[0], e.col[1], e.col[2], e.col[3]);
    st[stride + i] = make_uint4(e.col[4], e.col[5], e.col[6], e.col[7]);
    st[2 * stride + i] = make_uint4(e.col[8], e.col[9], e.q[0], e.q[1]);
    st[3 * stride + i] = pack_meta(e);
}

// piece `idx` of the 128-bit queue
__device__ __forceinline__ uint32_t queue_piece(const uint32_t (&q)[4], uint32_t idx) {
    uint32_t bit = 3u * idx, wd = bit >> 5, off = bit & 31u;
    uint32_t lo = wd == 0 ? q[0] : wd == 1 ? q[1] : wd == 2 ? q[2] : q[3];
    uint32_t hi = wd == 0 ? q[1] : wd == 1 ? q[2] : wd == 2 ? q[3] : 0u;
    return __funnelshift_r(lo, hi, off) & 7u;
}

// height of a bit-column = index of its highest set bit + 1 (bfind returns -1 for 0): one FLO + one add
__device__ __forceinline__ int col_height(uint32_t c) {
#ifdef TPL_HOST_EMUL
    return 32 - __clz(c);
#else
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(c));
    return r + 1;
#endif
}

// a * b + c pinned to the FMA pipe (IMAD), e.g. to pack bytes with `hi * 256 + lo` instead of an ALU-pipe PRMT/LEA
__device__ __forceinline__ uint32_t mad_fma_pipe(uint32_t a, uint32_t b, uint32_t c) {
#ifdef TPL_HOST_EMUL
    return a * b + c;
#else
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#endif
}

// two-way dot products of s16 pairs with u8 pairs (IDP.2A): a.lo * b.byte0 + a.hi * b.byte1 + c, resp. bytes 2 and 3
__device__ __forceinline__ int dp2a_lo(uint32_t a, uint32_t b, int c) {
#ifdef TPL_HOST_EMUL
    return (int)(int16_t)(a & 0xFFFFu) * (int)(b & 0xFFu) + (int)(int16_t)(a >> 16) * (int)((b >> 8) & 0xFFu) + c;
#else
    int r;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));       // signed halves of a, unsigned bytes of b
    return r;
#endif
}
__device__ __forceinline__ int dp2a_hi(uint32_t a, uint32_t b, int c) {
#ifdef TPL_HOST_EMUL
    return (int)(int16_t)(a & 0xFFFFu) * (int)((b >> 16) & 0xFFu) + (int)(int16_t)(a >> 16) * (int)(b >> 24) + c;
#else
    int r;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#endif
}

// sum over the four bytes of |a - b|, plus c: ONE instruction (VABSDIFF4.U8.ACC with a live accumulator).  Written as
// `__vsadu4(a, b) + c` the accumulator operand stays zero and ptxas appends a separate add.
__device__ __forceinline__ uint32_t vsad4_acc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef TPL_HOST_EMUL
    return __vsadu4(a, b) + c;
#else
    uint32_t r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#endif
}

// a * b on the FMA pipe (IMAD): used for `bits << y` as bits * (1 << y) so the variable shifts do not all
// land on the ALU pipe, which is the pipe that bounds the afterstate kernel
__device__ __forceinline__ uint32_t mul_fma_pipe(uint32_t a, uint32_t b) {
#ifdef TPL_HOST_EMUL
    return a * b;
#else
    uint32_t r;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#endif
}

// ---------------------------------------------------------------------------------------------
// C. the general move on bit-columns (any loc, static column indexing only).
//    Returns rows cleared (0..4); topout set when drop row < 0 (board untouched, :372-374).
// ---------------------------------------------------------------------------------------------
struct MoveOut { int k; bool topout; };

// Per-thread scratch for the one place where a column index is data-dependent (the move's location):
// SCR_ROWS words per thread at scr[k * ss]; with ss = blockDim.x the words of a warp sit in 32 different
// banks, so the dynamic accesses are conflict-free.  Rows 10..12 are padding for loc + j > 9.
constexpr int SCR_ROWS = 13;

// v >> s with s >= 32 giving 0 (PTX shift semantics; `>>` in C++ is undefined there)
__device__ __forceinline__ uint32_t shr_clamp(uint32_t v, uint32_t s) {
#ifdef TPL_HOST_EMUL
    return s >= 32 ? 0u : v >> s;
#else
    return __funnelshift_rc(v, 0u, s);
#endif
}

// The move on bit-columns for a run-time location.  Hard drop (:424-433) in one bit-scan:
//   max_j(H[loc+j] - bo_j) = bfind( OR_j (col[loc+j] >> bo_j) ) + 1
// (bo_j = 64 for the columns the shape does not cover, which shifts them out entirely).
__device__ __forceinline__ MoveOut place_general(uint32_t (&x)[COLS], uint32_t *scr, int ss, const uint4 &o, const uint4 &ob,
                                                 int loc_raw) {
    const int w = orient_w(o);
    const int loc = min(loc_raw, COLS - w);                                  // :364
#pragma unroll
    for (int k = 0; k < COLS; ++k) scr[k * ss] = x[k];
    uint32_t *win = scr + loc * ss;
    const uint32_t v0 = win[0], v1 = win[ss], v2 = win[2 * ss], v3 = win[3 * ss];
    const uint32_t t = shr_clamp(v0, o.y & 0xFFu) | shr_clamp(v1, (o.y >> 8) & 0xFFu) | shr_clamp(v2, (o.y >> 16) & 0xFFu) |
                       shr_clamp(v3, o.y >> 24);
    const int y = col_height(t);
    MoveOut r; r.k = 0; r.topout = (y > (int)ob.z);                           // drop = 20 - h - y < 0  (:372-374)
    if (r.topout) return r;
    const uint32_t pw = 1u << y;
    win[0] = v0 | ((o.x & 15u) * pw);                                         // :377-378
    win[ss] = v1 | (((o.x >> 4) & 15u) * pw);
    win[2 * ss] = v2 | (((o.x >> 8) & 15u) * pw);
    win[3 * ss] = v3 | (((o.x >> 12) & 15u) * pw);
    uint32_t full = ob.y * pw;                                                // only the piece's rows (:382-383)
#pragma unroll
    for (int k = 0; k < COLS; ++k) { x[k] = scr[k * ss]; full &= x[k]; }
    r.k = __popc(full);
    while (full) {                                                            // :397-407, highest row first
        const int q = 31 - __clz(full);
        const uint32_t low = (1u << q) - 1u;
#pragma unroll
        for (int k = 0; k < COLS; ++k) x[k] = (x[k] & low) | ((x[k] >> 1) & ~low);
        full &= ~(1u << q);
    }
    return r;
}

// win / lose bookkeeping shared by every caller (:379, :389-391, :409-422)
__device__ __forceinline__ uint32_t apply_outcome(Env &e, const MoveOut &m, int L, int M) {
    uint32_t fl = 0;
    if (m.topout) { e.state = S_LOST; return F_TOPOUT; }
    e.moves += 1;
    if (m.k == 0) {
        if ((int)e.moves >= M) { e.state = S_LOST; fl = F_LOSE; }
        return fl;
    }
    e.lines += (uint32_t)m.k;
    if ((int)e.lines >= L) { e.state = S_WON; fl = F_WIN; }
    else if ((int)e.moves >= M) { e.state = S_LOST; fl = F_LOSE; }
    return fl;
}

// features of a bit-column board: holes | bumpiness<<8 | agg<<16
__device__ __forceinline__ uint32_t board_features(const uint32_t (&x)[COLS], int cells) {
    int hprev = col_height(x[0]);
    int agg = hprev, bump = 0;
#pragma unroll
    for (int k = 1; k < COLS; ++k) {
        const int hk = col_height(x[k]);
        agg += hk; bump += abs(hk - hprev); hprev = hk;
    }
    return (uint32_t)(agg - cells) | ((uint32_t)bump << 8) | ((uint32_t)agg << 16);
}

__device__ __forceinline__ int board_cells(const uint32_t (&x)[COLS]) {
    int c = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) c += __popc(x[k]);
    return c;
}

// ---------------------------------------------------------------------------------------------
// G. Philox4x32-10 (Salmon et al. SC'11) and the 7-bag built on it
// ---------------------------------------------------------------------------------------------
enum : uint32_t { STREAM_PIECES = 0, STREAM_ACTION = 1, STREAM_CONFIG = 2 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ uint4 rng_words(uint64_t seed, uint64_t env, uint32_t episode, uint32_t stream, uint32_t index) {
    return philox4x32_10(make_uint4((uint32_t)env, (uint32_t)(env >> 32), episode, (stream << 28) | (index & 0x0FFFFFFFu)),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// one uniformly chosen permutation of 0..6 as seven 3-bit fields (field t = t-th piece of the bag)
__device__ __forceinline__ uint32_t bag_from_word(uint32_t u) {
    uint32_t k = __umulhi(u, 5040u);
    uint32_t perm = 0u;
#pragma unroll
    for (int t = 0; t < 7; ++t) perm |= (uint32_t)t << (3 * t);
#pragma unroll
    for (int i = 6; i >= 1; --i) {
        const uint32_t j = k % (uint32_t)(i + 1);
        k /= (uint32_t)(i + 1);
        const uint32_t x = ((perm >> (3 * i)) ^ (perm >> (3 * j))) & 7u;
        perm ^= (x << (3 * i)) | (x << (3 * j));
    }
    return perm;
}

// pieces [42 * block, 42 * block + count) (count <= 42) of episode `episode` of env `env`, packed 3 bits each.  The episode's
// sequence is a concatenation of 7-bags; bag g comes from word (g & 3) of Philox call (g >> 2) of the piece stream, so the six
// bags of a block are six consecutive words of two calls, starting at word 0 or 2 (block 0: calls 0 and 1, words 0..5).
__device__ __forceinline__ void gen_queue(uint64_t seed, uint64_t env, uint32_t episode, int count, uint32_t (&q)[4], uint32_t block = 0) {
    uint64_t lo = 0, hi = 0;
    const uint32_t g0 = 6u * block, c0 = g0 >> 2;
    const bool odd = (g0 & 2u) != 0u;
    uint4 w0 = rng_words(seed, env, episode, STREAM_PIECES, c0);
    uint4 w1 = rng_words(seed, env, episode, STREAM_PIECES, c0 + 1u);
    const uint32_t ws[6] = {odd ? w0.z : w0.x, odd ? w0.w : w0.y, odd ? w1.x : w0.z, odd ? w1.y : w0.w, odd ? w1.z : w1.x, odd ? w1.w : w1.y};
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const uint64_t perm = bag_from_word(ws[b]);
        const int off = 21 * b;
        if (off < 64) { lo |= perm << off; if (off + 21 > 64) hi |= perm >> (64 - off); }
        else hi |= perm << (off - 64);
    }
    const int bits = 3 * count;
    if (bits < 64) { lo &= (1ull << bits) - 1ull; hi = 0; }
    else if (bits < 128) hi &= (1ull << (bits - 64)) - 1ull;
    q[0] = (uint32_t)lo; q[1] = (uint32_t)(lo >> 32); q[2] = (uint32_t)hi; q[3] = (uint32_t)(hi >> 32);
}

// Out-of-line copy for the kernels' hot loops: a queue is generated only when an env is reset in generate mode or runs dry
// mid-episode, and two inlined Philox calls per call site would only bloat the loop bodies the instruction cache has to hold.
#ifdef TPL_HOST_EMUL
#define TPL_NOINLINE
#else
#define TPL_NOINLINE __noinline__
#endif
static __device__ TPL_NOINLINE uint4 gen_queue_cold(uint64_t seed, uint64_t env, uint32_t episode, int count, uint32_t block) {
    uint32_t q[4];
    gen_queue(seed, env, episode, count, q, block);
    return make_uint4(q[0], q[1], q[2], q[3]);
}

constexpr int QUEUE_PIECES = 42;                 // pieces the 128-bit queue holds
constexpr int GEN_MAX = QUEUE_PIECES * 256;      // longest generated sequence (the block number is one byte of the record)

// Generated sequences longer than the queue (gen_total > 42, i.e. M > 41): when the queue runs dry mid-episode, the next
// 42-piece block of the same counter-based sequence is generated in place (RandomPieceGenerator.get_random_sequence yields
// sequences of ANY length, game/tetris.py:95-102).  True if the env was running on an empty queue and got new pieces.
__device__ __forceinline__ bool refill_queue(Env &e, uint64_t seed, uint64_t env, uint32_t episode, int gen_total) {
    if (gen_total <= QUEUE_PIECES || e.state != S_RUNNING || e.head < e.npieces) return false;
    const int done = (int)(e.qblock + 1u) * QUEUE_PIECES;
    if (done >= gen_total) return false;
    e.qblock += 1u;
    const int count = min(QUEUE_PIECES, gen_total - done);
    const uint4 q = gen_queue_cold(seed, env, episode, count, e.qblock);
    e.q[0] = q.x; e.q[1] = q.y; e.q[2] = q.z; e.q[3] = q.w;
    e.head = 0; e.npieces = (uint32_t)count;
    return true;
}

__device__ __forceinline__ uint32_t config_index(uint64_t seed, uint64_t env, uint32_t episode, int K) {
    return __umulhi(rng_words(seed, env, episode, STREAM_CONFIG, 0).x, (uint32_t)K);
}

}  // namespace tpl
