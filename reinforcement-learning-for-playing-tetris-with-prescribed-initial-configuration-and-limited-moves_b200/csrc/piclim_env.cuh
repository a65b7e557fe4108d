// piclim_env.cuh -- per-env bodies of the kernels (one thread = one env), written against an output
// "sink" so the same code is launched by the CUDA kernels (piclim_kernels.cu) and unit-tested on a CPU by
// the host emulation harness under tests/emul/ (test infrastructure only).
#pragma once
#include "piclim_core.cuh"

namespace tpl {

// ---------------------------------------------------------------------------------------------
// boundary conversions: 20 x u16 bitrows (row 0 = top, bit c = column c) <-> 10 bit-columns
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rows_to_cols(const uint16_t *rows, uint32_t (&col)[COLS]) {
#pragma unroll
    for (int c = 0; c < COLS; ++c) col[c] = 0;
    for (int r = 0; r < ROWS; ++r) {
        const uint32_t row = rows[r];
#pragma unroll
        for (int c = 0; c < COLS; ++c) col[c] |= ((row >> c) & 1u) << (ROWS - 1 - r);
    }
}
__device__ __forceinline__ void cols_to_rows(const uint32_t (&col)[COLS], uint16_t *rows) {
    for (int r = 0; r < ROWS; ++r) {
        uint32_t row = 0;
#pragma unroll
        for (int c = 0; c < COLS; ++c) row |= ((col[c] >> (ROWS - 1 - r)) & 1u) << c;
        rows[r] = (uint16_t)row;
    }
}
__device__ __forceinline__ void pack_queue(const uint8_t *pieces, int np, uint32_t (&q)[4]) {
    uint64_t lo = 0, hi = 0;
    for (int p = 0; p < np; ++p) {
        const uint64_t v = pieces[p] & 7u;
        const int bit = 3 * p;
        if (bit < 64) { lo |= v << bit; if (bit > 61) hi |= v >> (64 - bit); }
        else hi |= v << (bit - 64);
    }
    q[0] = (uint32_t)lo; q[1] = (uint32_t)(lo >> 32); q[2] = (uint32_t)hi; q[3] = (uint32_t)(hi >> 32);
}

// ---------------------------------------------------------------------------------------------
// reset: install pool record k with ctor-fresh counters (game/tetris.py:447 and :149-151)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void install_config(Env &e, const uint4 *__restrict__ pool, uint32_t k, uint64_t seed,
                                               uint64_t env, uint32_t episode, int gen_count) {
    const uint4 a = pool[4 * (size_t)k], b = pool[4 * (size_t)k + 1], c = pool[4 * (size_t)k + 2], d = pool[4 * (size_t)k + 3];
    unpack_env(a, b, c, d, e);
    e.lines = 0; e.moves = 0; e.state = S_RUNNING; e.head = 0;
    if (gen_count > 0) { gen_queue(seed, env, episode, gen_count, e.q); e.npieces = (uint32_t)gen_count; }
}

// ---------------------------------------------------------------------------------------------
// step: Tetris.move (game/tetris.py:354-422).  Returns the TPL_FLAG_* bits; k = rows cleared.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t step_env(Env &e, const uint4 *tab, uint32_t rot, uint32_t loc, int L, int M, int &k,
                                             bool &board_changed) {
    k = 0; board_changed = false;
    if (e.head >= e.npieces) return F_NOPIECE;
    const uint32_t piece = queue_piece(e.q, e.head);                        // :356 pop(0)
    e.head += 1;
    const uint4 o = tab[piece * 4 + (rot & 3u)];                            // :359 / :61
    const MoveOut m = place_general(e.col, o, (int)loc);
    k = m.k; board_changed = !m.topout;
    return apply_outcome(e, m, L, M);
}

// ---------------------------------------------------------------------------------------------
// afterstates: slot (r, c) == clone(env).move(r, c); sink.put(slot, word, flags) with
//   word = dlines | holes << 8 | bumpiness << 16 | aggregate_height << 24
// ---------------------------------------------------------------------------------------------
template <class Sink>
__device__ __forceinline__ void afterstates_env(const Env &e, const uint4 *tab, int L, int M, Sink &sink) {
    if (e.head >= e.npieces) {
        for (int s = 0; s < 40; ++s) sink.put(s, 0u, F_NOPIECE);
        return;
    }
    const uint32_t piece = queue_piece(e.q, e.head);

    // per-env precompute, shared by all 40 slots.  Column k lives at index k (col) / k+1 (H); the padding
    // entries are neutral (full columns for the AND, height 0) and every use of them is statically excluded.
    uint32_t col[14]; int H[15];
    H[0] = 0;
    int agg = 0, cells = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) {
        col[k] = e.col[k]; H[k + 1] = col_height(col[k]);
        agg += H[k + 1]; cells += __popc(col[k]);
    }
#pragma unroll
    for (int k = COLS; k < 14; ++k) { col[k] = COL_FULL; H[k + 1] = 0; }
    int D[9]; int bump = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { D[k] = abs(H[k + 1] - H[k + 2]); bump += D[k]; }
    int oldB[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
        int s = 0;
#pragma unroll
        for (int k = c - 1; k <= c + 3; ++k) if (k >= 0 && k <= 8) s += D[k];
        oldB[c] = s;
    }
    uint32_t pre[11], suf[11], A[COLS];
    pre[0] = COL_FULL; suf[10] = COL_FULL;
#pragma unroll
    for (int k = 0; k < COLS; ++k) pre[k + 1] = pre[k] & col[k];
#pragma unroll
    for (int k = COLS - 1; k >= 0; --k) suf[k] = suf[k + 1] & col[k];
#pragma unroll
    for (int c = 0; c < COLS; ++c) A[c] = pre[c] & suf[c + 4 < 10 ? c + 4 : 10];

    const uint32_t U = ((uint32_t)(agg - cells) << 8) | ((uint32_t)bump << 16) | ((uint32_t)agg << 24);
    const uint32_t fl_noclear = ((int)e.moves + 1 >= M) ? F_LOSE : 0u;       // :389-391
    unsigned long long pending = 0ull;

    for (int r = 0; r < 4; ++r) {
        const uint4 o = tab[piece * 4 + r];
        const int w = orient_w(o), h = orient_h(o);
        const int bo0 = o.y & 0xFF, bo1 = (o.y >> 8) & 0xFF, bo2 = (o.y >> 16) & 0xFF, bo3 = o.y >> 24;
        const int to0 = o.z & 0xFF, to1 = (o.z >> 8) & 0xFF, to2 = (o.z >> 16) & 0xFF, to3 = o.z >> 24;
        const uint32_t cb0 = o.x & 15u, cb1 = (o.x >> 4) & 15u, cb2 = (o.x >> 8) & 15u, cb3 = (o.x >> 12) & 15u;
        const uint32_t hm = (1u << h) - 1u;
        const uint32_t afl = orient_alias(o) ? F_ALIAS : 0u;
        uint32_t word = 0, fl = 0; bool pend = false;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            if (c + w <= COLS) {
                // B: hard drop in height form: y = rows below the shape's bottom row
                const int y = max(max(H[c + 1] - bo0, H[c + 2] - bo1), max(H[c + 3] - bo2, H[c + 4] - bo3));
                const bool top = (y + h > ROWS);
                const uint32_t full = A[c] & (col[c] | (cb0 << y)) & (col[c + 1] | (cb1 << y)) &
                                      (col[c + 2] | (cb2 << y)) & (col[c + 3] | (cb3 << y)) & (hm << y);
                // incremental features when no row clears
                const int n0 = y + to0;
                const int n1 = (w > 1) ? y + to1 : H[c + 2];
                const int n2 = (w > 2) ? y + to2 : H[c + 3];
                const int n3 = (w > 3) ? y + to3 : H[c + 4];
                const int a2 = agg - (H[c + 1] + H[c + 2] + H[c + 3] + H[c + 4]) + (n0 + n1 + n2 + n3);
                int nb = 0;
                if (c >= 1) nb += abs(H[c] - n0);
                if (c + 1 <= 9) nb += abs(n0 - n1);
                if (c + 2 <= 9) nb += abs(n1 - n2);
                if (c + 3 <= 9) nb += abs(n2 - n3);
                if (c + 4 <= 9) nb += abs(n3 - H[c + 5]);
                const int b2 = bump - oldB[c] + nb;
                word = ((uint32_t)(a2 - cells - 4) << 8) | ((uint32_t)b2 << 16) | ((uint32_t)a2 << 24);
                fl = fl_noclear; pend = false;
                if (top) { word = U; fl = F_TOPOUT; }
                else if (full) { pend = true; if (!afl) pending |= 1ull << (r * 10 + c); }
            } else {
                fl |= F_ALIAS;
            }
            if (!pend) sink.put(r * 10 + c, word, fl | afl);
        }
    }

    // deferred line-clear slots (rare per slot, so kept out of the unrolled fast path)
    while (pending) {
        const int s = __ffsll((long long)pending) - 1;
        pending &= pending - 1ull;
        const int r = s / 10, c = s - 10 * r;
        const uint4 o = tab[piece * 4 + r];
        uint32_t x[COLS];
#pragma unroll
        for (int k = 0; k < COLS; ++k) x[k] = e.col[k];
        const MoveOut m = place_general(x, o, c);
        const uint32_t f3 = board_features(x, cells + 4 - 10 * m.k);
        const uint32_t word = (uint32_t)m.k | (f3 << 8);
        const uint32_t fl = ((int)e.lines + m.k >= L) ? F_WIN : fl_noclear;   // :415-422
        const int nrot = orient_nrot(o), w = orient_w(o);
        const int cend = (c == COLS - w) ? COLS - 1 : c;
        for (int r2 = r; r2 < 4; r2 += nrot)
            for (int c2 = c; c2 <= cend; ++c2)
                sink.put(r2 * 10 + c2, word, fl | ((r2 != r || c2 != c) ? F_ALIAS : 0u));
    }
}

// ---------------------------------------------------------------------------------------------
// one random-agent move with auto-reset (the body of the fused rollout loop)
//   acc[8] += {episodes, wins, top-outs, move-limit losses, lines, moves placed, steps, resets}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rollout_random_step(Env &e, uint32_t &ep, uint32_t &t, uint32_t (&acc)[8], const uint4 *tab,
                                                    const uint4 *__restrict__ pool, int K, uint64_t seed, uint64_t env,
                                                    int gen_count, int L, int M) {
    if (e.state != S_RUNNING || e.head >= e.npieces) {
        ep += 1; t = 0; acc[7] += 1;
        install_config(e, pool, config_index(seed, env, ep, K), seed, env, ep, gen_count);
    }
    const uint4 rw = rng_words(seed, env, ep, STREAM_ACTION, t);
    const uint32_t rot = rw.x & 3u, loc = __umulhi(rw.y, 10u);
    int k; bool changed;
    const uint32_t fl = step_env(e, tab, rot, loc, L, M, k, changed);
    t += 1;
    acc[6] += 1; acc[4] += (uint32_t)k; acc[5] += changed ? 1u : 0u;
    if (e.state != S_RUNNING) {
        acc[0] += 1;
        if (fl & F_WIN) acc[1] += 1; else if (fl & F_TOPOUT) acc[2] += 1; else acc[3] += 1;
    }
}

// ---------------------------------------------------------------------------------------------
// greedy policy over the afterstate grid: integer linear value, arg-max with lowest-slot tie-break
//   value = w[0]*dlines + w[1]*holes + w[2]*bumpiness + w[3]*agg_height + (win ? w[4] : 0)
//           + (lose or top-out ? w[5] : 0)
// ---------------------------------------------------------------------------------------------
struct GreedySink {
    int w0, w1, w2, w3, w4, w5;
    int best, best_slot;
    __device__ __forceinline__ void put(int slot, uint32_t word, uint32_t fl) {
        int v = w0 * (int)(word & 0xFFu) + w1 * (int)((word >> 8) & 0xFFu) + w2 * (int)((word >> 16) & 0xFFu) +
                w3 * (int)(word >> 24);
        if (fl & F_WIN) v += w4;
        if (fl & (F_LOSE | F_TOPOUT)) v += w5;
        if (v > best || (v == best && slot < best_slot)) { best = v; best_slot = slot; }
    }
};

struct GreedyWeights { int w[6]; };

__device__ __forceinline__ void rollout_greedy_step(Env &e, uint32_t &ep, uint32_t &t, uint32_t (&acc)[8], const uint4 *tab,
                                                    const uint4 *__restrict__ pool, int K, uint64_t seed, uint64_t env,
                                                    int gen_count, int L, int M, const GreedyWeights &gw) {
    if (e.state != S_RUNNING || e.head >= e.npieces) {
        ep += 1; t = 0; acc[7] += 1;
        install_config(e, pool, config_index(seed, env, ep, K), seed, env, ep, gen_count);
    }
    GreedySink sink{gw.w[0], gw.w[1], gw.w[2], gw.w[3], gw.w[4], gw.w[5], (int)0x80000000, 40};
    afterstates_env(e, tab, L, M, sink);
    const int slot = sink.best_slot < 40 ? sink.best_slot : 0;
    const uint32_t rot = (uint32_t)(slot / 10), loc = (uint32_t)(slot - 10 * (slot / 10));
    int k; bool changed;
    const uint32_t fl = step_env(e, tab, rot, loc, L, M, k, changed);
    t += 1;
    acc[6] += 1; acc[4] += (uint32_t)k; acc[5] += changed ? 1u : 0u;
    if (e.state != S_RUNNING) {
        acc[0] += 1;
        if (fl & F_WIN) acc[1] += 1; else if (fl & F_TOPOUT) acc[2] += 1; else acc[3] += 1;
    }
}

}  // namespace tpl
