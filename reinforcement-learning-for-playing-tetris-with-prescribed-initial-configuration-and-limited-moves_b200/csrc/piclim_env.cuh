// piclim_env.cuh -- per-env bodies of the kernels (one thread = one env), written against an output
// "sink" so the same code is launched by the CUDA kernels (piclim_kernels.cu) and unit-tested on a CPU by
// the host emulation harness under tests/emul/ (test infrastructure only).
#pragma once
#include "piclim_core.cuh"

namespace tpl {

// Sinks that only want the VALUE of a placement (the greedy rollout's arg-max) declare `static constexpr bool VALUE_ONLY = true`
// and take the fast path's words through begin_env / put_value (see GreedySinkT): no flags word, no alias copies.
template <class S, class = void> struct SinkValueOnly { static constexpr bool value = false; };
template <class S> struct SinkValueOnly<S, decltype((void)S::VALUE_ONLY)> { static constexpr bool value = S::VALUE_ONLY; };

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
        const uint64_t v = pieces[p] > 6 ? 6u : pieces[p];       // ids are 0..6 (host entry points reject others); never index past the table
        const int bit = 3 * p;
        if (bit < 64) { lo |= v << bit; if (bit > 61) hi |= v >> (64 - bit); }
        else hi |= v << (bit - 64);
    }
    q[0] = (uint32_t)lo; q[1] = (uint32_t)(lo >> 32); q[2] = (uint32_t)hi; q[3] = (uint32_t)(hi >> 32);
}

// ---------------------------------------------------------------------------------------------
// reset: install pool record k with ctor-fresh counters (game/tetris.py:447 and :149-151)
// ---------------------------------------------------------------------------------------------
// (install_record: the four chunks of the pool record come from the caller -- the fused step has them copied into shared
// memory a whole tile ahead, see step_observe_kernel)
__device__ __forceinline__ void install_record(Env &e, const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d, uint64_t seed,
                                               uint64_t env, uint32_t episode, int gen_count);
__device__ __forceinline__ void install_config(Env &e, const uint4 *__restrict__ pool, uint32_t k, uint64_t seed,
                                               uint64_t env, uint32_t episode, int gen_count) {
    const uint4 a = pool[4 * (size_t)k], b = pool[4 * (size_t)k + 1], c = pool[4 * (size_t)k + 2], d = pool[4 * (size_t)k + 3];
    install_record(e, a, b, c, d, seed, env, episode, gen_count);
}
__device__ __forceinline__ void install_record(Env &e, const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d, uint64_t seed,
                                               uint64_t env, uint32_t episode, int gen_count) {
    unpack_env(a, b, c, d, e);
    e.lines = 0; e.moves = 0; e.state = S_RUNNING; e.head = 0; e.qblock = 0;
    if (gen_count > 0) {                      // (more than 42: the first block now, the rest by refill_queue as the episode goes on)
        const int first = gen_count < QUEUE_PIECES ? gen_count : QUEUE_PIECES;
        const uint4 q = gen_queue_cold(seed, env, episode, first, 0u);
        e.q[0] = q.x; e.q[1] = q.y; e.q[2] = q.z; e.q[3] = q.w; e.npieces = (uint32_t)first;
    }
}

// ---------------------------------------------------------------------------------------------
// step: Tetris.move (game/tetris.py:354-422).  Returns the TPL_FLAG_* bits; k = rows cleared.
// tab: 28 entries x 2 uint4 (see piclim_core.cuh)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t step_env(Env &e, const uint4 *tab, uint32_t *scr, int ss, uint32_t rot, uint32_t loc, int L, int M,
                                             int &k, bool &board_changed) {
    k = 0; board_changed = false;
    if (e.head >= e.npieces) return F_NOPIECE;
    const uint32_t piece = queue_piece(e.q, e.head);                        // :356 pop(0)
    e.head += 1;
    const uint32_t t = (piece * 4 + (rot & 3u)) * 2;                        // :359 / :61
    const MoveOut m = place_general(e.col, scr, ss, tab[t], tab[t + 1], (int)loc);
    k = m.k; board_changed = !m.topout;
    return apply_outcome(e, m, L, M);
}

// ---------------------------------------------------------------------------------------------
// afterstates: slot (r, c) == clone(env).move(r, c); sink.put(slot, word, flags) with
//   word = dlines | holes << 8 | bumpiness << 16 | aggregate_height << 24
//
// Fast path (no row clears), per slot, all on packed bytes / halfwords:
//   y      = max_j(H[c+j] - bo_j)                          hard drop in height form (:424-433)
//   full   = A[c] & ((col[c+j] + cb_j << y) for j<4)       rows of the piece that become full (:382-386); A excludes old full rows
//   N      = y + to_j on the columns the shape covers, H[c+j] elsewhere (one byte-wise IMAD + one LOP3 select)
//   agg'   = agg + sum_j (N_j - H_j)                        one VABSDIFF4.U8.ACC
//   bump'  = bump - old pairs + new pairs                   one VABSDIFF4.U8.ACC + one VABSDIFF
//   holes' = agg' - (cells + 4)
// Slots whose placement completes a row are deferred to the general move below (rare per slot).
// ---------------------------------------------------------------------------------------------
// static byte-permute selectors: X = (H[c-1], n0, n1, n2) from (N4, Hprev); Y = (n0, n1, n2, n3) from N4;
// pairs that touch a column >= 10 are made degenerate (|a - a| = 0) so they never count.
__host__ __device__ constexpr uint32_t sel_x(int c) {
    return c == 0 ? 0x2100u : c <= 7 ? 0x2104u : c == 8 ? 0x1104u : 0x0004u;
}
__host__ __device__ constexpr uint32_t sel_y(int c) {
    return c <= 6 ? 0x3210u : c == 7 ? 0x2210u : c == 8 ? 0x1110u : 0x0000u;
}

// sum over the window pairs (H[c-1],n0),(n0,n1),(n1,n2),(n2,n3),(n3,H[c+4]) restricted to columns 0..9
template <int C>
__device__ __forceinline__ uint32_t window_pairs(uint32_t N4, uint32_t hprev, uint32_t hnext, uint32_t acc) {
    // X = (H[c-1], n0, n1, n2): for the interior columns that is N4 * 256 + H[c-1], one IMAD on the FMA pipe
    const uint32_t X = (C >= 1 && C <= 7) ? mad_fma_pipe(N4, 256u, hprev) : __byte_perm(N4, hprev, sel_x(C));
    const uint32_t Y = (C <= 6) ? N4 : __byte_perm(N4, 0u, sel_y(C));
    uint32_t s = vsad4_acc(X, Y, acc);
    if (C + 4 <= 9) s = __sad((int)(N4 >> 24), (int)hnext, s);
    return s;
}

template <int C, bool UNIFORM, class Sink>
__device__ __forceinline__ void slot_fast(Sink &sink, int r, const int (&H)[14], const uint32_t (&Hw)[COLS],
                                          const uint32_t (&col)[14],
                                          const uint32_t (&A)[COLS], const uint32_t (&Bb)[COLS], uint32_t agg, uint32_t K, uint32_t U,
                                          uint32_t flN, uint32_t flT, int w,
                                          int nb0, int nb1, int nb2, int nb3, uint32_t cb0, uint32_t cb1, uint32_t cb2, uint32_t cb3,
                                          uint32_t TO4, uint32_t cover, int thr, uint32_t one,
                                          uint32_t &word, uint32_t &fl, bool &pend, uint32_t &pmask, uint32_t &nsmask, int cmax_warp) {
    // Distinct-placements form: alias slots are not stored at all, so a column that no distinct placement of ANY piece reaches in
    // this rotation is skipped outright -- rotation 0: column 9; rotation 2: columns 8 and 9 (the only two-wide shape there is O,
    // whose rotation 2 is an alias); rotation 3: column 9 (the one-wide vertical I is an alias there).  36 of 40 slots are left.
    if constexpr (Sink::RAGGED) {
        if (C >= 8 && C > cmax_warp) return;
    } else
    // No lane of this warp fits a shape at column C: only the clamp alias.  Without piece-sorted warps (UNIFORM) that is known
    // statically for column 9 of rotations 0 and 2, where the narrowest shape of any piece is two wide (cmax_warp = 8).
    if (UNIFORM ? (C > 6 && C > cmax_warp) : (C == 9 && cmax_warp < 9)) {
        if constexpr (SinkValueOnly<Sink>::value) return;         // (an alias: same value as the slot it repeats, higher number)
        if constexpr (Sink::PACKED) { word |= F_ALIAS << 3; if (!pend) sink.put_packed_col(C, word); }
        else { fl |= F_ALIAS; if (!pend) sink.put(r * 10 + C, word, fl); }
        if (pend) nsmask |= 1u << C;
        if constexpr (Sink::PACKED) sink.next_col();
        return;
    }
    // hard drop in height form.  nb_j = -bo_j is negated once per rotation: nvcc 12.9 / ptxas drops the operand negation when
    // it folds `H - bo` terms with a constant-zero H (the padding columns) into the 3-input VIMNMX3 (found by the GPU parity
    // tests: +64 instead of -64 won the max), so no negation is left for it to fold.
    // For C <= 6 all four heights are live registers: the sums H + nb are taken on the FMA pipe as H * one + nb, `one` being
    // the value 1 from a kernel parameter -- opaque to ptxas, which would otherwise fuse each sum into an ALU-pipe
    // VIADDMNMX -- and the four-way max is VIMNMX3 + VIMNMX: 2 ALU-pipe instructions instead of 3.
    int y;
    if (C <= 6) {
        const int t0 = (int)mad_fma_pipe((uint32_t)H[C], one, (uint32_t)nb0), t1 = (int)mad_fma_pipe((uint32_t)H[C + 1], one, (uint32_t)nb1);
        const int t2 = (int)mad_fma_pipe((uint32_t)H[C + 2], one, (uint32_t)nb2), t3 = (int)mad_fma_pipe((uint32_t)H[C + 3], one, (uint32_t)nb3);
        y = max(max(max(t0, t1), t2), t3);
    } else {
        y = max(max(H[C] + nb0, H[C + 1] + nb1), __viaddmax_s32(H[C + 3], nb3, H[C + 2] + nb2));
    }
    const bool top = y > thr;
    const uint32_t pw = 1u << y;
    // the piece lands on empty cells, so `column | piece image` is `column + piece image`: one IMAD per window column
    // No mask for the piece's rows (:382-383) is needed: A excludes the rows that were full before the placement (see
    // afterstates_env_impl), every row the piece spans receives at least one of its cells -- so it was not full before -- and
    // no other row changes: the rows that are full now and were not before ARE the completed rows of the piece.
    const uint32_t full = A[C] & mad_fma_pipe(cb0, pw, col[C]) & mad_fma_pipe(cb1, pw, col[C + 1]) &
                          mad_fma_pipe(cb2, pw, col[C + 2]) & mad_fma_pipe(cb3, pw, col[C + 3]);
    // new heights of the window as bytes: y + to_j where the shape covers the column, the old height elsewhere
    const uint32_t N4 = (mad_fma_pipe((uint32_t)y, 0x01010101u, TO4) & cover) | (Hw[C] & ~cover);
    const uint32_t agg2 = vsad4_acc(N4, Hw[C], agg);
    const uint32_t b2 = window_pairs<C>(N4, (uint32_t)H[C > 0 ? C - 1 : 0], (uint32_t)H[C + 4 <= 13 ? C + 4 : 13], Bb[C]);
    uint32_t wnew = agg2 * 0x01000100u + K;
    wnew = b2 * 0x10000u + wnew;
    if constexpr (Sink::RAGGED) {
        // distinct placements only: no alias bookkeeping (no carried word / pend), a column the shape does not fit is simply not
        // stored; lanes whose rotation is an alias store to a dummy row (begin_rotation_ragged) and their pmask is dropped
        const bool pnew = full != 0u;
        if (C <= 6) {                                  // every width fits (w <= 4)
            if (!pnew) sink.template put_col<C>(wnew);
            if (top) sink.template put_col<C>(U);
            if (pnew) pmask = mad_fma_pipe(one, 1u << C, pmask);       // (bit C is not set yet: + is |, and the add runs on the FMA pipe)
        } else {
            const bool fits = C + w <= COLS;
            if (fits && !pnew) sink.template put_col<C>(wnew);
            if (fits && top) sink.template put_col<C>(U);
            if (fits && pnew) pmask = mad_fma_pipe(one, 1u << C, pmask);
        }
    } else if constexpr (Sink::PACKED) {
        // compact form: the flags ride in byte 0 (K and U already carry flN << 3 / flT << 3 for this rotation)
        if (C <= 5) {
            // Columns every width fits and no later column aliases: no select at all.  The no-clear word is stored unless
            // a row completes, then U overwrites it when the piece tops out (same thread, same address: ordered).  A slot
            // that is both (a row of an overflowing piece completes -- very rare) is deferred too; resolve_slot handles it.
            const bool pnew = full != 0u;
            if (!pnew) sink.put_packed_col(C, wnew);
            if (top) sink.put_packed_col_again(C, U);   // (a second, separately predicated store -- see GlobalSink)
            if (pnew) { pmask = mad_fma_pipe(one, 1u << C, pmask); nsmask |= 1u << C; }
            pend = pnew && !top;
        } else {
            const bool pnew = !top && full != 0u;
            wnew = top ? U : wnew;
            if (C == 6) { word = wnew; pend = pnew; }
            else {
                const bool fits = C + w <= COLS;
                word = fits ? wnew : (word | (F_ALIAS << 3));
                pend = fits ? pnew : pend;
            }
            if (pend && (C == 6 || C + w <= COLS)) pmask = mad_fma_pipe(one, 1u << C, pmask);
            if (!pend) sink.put_packed_col(C, word); else nsmask |= 1u << C;
        }
        sink.next_col();
    } else if constexpr (SinkValueOnly<Sink>::value) {
        // value only: a column the shape does not fit repeats an earlier slot (same value, higher number: it can never win the
        // lowest-slot tie-break), so it is simply not offered; nothing is carried from column to column
        const bool fits = C <= 6 || C + w <= COLS;
        const bool pnew = !top && full != 0u;
        if (fits && pnew) pmask = mad_fma_pipe(one, 1u << C, pmask);
        if (fits && !pnew) sink.put_value(r * 10 + C, wnew, top);
    } else {
        const bool pnew = !top && full != 0u;
        wnew = top ? U : wnew;
        const uint32_t fnew = top ? flT : flN;
        if (C <= 6) {                                  // every width fits (w <= 4)
            word = wnew; fl = fnew; pend = pnew;
        } else {                                       // loc clamps to 10 - w (:364): repeat the last fitting column
            const bool fits = C + w <= COLS;
            word = fits ? wnew : word;
            fl = fits ? fnew : (fl | F_ALIAS);
            pend = fits ? pnew : pend;
        }
        if (pend && (C <= 6 || C + w <= COLS)) pmask = mad_fma_pipe(one, 1u << C, pmask);
        if (!pend) sink.put(r * 10 + C, word, fl); else nsmask |= 1u << C;
    }
}

// What a deferred (row-completing) slot needs besides the ten columns
struct PendingCtx { unsigned long long mask; uint32_t piece, cells, lines, fl_noclear; };

// One deferred slot: the general move on a copy of the columns, features of the cleared board, and the stores to the
// slot and to every slot that aliases it (rot + k * n_rot, and loc > 10 - w when the slot is the last fitting column).
template <class Sink>
__device__ __forceinline__ void resolve_slot(const uint32_t (&cols)[COLS], const PendingCtx &cx, int s, const uint4 *tab,
                                             uint32_t *scr, int ss, int L, Sink &sink) {
    const int r = s / 10, c = s - 10 * r;
    const uint4 o = tab[(cx.piece * 4 + r) * 2], ob = tab[(cx.piece * 4 + r) * 2 + 1];
    uint32_t x[COLS];
#pragma unroll
    for (int k = 0; k < COLS; ++k) x[k] = cols[k];
    const MoveOut m = place_general(x, scr, ss, o, ob, c);
    // (a deferred slot may turn out to be a top-out: board unchanged, :372-374)
    const uint32_t f3 = board_features(x, m.topout ? (int)cx.cells : (int)cx.cells + 4 - 10 * m.k);
    const uint32_t word = (uint32_t)m.k | (f3 << 8);
    const uint32_t fl = m.topout ? F_TOPOUT :
                        m.k == 0 ? cx.fl_noclear : (((int)cx.lines + m.k >= L) ? F_WIN : cx.fl_noclear);   // :389-391, :415-422
    if constexpr (Sink::RAGGED) {
        sink.put_canon((int)orient_rot_base(ob) + c, word | (fl << 3));       // deferred slots are canonical placements
    } else {
        const int nrot = orient_nrot(o), w = orient_w(o);
        const int cend = (c == COLS - w) ? COLS - 1 : c;
        for (int r2 = r; r2 < 4; r2 += nrot)
            for (int c2 = c; c2 <= cend; ++c2)
                sink.put(r2 * 10 + c2, word, fl | ((r2 != r || c2 != c) ? F_ALIAS : 0u));
    }
}

// UNIFORM = true is the variant for warps whose lanes were sorted by piece (afterstates_sorted_kernel): every lane of
// the warp must call it (no early exit), rotations that are aliases for every lane (r >= max n_rot in the warp) and
// columns no lane can reach (c > max(10 - w)) are not enumerated but copied from the slot they alias -- on average
// 23 of the 40 slots are distinct placements.  Needs a sink with copy(dst, src, extra_flags).
// [r_begin, r_end) restricts the enumeration to some rotations (the small-batch kernel gives each rotation of an env
// to its own thread); deferred line-clear slots still write every alias row of the rotation that owns them.
// `defer` != nullptr: the deferred slots are not resolved here but returned (mask + context), so that a CTA can pool
// them and resolve them with full warps (see afterstates_kernel).
template <bool UNIFORM, class Sink>
__device__ __forceinline__ void afterstates_env_impl(const Env &e, const uint4 *tab, uint32_t *scr, int ss, int L, int M, Sink &sink,
                                                     int r_begin = 0, int r_end = 4, PendingCtx *defer = nullptr, uint32_t one = 1u) {
    if (defer) defer->mask = 0ull;
    const bool nopiece = e.head >= e.npieces;
    static_assert(!(UNIFORM && Sink::RAGGED), "the distinct-placements form is not built for piece-sorted warps");
    if (!UNIFORM && nopiece) {
        if constexpr (!Sink::RAGGED)                    // (distinct-placements form: an env without a piece has an empty run)
            for (int s = r_begin * 10; s < r_end * 10; ++s) sink.put(s, 0u, F_NOPIECE);
        return;
    }
    // a lane without a piece walks through as an O piece (its stores are overwritten at the end)
    const uint32_t piece = nopiece ? 6u : queue_piece(e.q, e.head);
    int nrot_warp = 4;
    if constexpr (UNIFORM) nrot_warp = (int)__reduce_max_sync(0xFFFFFFFFu, (uint32_t)orient_nrot(tab[piece * 8]));

    // ---- per-env precompute, shared by all 40 slots.  Indices 10..13 are neutral padding (full columns for
    //      the AND, height 0); every use of them in a pair sum is statically excluded.
    uint32_t col[14]; int H[14];
    uint32_t cells = 0;
#pragma unroll
    for (int k = 0; k < COLS; ++k) { col[k] = e.col[k]; H[k] = col_height(col[k]); cells += __popc(col[k]); }
#pragma unroll
    for (int k = COLS; k < 14; ++k) { col[k] = COL_FULL; H[k] = 0; }
    const uint32_t HB0 = H[0] | (H[1] << 8) | (H[2] << 16) | (H[3] << 24);
    const uint32_t HB1 = H[4] | (H[5] << 8) | (H[6] << 16) | (H[7] << 24);
    const uint32_t HB2 = H[8] | (H[9] << 8);
    uint32_t Hw[COLS];                               // Hw[c] = bytes (H[c], H[c+1], H[c+2], H[c+3])
    Hw[0] = HB0; Hw[4] = HB1; Hw[8] = HB2;
    Hw[1] = __byte_perm(HB0, HB1, 0x4321); Hw[2] = __byte_perm(HB0, HB1, 0x5432); Hw[3] = __byte_perm(HB0, HB1, 0x6543);
    Hw[5] = __byte_perm(HB1, HB2, 0x4321); Hw[6] = __byte_perm(HB1, HB2, 0x5432); Hw[7] = __byte_perm(HB1, HB2, 0x6543);
    Hw[9] = HB2 >> 8;
    const uint32_t agg = vsad4_acc(HB2, 0u, vsad4_acc(HB1, 0u, vsad4_acc(HB0, 0u, 0u)));
    // Running sums of the height steps: P[k] = sum over j < k of |H[j] - H[j+1]| (one abs-diff-accumulate each), so bumpiness = P[9]
    // and the steps a placement at column c can change -- those between columns c-1 .. c+4 -- are P[min(c+4, 9)] - P[max(c-1, 0)].
    uint32_t P[COLS];
    P[0] = 0u;
#pragma unroll
    for (int k = 1; k < COLS; ++k) P[k] = (uint32_t)__sad(H[k - 1], H[k], P[k - 1]);
    const uint32_t bump = P[9];
    uint32_t Bb[COLS];                               // bump minus the pairs a placement at c can change
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
        Bb[c] = bump + P[c > 0 ? c - 1 : 0] - P[c + 4 < 9 ? c + 4 : 9];
#if defined(__CUDA_ARCH__)
        asm volatile("" : "+r"(Bb[c]));     // (kept in registers: re-deriving them from the running sums would cost an add per slot)
#endif
    }
    uint32_t pre[11], suf[11], A[COLS];              // A[c] = AND of the columns outside [c, c+3]
    pre[0] = COL_FULL; suf[10] = COL_FULL;
#pragma unroll
    for (int k = 0; k < COLS; ++k) pre[k + 1] = pre[k] & col[k];
#pragma unroll
    for (int k = COLS - 1; k >= 0; --k) suf[k] = suf[k + 1] & col[k];
    // ... and not among the rows that are full already (pre[10]; an untouched full row stays, :382-383, and must not count)
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
        A[c] = pre[c] & suf[c + 4 < 10 ? c + 4 : 10] & ~pre[10];
#if defined(__CUDA_ARCH__) && !defined(TPL_NO_PIN_A)
        asm volatile("" : "+r"(A[c]));      // keep the ten masks in registers: re-deriving them from pre / suf costs a LOP3 per slot
#endif
    }

    const uint32_t U = ((agg - cells) << 8) | (bump << 16) | (agg << 24);   // unchanged board (top-out slots)
    const uint32_t K = 0u - ((cells + 4u) << 8);                             // holes' = agg' - (cells + 4)
    const uint32_t fl_noclear = ((int)e.moves + 1 >= M) ? F_LOSE : 0u;       // :389-391
    unsigned long long pending = 0ull, notstored = 0ull;
    if constexpr (SinkValueOnly<Sink>::value) sink.begin_env(U, fl_noclear);

#ifndef TPL_ROT_UNROLL
#define TPL_ROT_UNROLL 1            // tuning knob: unrolling the rotation loop was measured and does not pay
#endif
    constexpr int rot_unroll = TPL_ROT_UNROLL;
#pragma unroll rot_unroll
    for (int r = r_begin; r < (UNIFORM ? nrot_warp : r_end); ++r) {
        const uint4 o = tab[(piece * 4 + r) * 2];
        const int w = orient_w(o);
        // widest reach of any piece in this rotation: the only one-wide shape is the vertical I, rotations 1 and 3 (game/tetris.py:25-55)
        int cmax_warp = (r & 1) ? 9 : 8;
        if constexpr (Sink::RAGGED) cmax_warp = (r == 1) ? 9 : (r == 2) ? 7 : 8;        // see slot_fast
        if constexpr (UNIFORM) cmax_warp = (int)__reduce_max_sync(0xFFFFFFFFu, (uint32_t)(COLS - w));
        const uint4 *wide = tab + TAB_COMPACT4 + (piece * 4 + r) * 3;              // the same facts, one register each
        const uint4 wn = wide[0], wc = wide[1], wm = wide[2];
        const int nb0 = (int)wn.x, nb1 = (int)wn.y, nb2 = (int)wn.z, nb3 = (int)wn.w;
        const uint32_t cb0 = wc.x, cb1 = wc.y, cb2 = wc.z, cb3 = wc.w;
        const uint32_t afl = orient_alias(o) ? F_ALIAS : 0u;
        const bool canon = afl == 0u;               // deferred slots are resolved once, from the canonical rotation
        const uint32_t flN = fl_noclear | afl, flT = F_TOPOUT | afl;
        const uint32_t Kr = Sink::PACKED ? K + (flN << 3) : K, Ur = Sink::PACKED ? (U | (flT << 3)) : U;
        uint32_t word = 0, fl = 0, pmask = 0, nsmask = 0; bool pend = false;
        if constexpr (Sink::RAGGED) sink.begin_rotation_ragged(canon, orient_rot_base(tab[(piece * 4 + r) * 2 + 1]));
        else if constexpr (Sink::PACKED) sink.begin_rotation(r);
#define TPL_SLOT(C) slot_fast<C, UNIFORM>(sink, r, H, Hw, col, A, Bb, agg, Kr, Ur, flN, flT, w, nb0, nb1, nb2, nb3, cb0, cb1, cb2, cb3, \
                                 wm.x, wm.y, (int)wm.w, one, word, fl, pend, pmask, nsmask, cmax_warp);
        TPL_SLOT(0) TPL_SLOT(1) TPL_SLOT(2) TPL_SLOT(3) TPL_SLOT(4) TPL_SLOT(5) TPL_SLOT(6) TPL_SLOT(7) TPL_SLOT(8) TPL_SLOT(9)
#undef TPL_SLOT
        if (canon) pending |= (unsigned long long)pmask << (10 * r);
        if constexpr (UNIFORM) notstored |= (unsigned long long)nsmask << (10 * r);
    }

    // ---- deferred line-clear slots
    PendingCtx cx{pending, piece, cells, e.lines, fl_noclear};
    if (!defer)
        while (pending) {
            const int s = __ffsll((long long)pending) - 1;
            pending &= pending - 1ull;
            resolve_slot(e.col, cx, s, tab, scr, ss, L, sink);
        }

    if constexpr (UNIFORM) {
        // rotations >= nrot_warp are aliases for every lane (:61): copy them from rot % n_rot.  Slots the fast path
        // left to the resolver are skipped: resolve_slot stores every alias row of a deferred slot itself.
        const int nrot = orient_nrot(tab[piece * 8]);
        for (int r = nrot_warp; r < 4; ++r)
            for (int c = 0; c < COLS; ++c) {
                const int src = (r % nrot) * 10 + c;
                if (!defer || !((notstored >> src) & 1ull)) sink.copy(r * 10 + c, src, F_ALIAS);
            }
        if (nopiece) {
            for (int s = 0; s < 40; ++s) sink.put(s, 0u, F_NOPIECE);
            cx.mask = 0ull;
        }
    }
    if (defer) *defer = cx;
}

template <class Sink>
__device__ __forceinline__ void afterstates_env(const Env &e, const uint4 *tab, uint32_t *scr, int ss, int L, int M, Sink &sink,
                                                int r_begin = 0, int r_end = 4, PendingCtx *defer = nullptr, uint32_t one = 1u) {
    afterstates_env_impl<false>(e, tab, scr, ss, L, M, sink, r_begin, r_end, defer, one);
}

// ---------------------------------------------------------------------------------------------
// one random-agent move with auto-reset (the body of the fused rollout loop)
//   acc[8] += {episodes, wins, top-outs, move-limit losses, lines, moves placed, steps, resets}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rollout_random_step(Env &e, uint32_t &ep, uint32_t &t, uint32_t (&acc)[8], const uint4 *tab,
                                                    uint32_t *scr, int ss, const uint4 *__restrict__ pool, int K, uint64_t seed, uint64_t env,
                                                    int gen_count, int L, int M) {
    if (!refill_queue(e, seed, env, ep, gen_count) && (e.state != S_RUNNING || e.head >= e.npieces)) {
        ep += 1; t = 0; acc[7] += 1;
        install_config(e, pool, config_index(seed, env, ep, K), seed, env, ep, gen_count);
    }
    const uint4 rw = rng_words(seed, env, ep, STREAM_ACTION, t);
    const uint32_t rot = rw.x & 3u, loc = __umulhi(rw.y, 10u);
    int k; bool changed;
    const uint32_t fl = step_env(e, tab, scr, ss, rot, loc, L, M, k, changed);
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
// W16: the four feature weights fit int16, so the value of a slot is two dot-product instructions on the packed feature word
// (IDP.2A: two s16 weights x two u8 features each) instead of four byte extractions and four multiply-adds.
template <bool W16>
__device__ __forceinline__ int greedy_value(int w0, int w1, int w2, int w3, int w4, int w5, uint32_t word, uint32_t fl) {
    int v;
    if constexpr (W16) {
        v = dp2a_hi((uint32_t)w2 & 0xFFFFu | ((uint32_t)w3 << 16), word, dp2a_lo((uint32_t)w0 & 0xFFFFu | ((uint32_t)w1 << 16), word, 0));
    } else {
        v = w0 * (int)(word & 0xFFu) + w1 * (int)((word >> 8) & 0xFFu) + w2 * (int)((word >> 16) & 0xFFu) +
            w3 * (int)(word >> 24);
    }
    if (fl & F_WIN) v += w4;
    if (fl & (F_LOSE | F_TOPOUT)) v += w5;
    return v;
}

// INORDER: every put comes with a higher slot number than the one before (the enumeration's own order, when the deferred slots
// are resolved elsewhere): "first strictly greater wins" then IS the lowest-slot tie-break, one compare less per slot.
// ... and the fast path then hands over values only (VALUE_ONLY): without a completed row a placement cannot win, so its flags are
// either those of a top-out -- value of the unchanged board + w5, one constant per env (vU) -- or the env's no-clear flags (nothing
// or LOSE: the constant bN, which rides as the accumulator of the dot products): two dot products and one select per slot.
template <bool W16, bool INORDER = false>
struct GreedySinkT {
    static constexpr bool PACKED = false, RAGGED = false, VALUE_ONLY = INORDER;
    int w0, w1, w2, w3, w4, w5;
    int best, best_slot;
    int vU = 0, bN = 0;
    __device__ __forceinline__ void put(int slot, uint32_t word, uint32_t fl) {
        const int v = greedy_value<W16>(w0, w1, w2, w3, w4, w5, word, fl);
        if (INORDER ? v > best : (v > best || (v == best && slot < best_slot))) { best = v; best_slot = slot; }
    }
    __device__ __forceinline__ void begin_env(uint32_t U, uint32_t fl_noclear) {
        vU = greedy_value<W16>(w0, w1, w2, w3, w4, w5, U, F_TOPOUT);
        bN = (fl_noclear & (F_LOSE | F_TOPOUT)) ? w5 : 0;
    }
    __device__ __forceinline__ void put_value(int slot, uint32_t wnew, bool top) {
        int v;
        if constexpr (W16) v = dp2a_hi((uint32_t)w2 & 0xFFFFu | ((uint32_t)w3 << 16), wnew, dp2a_lo((uint32_t)w0 & 0xFFFFu | ((uint32_t)w1 << 16), wnew, bN));
        else v = greedy_value<false>(w0, w1, w2, w3, w4, w5, wnew, 0u) + bN;
        v = top ? vU : v;
        if (v > best) { best = v; best_slot = slot; }
    }
};

// records the first put only: resolve_slot stores the canonical slot first, its aliases (equal value, higher index) after
struct FirstPutSink {
    static constexpr bool PACKED = false, RAGGED = false;
    uint32_t word, fl; bool got;
    __device__ __forceinline__ void put(int, uint32_t w, uint32_t f) { if (!got) { word = w; fl = f; got = true; } }
};

using GreedySink = GreedySinkT<false>;

struct GreedyWeights { int w[6]; };

template <bool W16 = false>
__device__ __forceinline__ void rollout_greedy_step(Env &e, uint32_t &ep, uint32_t &t, uint32_t (&acc)[8], const uint4 *tab,
                                                    uint32_t *scr, int ss, const uint4 *__restrict__ pool, int K, uint64_t seed, uint64_t env,
                                                    int gen_count, int L, int M, const GreedyWeights &gw) {
    if (!refill_queue(e, seed, env, ep, gen_count) && (e.state != S_RUNNING || e.head >= e.npieces)) {
        ep += 1; t = 0; acc[7] += 1;
        install_config(e, pool, config_index(seed, env, ep, K), seed, env, ep, gen_count);
    }
    GreedySinkT<W16> sink{gw.w[0], gw.w[1], gw.w[2], gw.w[3], gw.w[4], gw.w[5], (int)0x80000000, 40};
    afterstates_env(e, tab, scr, ss, L, M, sink);
    const int slot = sink.best_slot < 40 ? sink.best_slot : 0;
    const uint32_t rot = (uint32_t)(slot / 10), loc = (uint32_t)(slot - 10 * (slot / 10));
    int k; bool changed;
    const uint32_t fl = step_env(e, tab, scr, ss, rot, loc, L, M, k, changed);
    t += 1;
    acc[6] += 1; acc[4] += (uint32_t)k; acc[5] += changed ? 1u : 0u;
    if (e.state != S_RUNNING) {
        acc[0] += 1;
        if (fl & F_WIN) acc[1] += 1; else if (fl & F_TOPOUT) acc[2] += 1; else acc[3] += 1;
    }
}

}  // namespace tpl
