// emul.cpp -- host emulation harness for the per-env device logic.  TEST INFRASTRUCTURE ONLY.
//
// Compiles csrc/piclim_core.cuh + piclim_env.cuh with g++ (tests/emul/host_shim.h stands in for the CUDA
// intrinsics) and runs the SAME per-env bodies the CUDA kernels run, one env after another, on host
// arrays.  It lets the CPU-only test tier check the bit-column algorithms, record layout, RNG and
// afterstate bookkeeping against the oracle before any GPU time is spent.  The product package never
// loads this library and has no CPU path; entry points mirror tpl_* (minus the stream argument).
#define TPL_HOST_EMUL 1
#include "../../reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200/csrc/piclim_env.cuh"

using namespace tpl;

static const uint4 *table() { return reinterpret_cast<const uint4 *>(&c_orient); }

struct HostSink {
    static constexpr bool PACKED = false, RAGGED = false;
    uint32_t *feats; uint8_t *flags; float *ff; size_t n, i;
    void put(int slot, uint32_t word, uint32_t fl) {
        const size_t o = (size_t)slot * n + i;
        if (feats) feats[o] = flags ? word : (word | (fl << 3));      // no flags array: pack them into byte 0
        if (flags) flags[o] = (uint8_t)fl;
        if (ff) { ff[4 * o] = (float)(word & 0xFF); ff[4 * o + 1] = (float)((word >> 8) & 0xFF);
                  ff[4 * o + 2] = (float)((word >> 16) & 0xFF); ff[4 * o + 3] = (float)(word >> 24); }
    }
};

// sink with copy(): the packed staging the sorted kernel uses (word | flags << 3), one row of 40 per env
struct PackedRowSink {
    static constexpr bool PACKED = true, RAGGED = false;
    uint32_t row[40];
    int rot;
    void begin_rotation(int r) { rot = r; }
    void put_packed_col(int c, uint32_t packed) { row[rot * 10 + c] = packed; }
    void put_packed_col_again(int c, uint32_t packed) { row[rot * 10 + c] = packed; }
    void next_col() {}
    void put_packed(int slot, uint32_t packed) { row[slot] = packed; }
    void put(int slot, uint32_t word, uint32_t fl) { row[slot] = word | (fl << 3); }
    void copy(int dst, int src, uint32_t extra) { row[dst] = row[src] | (extra << 3); }
};

// the kernels' compact-form sink (GlobalSink<0>): a cursor that walks along the columns of one rotation
struct PackedCursorSink {
    static constexpr bool PACKED = true, RAGGED = false;
    uint32_t row[40];
    int rot, colc;
    void begin_rotation(int r) { rot = r; colc = 0; }
    void put_packed_col(int, uint32_t packed) { row[rot * 10 + colc] = packed; }
    void put_packed_col_again(int, uint32_t packed) { row[rot * 10 + colc] = packed; }
    void next_col() { ++colc; }
    void put_packed(int slot, uint32_t packed) { row[slot] = packed; }
    void put(int slot, uint32_t word, uint32_t fl) { row[slot] = word | (fl << 3); }
};

// distinct-placements form: the run of one env (at most 34 words); alias rotations write to a dummy row
struct HostRaggedSink {
    static constexpr bool PACKED = true, RAGGED = true;
    uint32_t row[34 + 10];
    uint32_t *prow;
    void begin_rotation_ragged(bool canon, uint32_t rot_base) { prow = canon ? row + rot_base : row + 34; }
    template <int C> void put_col(uint32_t w) { prow[C] = w; }
    void put_canon(int idx, uint32_t w) { row[idx] = w; }
};

static uint32_t emul_distinct_env(const Env &e, int L, int M, int defer, uint32_t *rows, uint32_t &cursor, uint32_t run_base) {
    uint32_t piece = 7u, cnt = 0u;
    if (e.head < e.npieces) { piece = queue_piece(e.q, e.head); cnt = orient_run_len(table()[piece * 8 + 1]); }
    HostRaggedSink sink; for (int s = 0; s < 44; ++s) sink.row[s] = 0xDEADBEEFu;
    uint32_t scr[SCR_ROWS];
    if (!defer) afterstates_env(e, table(), scr, 1, L, M, sink);
    else {                                               // the kernels' way: deferred slots resolved afterwards, into the same run
        PendingCtx cx;
        afterstates_env(e, table(), scr, 1, L, M, sink, 0, 4, &cx);
        unsigned long long m = cx.mask;
        while (m) {
            const int s = __builtin_ffsll((long long)m) - 1; m &= m - 1ull;
            resolve_slot(e.col, cx, s, table(), scr, 1, L, sink);
        }
    }
    const uint32_t off = cursor;
    for (uint32_t j = 0; j < cnt; ++j) rows[off + j] = sink.row[j];
    cursor += cnt;
    return (off + run_base) | (piece << 29);
}

extern "C" {

int emul_afterstates_distinct(const void *state, int64_t stride, int n, uint32_t *rows, uint32_t *runs, uint32_t run_base, uint32_t *cursor,
                              int L, int M, int defer) {
    for (int64_t i = 0; i < n; ++i) {
        Env e; load_env((const uint4 *)state, stride, i, e);
        runs[i] = emul_distinct_env(e, L, M, defer, rows, *cursor, run_base);
    }
    return 0;
}

// afterstates through the warp-uniform (alias-skipping) variant, each env being its own one-lane "warp".
// defer != 0: the row-completing slots are returned by the enumeration and resolved afterwards, the way the
// CTA-pooled kernels do it.
int emul_afterstates_uniform(const void *state, int64_t stride, int n, uint8_t *feats_packed, int L, int M, int defer) {
    uint32_t *out = (uint32_t *)feats_packed;
    for (int64_t i = 0; i < n; ++i) {
        Env e; load_env((const uint4 *)state, stride, i, e);
        PackedRowSink sink; for (int s = 0; s < 40; ++s) sink.row[s] = 0xDEADBEEFu;
        uint32_t scr[SCR_ROWS];
        if (!defer) {
            afterstates_env_impl<true>(e, table(), scr, 1, L, M, sink);
        } else {
            PendingCtx cx;
            afterstates_env_impl<true>(e, table(), scr, 1, L, M, sink, 0, 4, &cx);
            unsigned long long m = cx.mask;
            while (m) {
                const int s = __builtin_ffsll((long long)m) - 1; m &= m - 1ull;
                resolve_slot(e.col, cx, s, table(), scr, 1, L, sink);
            }
        }
        for (int s = 0; s < 40; ++s) out[(size_t)s * n + i] = sink.row[s];
    }
    return 0;
}

// afterstates through the compact-form path of the persistent kernels (afterstates_env with a packed cursor sink), the
// row-completing slots deferred and resolved afterwards as the kernels' queues do
int emul_afterstates_cursor(const void *state, int64_t stride, int n, uint8_t *feats_packed, int L, int M) {
    uint32_t *out = (uint32_t *)feats_packed;
    for (int64_t i = 0; i < n; ++i) {
        Env e; load_env((const uint4 *)state, stride, i, e);
        PackedCursorSink sink; for (int s = 0; s < 40; ++s) sink.row[s] = 0xDEADBEEFu;
        uint32_t scr[SCR_ROWS];
        PendingCtx cx;
        afterstates_env(e, table(), scr, 1, L, M, sink, 0, 4, &cx);
        for (unsigned long long m = cx.mask; m; m &= m - 1ull) resolve_slot(e.col, cx, __builtin_ffsll((long long)m) - 1, table(), scr, 1, L, sink);
        for (int s = 0; s < 40; ++s) out[(size_t)s * n + i] = sink.row[s];
    }
    return 0;
}

// afterstates with each rotation enumerated by a separate call (the small-batch kernel's thread mapping)
int emul_afterstates_split(const void *state, int64_t stride, int n, uint8_t *feats_packed, int L, int M) {
    for (int64_t i = 0; i < n; ++i)
        for (int r = 3; r >= 0; --r) {                     // any order must work
            Env e; load_env((const uint4 *)state, stride, i, e);
            HostSink sink{(uint32_t *)feats_packed, nullptr, nullptr, (size_t)n, (size_t)i};
            uint32_t scr[SCR_ROWS];
            afterstates_env(e, table(), scr, 1, L, M, sink, r, r + 1);
        }
    return 0;
}

int emul_pack(void *out, int64_t stride, int aos, int n, const uint16_t *rows, const uint8_t *pieces, int pstride,
              const uint8_t *npieces, const int32_t *lines, const int32_t *moves, const int8_t *st, const uint8_t *head) {
    for (int64_t i = 0; i < n; ++i) {
        Env e;
        rows_to_cols(rows + i * ROWS, e.col);
        const int np = npieces[i] < 42 ? npieces[i] : 42;
        pack_queue(pieces + i * pstride, np, e.q);
        e.lines = lines ? (uint32_t)lines[i] : 0u; e.moves = moves ? (uint32_t)moves[i] : 0u;
        e.state = st ? (uint32_t)st[i] : 0u; e.head = head ? head[i] : 0u; e.npieces = (uint32_t)np;
        if (aos) store_env((uint4 *)out + 4 * i, 1, 0, e); else store_env((uint4 *)out, stride, i, e);
    }
    return 0;
}

int emul_unpack(const void *state, int64_t stride, int n, uint16_t *rows, uint8_t *cur, uint8_t *next, int32_t *lines,
                int32_t *moves, int8_t *sto, uint8_t *head, uint8_t *npieces, uint8_t *queue) {
    for (int64_t i = 0; i < n; ++i) {
        Env e; load_env((const uint4 *)state, stride, i, e);
        if (rows) cols_to_rows(e.col, rows + i * ROWS);
        if (cur) cur[i] = e.head < e.npieces ? (uint8_t)queue_piece(e.q, e.head) : (uint8_t)255;
        if (next) next[i] = e.head + 1 < e.npieces ? (uint8_t)queue_piece(e.q, e.head + 1) : (uint8_t)255;
        if (lines) lines[i] = (int32_t)e.lines;
        if (moves) moves[i] = (int32_t)e.moves;
        if (sto) sto[i] = (int8_t)e.state;
        if (head) head[i] = (uint8_t)e.head;
        if (npieces) npieces[i] = (uint8_t)e.npieces;
        if (queue) for (int p = 0; p < 42; ++p) queue[i * 42 + p] = (uint8_t)queue_piece(e.q, p);
    }
    return 0;
}

int emul_reset_from_pool(void *state, int64_t stride, int n, const void *pool, int K, const int32_t *idx, const uint8_t *mask,
                         int mode, uint32_t *episode, uint32_t *tstep, uint64_t seed, uint64_t env_base, int gen_count) {
    uint4 *st = (uint4 *)state;
    for (int64_t i = 0; i < n; ++i) {
        if (mode == 1 && !mask[i]) continue;
        uint32_t ep = episode ? episode[i] : 0u;
        if (mode == 2) {
            const uint4 d = st[3 * stride + i];
            const uint32_t s = d.w & 0xFFu, head = (d.w >> 8) & 0xFFu, np = (d.w >> 16) & 0xFFu;
            if (s == S_RUNNING && head < np) continue;
            if (gen_count > QUEUE_PIECES && s == S_RUNNING) {
                Env e; load_env(st, stride, i, e);
                if (refill_queue(e, seed, env_base + (uint64_t)i, ep, gen_count)) { store_env(st, stride, i, e); continue; }
            }
        }
        if (mode == 2 || (mode == 1 && !idx)) { ep += 1; if (episode) episode[i] = ep; }
        if (tstep) tstep[i] = 0u;
        uint32_t k;
        if (idx) { const int32_t v = idx[i]; k = (uint32_t)(v < 0 ? 0 : (v >= K ? K - 1 : v)); }
        else k = config_index(seed, env_base + (uint64_t)i, ep, K);
        Env e; install_config(e, (const uint4 *)pool, k, seed, env_base + (uint64_t)i, ep, gen_count);
        store_env(st, stride, i, e);
    }
    return 0;
}

int emul_step(void *state, int64_t stride, int n, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags,
              int8_t *sto, int L, int M) {
    uint4 *st = (uint4 *)state;
    for (int64_t i = 0; i < n; ++i) {
        Env e; load_env(st, stride, i, e);
        int k; bool changed;
        uint32_t scr[SCR_ROWS];
        const uint32_t fl = step_env(e, table(), scr, 1, rot[i], loc[i], L, M, k, changed);
        if (changed) {
            st[i] = make_uint4(e.col[0], e.col[1], e.col[2], e.col[3]);
            st[stride + i] = make_uint4(e.col[4], e.col[5], e.col[6], e.col[7]);
            st[2 * stride + i] = make_uint4(e.col[8], e.col[9], e.q[0], e.q[1]);
        }
        if (!(fl & F_NOPIECE)) st[3 * stride + i] = pack_meta(e);
        if (dlines) dlines[i] = (int8_t)k;
        if (flags) flags[i] = (uint8_t)fl;
        if (sto) sto[i] = (int8_t)e.state;
    }
    return 0;
}

int emul_afterstates(const void *state, int64_t stride, int n, uint8_t *feats, uint8_t *flags, float *ff, int L, int M) {
    for (int64_t i = 0; i < n; ++i) {
        Env e; load_env((const uint4 *)state, stride, i, e);
        HostSink sink{(uint32_t *)feats, flags, ff, (size_t)n, (size_t)i};
        uint32_t scr2[SCR_ROWS];
        afterstates_env(e, table(), scr2, 1, L, M, sink);
    }
    return 0;
}

int emul_step_observe(void *state, int64_t stride, int n, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags,
                      int8_t *sto, long long *stats, const void *pool, int K, uint32_t *episode, uint32_t *tstep, uint64_t seed,
                      uint64_t env_base, int gen_count, uint8_t *feats, uint8_t *aflags, float *ff, int L, int M,
                      uint32_t *drows, uint32_t *druns, uint32_t *dcursor) {
    uint4 *st = (uint4 *)state;
    for (int64_t i = 0; i < n; ++i) {
        Env e; load_env(st, stride, i, e);
        const uint32_t was = e.state;
        int k; bool changed;
        uint32_t scr[SCR_ROWS];
        const uint32_t fl = step_env(e, table(), scr, 1, rot[i], loc[i], L, M, k, changed);
        if (dlines) dlines[i] = (int8_t)k;
        if (flags) flags[i] = (uint8_t)fl;
        if (sto) sto[i] = (int8_t)e.state;
        if (stats) {
            stats[6] += 1; stats[4] += k; stats[5] += changed ? 1 : 0;
            if (was == S_RUNNING && e.state != S_RUNNING) {
                stats[0] += 1;
                if (fl & F_WIN) stats[1] += 1; else if (fl & F_TOPOUT) stats[2] += 1; else stats[3] += 1;
            }
        }
        if (refill_queue(e, seed, env_base + (uint64_t)i, episode ? episode[i] : 0u, gen_count)) {
        } else if (pool && (e.state != S_RUNNING || e.head >= e.npieces)) {
            uint32_t ep = episode ? episode[i] + 1u : 1u;
            if (episode) episode[i] = ep;
            if (tstep) tstep[i] = 0u;
            install_config(e, (const uint4 *)pool, config_index(seed, env_base + (uint64_t)i, ep, K), seed, env_base + (uint64_t)i, ep, gen_count);
            if (stats) stats[7] += 1;
        }
        store_env(st, stride, i, e);
        if (drows) { druns[i] = emul_distinct_env(e, L, M, (int)(i & 1), drows, *dcursor, 0u); continue; }
        HostSink sink{(uint32_t *)feats, aflags, ff, (size_t)n, (size_t)i};
        uint32_t scr2[SCR_ROWS];
        afterstates_env(e, table(), scr2, 1, L, M, sink);
    }
    return 0;
}

int emul_gen_pieces(uint8_t *out, int n, int count, uint64_t seed, uint64_t env_base, const uint32_t *episode, uint32_t episode0) {
    for (int64_t i = 0; i < n; ++i)
        for (int base = 0, block = 0; base < count; base += QUEUE_PIECES, ++block) {
            uint32_t q[4];
            const int c = count - base < QUEUE_PIECES ? count - base : QUEUE_PIECES;
            gen_queue(seed, env_base + (uint64_t)i, episode ? episode[i] : episode0, c, q, (uint32_t)block);
            for (int p = 0; p < c; ++p) out[i * count + base + p] = (uint8_t)queue_piece(q, p);
        }
    return 0;
}

static int rollout(bool greedy, void *state, int64_t stride, int n, const void *pool, int K, uint32_t *episode, uint32_t *tstep,
                   long long *stats, int steps, const int32_t *w6, uint64_t seed, uint64_t env_base, int gen_count, int L, int M) {
    uint4 *st = (uint4 *)state;
    GreedyWeights gw{};
    if (greedy) for (int q = 0; q < 6; ++q) gw.w[q] = w6[q];
    for (int64_t i = 0; i < n; ++i) {
        uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const uint64_t env = env_base + (uint64_t)i;
        Env e; load_env(st, stride, i, e);
        uint32_t ep = episode[i], t = tstep[i];
        uint32_t scr[SCR_ROWS];
        for (int s = 0; s < steps; ++s) {
            if (greedy) rollout_greedy_step(e, ep, t, acc, table(), scr, 1, (const uint4 *)pool, K, seed, env, gen_count, L, M, gw);
            else rollout_random_step(e, ep, t, acc, table(), scr, 1, (const uint4 *)pool, K, seed, env, gen_count, L, M);
        }
        store_env(st, stride, i, e);
        episode[i] = ep; tstep[i] = t;
        for (int q = 0; q < 8; ++q) stats[q] += acc[q];
    }
    return 0;
}

int emul_rollout_random(void *state, int64_t stride, int n, const void *pool, int K, uint32_t *episode, uint32_t *tstep,
                        long long *stats, int steps, uint64_t seed, uint64_t env_base, int gen_count, int L, int M) {
    return rollout(false, state, stride, n, pool, K, episode, tstep, stats, steps, nullptr, seed, env_base, gen_count, L, M);
}
int emul_rollout_greedy(void *state, int64_t stride, int n, const void *pool, int K, uint32_t *episode, uint32_t *tstep,
                        long long *stats, int steps, const int32_t *w6, uint64_t seed, uint64_t env_base, int gen_count, int L, int M) {
    return rollout(true, state, stride, n, pool, K, episode, tstep, stats, steps, w6, seed, env_base, gen_count, L, M);
}

}  // extern "C"
