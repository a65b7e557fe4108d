// piclim_kernels.cu -- sm_100a kernels + the device-pointer half of the C ABI (include/tetris_piclim.h).
//
// Thread mapping: one thread per env, one warp per group of 32 consecutive envs.  The env state is four
// planes of 16-byte chunks, so every warp-level load/store of a chunk is one fully coalesced 512-byte,
// 128-bit-per-lane transaction and no shared-memory staging or transposition is needed.
#include "piclim_env.cuh"
#include "../../include/tetris_piclim.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace tpl {

constexpr int THREADS = 128;

static thread_local char g_err[256] = "";
static std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}
long long launches() { return g_launches.load(); }
const char *last_error() { return g_err; }

__device__ __forceinline__ void load_table(uint4 *s_tab) {
    if (threadIdx.x < TAB_WORDS4) s_tab[threadIdx.x] = reinterpret_cast<const uint4 *>(c_orient)[threadIdx.x];
    __syncthreads();
}

// =================================================================================================
// pack / unpack: boundary format (20 x u16 bitrows, piece bytes) <-> env records
// =================================================================================================
__global__ void __launch_bounds__(THREADS)
pack_kernel(uint4 *out, int64_t stride, int aos, int n, const uint16_t *__restrict__ rows,
            const uint8_t *__restrict__ pieces, int pstride, const uint8_t *__restrict__ npieces,
            const int32_t *__restrict__ lines, const int32_t *__restrict__ moves, const int8_t *__restrict__ st,
            const uint8_t *__restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    Env e;
    rows_to_cols(rows + i * ROWS, e.col);
    const int np = min((int)npieces[i], TPL_MAX_PIECES);
    pack_queue(pieces + i * pstride, np, e.q);
    e.lines = lines ? (uint32_t)lines[i] : 0u;
    e.moves = moves ? (uint32_t)moves[i] : 0u;
    e.state = st ? (uint32_t)st[i] : 0u;
    e.head = head ? head[i] : 0u;
    e.npieces = (uint32_t)np;
    if (aos) store_env(out + 4 * i, 1, 0, e);       // record k = 4 consecutive chunks
    else store_env(out, stride, i, e);
}

__global__ void __launch_bounds__(THREADS)
unpack_kernel(const uint4 *__restrict__ st, int64_t stride, int n, uint16_t *rows, uint8_t *cur, uint8_t *next,
              int32_t *lines, int32_t *moves, int8_t *sto, uint8_t *head, uint8_t *npieces, uint8_t *queue) {
    const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    Env e; load_env(st, stride, i, e);
    if (rows) cols_to_rows(e.col, rows + i * ROWS);
    if (cur) cur[i] = e.head < e.npieces ? (uint8_t)queue_piece(e.q, e.head) : (uint8_t)255;          // pieces[0] (:436)
    if (next) next[i] = e.head + 1 < e.npieces ? (uint8_t)queue_piece(e.q, e.head + 1) : (uint8_t)255; // pieces[1]
    if (lines) lines[i] = (int32_t)e.lines;
    if (moves) moves[i] = (int32_t)e.moves;
    if (sto) sto[i] = (int8_t)e.state;
    if (head) head[i] = (uint8_t)e.head;
    if (npieces) npieces[i] = (uint8_t)e.npieces;
    if (queue)
        for (int p = 0; p < TPL_MAX_PIECES; ++p) queue[i * TPL_MAX_PIECES + p] = (uint8_t)queue_piece(e.q, p);
}

// =================================================================================================
// reset from a prescribed-config pool (game/tetris.py:438-449, the "install a reset point" half)
// =================================================================================================
__global__ void __launch_bounds__(THREADS)
reset_kernel(uint4 *st, int64_t stride, int n, const uint4 *__restrict__ pool, int K, const int32_t *__restrict__ idx,
             const uint8_t *__restrict__ mask, int mode, uint32_t *episode, uint64_t seed, uint64_t env_base, int gen_count) {
    const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    if (mode == TPL_RESET_MASK && !mask[i]) return;
    uint32_t ep = episode ? episode[i] : 0u;
    if (mode == TPL_RESET_DONE) {
        const uint4 d = st[3 * stride + i];
        const uint32_t state = d.w & 0xFFu, head = (d.w >> 8) & 0xFFu, np = (d.w >> 16) & 0xFFu;
        if (state == S_RUNNING && head < np) return;
        ep += 1;
        if (episode) episode[i] = ep;
    }
    uint32_t k;
    if (idx) { const int32_t v = idx[i]; k = (uint32_t)(v < 0 ? 0 : (v >= K ? K - 1 : v)); }
    else k = config_index(seed, env_base + (uint64_t)i, ep, K);
    Env e;
    install_config(e, pool, k, seed, env_base + (uint64_t)i, ep, gen_count);
    store_env(st, stride, i, e);
}

// =================================================================================================
// step: Tetris.move for every env (game/tetris.py:354-422)
// =================================================================================================
// Episode statistics: per-thread counters -> warp shuffle-reduce -> shared-memory atomics -> one global atomic
// per counter and CTA.  (One global atomic per warp serialised ~33k same-address atomics per counter at 2^20
// envs and took 3x longer than the step kernel itself.)
__device__ __forceinline__ void flush_stats(const uint32_t (&loc)[8], unsigned long long *stats) {
    __shared__ unsigned int s_acc[8];
    if (threadIdx.x < 8) s_acc[threadIdx.x] = 0u;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, loc[q]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_acc[q], v);
    }
    __syncthreads();
    if (threadIdx.x < 8 && s_acc[threadIdx.x]) atomicAdd(stats + threadIdx.x, (unsigned long long)s_acc[threadIdx.x]);
}

__global__ void __launch_bounds__(THREADS)
step_kernel(uint4 *st, int64_t stride, int n, const uint8_t *__restrict__ rot, const uint8_t *__restrict__ loc,
            int8_t *dlines, uint8_t *flags, int8_t *sto, unsigned long long *stats, int L, int M) {
    __shared__ uint4 s_tab[TAB_WORDS4];
    load_table(s_tab);
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
        Env e; load_env(st, stride, i, e);
        const uint32_t was = e.state;
        int k; bool changed;
        const uint32_t fl = step_env(e, s_tab, rot[i], loc[i], L, M, k, changed);
        if (changed) {
            st[i] = make_uint4(e.col[0], e.col[1], e.col[2], e.col[3]);
            st[stride + i] = make_uint4(e.col[4], e.col[5], e.col[6], e.col[7]);
            st[2 * stride + i] = make_uint4(e.col[8], e.col[9], e.q[0], e.q[1]);
        }
        if (!(fl & F_NOPIECE)) st[3 * stride + i] = pack_meta(e);
        if (dlines) dlines[i] = (int8_t)k;
        if (flags) flags[i] = (uint8_t)fl;
        if (sto) sto[i] = (int8_t)e.state;
        acc[6] += 1; acc[4] += (uint32_t)k; acc[5] += changed ? 1u : 0u;
        if (was == S_RUNNING && e.state != S_RUNNING) {
            acc[0] += 1;
            if (fl & F_WIN) acc[1] += 1; else if (fl & F_TOPOUT) acc[2] += 1; else acc[3] += 1;
        }
    }
    if (stats) flush_stats(acc, stats);
}

// =================================================================================================
// afterstates: slot (r, c) == clone(env).move(r, c), features on the post-move board
// =================================================================================================
// Output modes: 0 = words only, flags packed into byte 0 (dlines | flags << 3): the compact 160 B/env form;
//               1 = words (byte 0 = dlines) + separate flags array: the 200 B/env parity form;
//               2 = float4 features + flags (value-net input rows);   3 = all three.
// All arrays are slot-major [40][n]; element (slot, i) sits at offset slot * n + i in each of them, so one
// 32-bit offset serves every array and a warp's store of one slot is one contiguous 128-byte line.
template <int MODE>
struct GlobalSink {
    uint32_t *words; uint8_t *flags; float4 *ff;      // already offset by the env index
    uint32_t n;
    __device__ __forceinline__ void put(int slot, uint32_t word, uint32_t fl) {
        const uint32_t o = (uint32_t)slot * n;
        if (MODE == 0) words[o] = word | (fl << 3);
        if (MODE == 1 || MODE == 3) words[o] = word;
        if (MODE >= 1) flags[o] = (uint8_t)fl;
        if (MODE >= 2)
            ff[o] = make_float4((float)(word & 0xFFu), (float)((word >> 8) & 0xFFu), (float)((word >> 16) & 0xFFu),
                                (float)(word >> 24));
    }
};

template <int MODE>
__global__ void __launch_bounds__(THREADS)
afterstates_kernel(const uint4 *__restrict__ st, int64_t stride, int n, uint32_t *__restrict__ words,
                   uint8_t *__restrict__ flags, float4 *__restrict__ ff, int L, int M) {
    __shared__ uint4 s_tab[TAB_WORDS4];
    load_table(s_tab);
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS) {
        Env e; load_env(st, stride, i, e);
        GlobalSink<MODE> sink{words + i, flags + i, ff + i, (uint32_t)n};
        afterstates_env(e, s_tab, L, M, sink);
    }
}

// =================================================================================================
// counter-based 7-bag sequences
// =================================================================================================
__global__ void __launch_bounds__(THREADS)
gen_pieces_kernel(uint8_t *out, int n, int count, uint64_t seed, uint64_t env_base, const uint32_t *__restrict__ episode,
                  uint32_t episode0) {
    const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    uint32_t q[4];
    gen_queue(seed, env_base + (uint64_t)i, episode ? episode[i] : episode0, count, q);
    for (int p = 0; p < count; ++p) out[i * count + p] = (uint8_t)queue_piece(q, p);
}

// =================================================================================================
// fused rollouts: state stays in registers for `steps` moves
// =================================================================================================
template <bool GREEDY>
__global__ void __launch_bounds__(THREADS)
rollout_kernel(uint4 *st, int64_t stride, int n, const uint4 *__restrict__ pool, int K, uint32_t *episode,
               uint32_t *tstep, unsigned long long *stats, int steps, uint64_t seed, uint64_t env_base,
               int gen_count, int L, int M, GreedyWeights gw) {
    __shared__ uint4 s_tab[TAB_WORDS4];
    load_table(s_tab);
    const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (i < n) {
        const uint64_t env = env_base + (uint64_t)i;
        Env e; load_env(st, stride, i, e);
        uint32_t ep = episode[i], t = tstep[i];
        for (int s = 0; s < steps; ++s) {
            if (GREEDY) rollout_greedy_step(e, ep, t, acc, s_tab, pool, K, seed, env, gen_count, L, M, gw);
            else rollout_random_step(e, ep, t, acc, s_tab, pool, K, seed, env, gen_count, L, M);
        }
        store_env(st, stride, i, e);
        episode[i] = ep; tstep[i] = t;
    }
    flush_stats(acc, stats);
}

}  // namespace tpl

// =================================================================================================
// C ABI (device-pointer half)
// =================================================================================================
using namespace tpl;

static inline unsigned grid_for(int n) { return (unsigned)((n + THREADS - 1) / THREADS); }

// Grid-stride kernels: at most `blocks_per_sm` resident CTAs per SM, i.e. a multiple of the SM count (148 on
// B200), so per-CTA set-up and the statistics flush are paid once per CTA slot instead of once per 128 envs.
static unsigned grid_persistent(int n, int blocks_per_sm) {
    static int sms = 0;
    if (!sms) {
        int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    if (blocks_per_sm <= 0) blocks_per_sm = 16;
    const unsigned need = grid_for(n), cap = (unsigned)(sms * blocks_per_sm);
    return need < cap ? need : cap;
}

extern "C" {

int tpl_abi_version(void) { return TPL_ABI_VERSION; }
const char *tpl_last_error(void) { return tpl::last_error(); }
long long tpl_launch_count(void) { return tpl::launches(); }

int tpl_pack(void *out, int64_t plane_stride, int aos, int n, const uint16_t *rows, const uint8_t *pieces,
             int pieces_stride, const uint8_t *npieces, const int32_t *lines, const int32_t *moves, const int8_t *st,
             const uint8_t *head, void *stream) {
    if (n < 0 || !out || !rows || !pieces || !npieces) return fail(TPL_EINVAL, "tpl_pack: null argument");
    if (!aos && plane_stride < n) return fail(TPL_ERANGE, "tpl_pack: plane_stride < n");
    if (n == 0) return 0;
    pack_kernel<<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>((uint4 *)out, plane_stride, aos, n, rows, pieces,
                                                                    pieces_stride, npieces, lines, moves, st, head);
    return check_launch("tpl_pack");
}

int tpl_unpack(const void *state, int64_t plane_stride, int n, uint16_t *rows, uint8_t *cur, uint8_t *next, int32_t *lines,
               int32_t *moves, int8_t *st, uint8_t *head, uint8_t *npieces, uint8_t *queue, void *stream) {
    if (n < 0 || !state) return fail(TPL_EINVAL, "tpl_unpack: null state");
    if (plane_stride < n) return fail(TPL_ERANGE, "tpl_unpack: plane_stride < n");
    if (n == 0) return 0;
    unpack_kernel<<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>((const uint4 *)state, plane_stride, n, rows, cur, next,
                                                                      lines, moves, st, head, npieces, queue);
    return check_launch("tpl_unpack");
}

int tpl_reset_from_pool(void *state, int64_t plane_stride, int n, const void *pool, int K, const int32_t *idx,
                        const uint8_t *mask, int mode, uint32_t *episode, uint64_t seed, uint64_t env_base, int gen_count,
                        void *stream) {
    if (n < 0 || !state || !pool || K <= 0) return fail(TPL_EINVAL, "tpl_reset_from_pool: null state/pool or K <= 0");
    if (plane_stride < n) return fail(TPL_ERANGE, "tpl_reset_from_pool: plane_stride < n");
    if (mode < TPL_RESET_ALL || mode > TPL_RESET_DONE) return fail(TPL_EINVAL, "tpl_reset_from_pool: bad mode %d", mode);
    if (mode == TPL_RESET_MASK && !mask) return fail(TPL_EINVAL, "tpl_reset_from_pool: mode MASK needs a mask");
    if (gen_count < 0 || gen_count > TPL_MAX_PIECES) return fail(TPL_ERANGE, "tpl_reset_from_pool: gen_count %d > 42", gen_count);
    if (n == 0) return 0;
    reset_kernel<<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>((uint4 *)state, plane_stride, n, (const uint4 *)pool, K, idx,
                                                                     mask, mode, episode, seed, env_base, gen_count);
    return check_launch("tpl_reset_from_pool");
}

int tpl_step(void *state, int64_t plane_stride, int n, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags,
             int8_t *st, long long *stats, int L, int M, void *stream) {
    if (n < 0 || !state || !rot || !loc) return fail(TPL_EINVAL, "tpl_step: null argument");
    if (plane_stride < n) return fail(TPL_ERANGE, "tpl_step: plane_stride < n");
    if (L < 0 || M < 0 || M > 65535 || L > 65535) return fail(TPL_ERANGE, "tpl_step: L/M out of range");
    if (n == 0) return 0;
    step_kernel<<<grid_persistent(n, 16), THREADS, 0, (cudaStream_t)stream>>>((uint4 *)state, plane_stride, n, rot, loc, dlines, flags, st,
                                                                    (unsigned long long *)stats, L, M);
    return check_launch("tpl_step");
}

int tpl_afterstates(const void *state, int64_t plane_stride, int n, uint8_t *feats, uint8_t *flags, float *feats_f32, int L, int M,
                    void *stream) {
    if (n < 0 || !state) return fail(TPL_EINVAL, "tpl_afterstates: null state");
    if (plane_stride < n) return fail(TPL_ERANGE, "tpl_afterstates: plane_stride < n");
    if (n > (1 << 25)) return fail(TPL_ERANGE, "tpl_afterstates: at most 2^25 envs per call (32-bit output offsets)");
    if (!feats && !feats_f32) return fail(TPL_EINVAL, "tpl_afterstates: no feature output requested");
    if (feats_f32 && !flags) return fail(TPL_EINVAL, "tpl_afterstates: the float form needs the flags array");
    if (n == 0) return 0;
    const cudaStream_t s = (cudaStream_t)stream;
    const uint4 *st = (const uint4 *)state; uint32_t *w = (uint32_t *)feats; float4 *f = (float4 *)feats_f32;
    const unsigned g = grid_persistent(n, 8);
    if (feats && !flags) afterstates_kernel<0><<<g, THREADS, 0, s>>>(st, plane_stride, n, w, flags, f, L, M);
    else if (feats && !feats_f32) afterstates_kernel<1><<<g, THREADS, 0, s>>>(st, plane_stride, n, w, flags, f, L, M);
    else if (!feats) afterstates_kernel<2><<<g, THREADS, 0, s>>>(st, plane_stride, n, w, flags, f, L, M);
    else afterstates_kernel<3><<<g, THREADS, 0, s>>>(st, plane_stride, n, w, flags, f, L, M);
    return check_launch("tpl_afterstates");
}

int tpl_gen_pieces(uint8_t *out, int n, int count, uint64_t seed, uint64_t env_base, const uint32_t *episode, uint32_t episode0,
                   void *stream) {
    if (n < 0 || !out) return fail(TPL_EINVAL, "tpl_gen_pieces: null output");
    if (count < 0 || count > TPL_MAX_PIECES) return fail(TPL_ERANGE, "tpl_gen_pieces: count %d outside 0..42", count);
    if (n == 0 || count == 0) return 0;
    gen_pieces_kernel<<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>(out, n, count, seed, env_base, episode, episode0);
    return check_launch("tpl_gen_pieces");
}

static int rollout_common(const char *who, bool greedy, void *state, int64_t plane_stride, int n, const void *pool, int K,
                          uint32_t *episode, uint32_t *tstep, long long *stats, int steps, const int32_t *w6, uint64_t seed,
                          uint64_t env_base, int gen_count, int L, int M, void *stream) {
    if (n < 0 || !state || !pool || K <= 0 || !episode || !tstep || !stats) return fail(TPL_EINVAL, "%s: null argument or K <= 0", who);
    if (plane_stride < n) return fail(TPL_ERANGE, "%s: plane_stride < n", who);
    if (gen_count < 0 || gen_count > TPL_MAX_PIECES) return fail(TPL_ERANGE, "%s: gen_count %d > 42", who, gen_count);
    if (greedy && !w6) return fail(TPL_EINVAL, "%s: null weights", who);
    if (n == 0 || steps <= 0) return 0;
    GreedyWeights gw{};
    if (greedy) for (int q = 0; q < 6; ++q) gw.w[q] = w6[q];
    if (greedy)
        rollout_kernel<true><<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>((uint4 *)state, plane_stride, n, (const uint4 *)pool, K, episode,
                                                                               tstep, (unsigned long long *)stats, steps, seed, env_base,
                                                                               gen_count, L, M, gw);
    else
        rollout_kernel<false><<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>((uint4 *)state, plane_stride, n, (const uint4 *)pool, K, episode,
                                                                                tstep, (unsigned long long *)stats, steps, seed, env_base,
                                                                                gen_count, L, M, gw);
    return check_launch(who);
}

int tpl_rollout_random(void *state, int64_t plane_stride, int n, const void *pool, int K, uint32_t *episode, uint32_t *tstep,
                       long long *stats, int steps, uint64_t seed, uint64_t env_base, int gen_count, int L, int M, void *stream) {
    return rollout_common("tpl_rollout_random", false, state, plane_stride, n, pool, K, episode, tstep, stats, steps, nullptr, seed,
                          env_base, gen_count, L, M, stream);
}

int tpl_rollout_greedy(void *state, int64_t plane_stride, int n, const void *pool, int K, uint32_t *episode, uint32_t *tstep,
                       long long *stats, int steps, const int32_t *weights6_host, uint64_t seed, uint64_t env_base, int gen_count,
                       int L, int M, void *stream) {
    return rollout_common("tpl_rollout_greedy", true, state, plane_stride, n, pool, K, episode, tstep, stats, steps, weights6_host, seed,
                          env_base, gen_count, L, M, stream);
}

}  // extern "C"
