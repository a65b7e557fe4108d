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
#include <cstdlib>

namespace tpl {

#ifndef TPL_THREADS
#define TPL_THREADS 128              // threads per CTA of every kernel here (tuning knob; the launch bounds below assume 128)
#endif
constexpr int THREADS = TPL_THREADS;

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

#define TPL_SCRATCH __shared__ uint32_t s_scr[SCR_ROWS * THREADS]; uint32_t *scr = s_scr + threadIdx.x

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------------
// The persistent kernels are launched with the programmatic-stream-serialization attribute (launch_pdl below).  Each CTA lets
// the NEXT kernel of the stream start launching right away (pdl_trigger) -- its CTAs become resident as the CTAs of this grid
// retire, copy the orientation table and set up their shared memory during this grid's tail -- and touches nothing a previous
// kernel may still be producing before pdl_wait(), which returns once every grid this one depends on has completed and its
// memory is visible.  Without the launch attribute both instructions do nothing.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void load_table(uint4 *s_tab) {
    for (int t = threadIdx.x; t < TAB_WORDS4; t += blockDim.x) s_tab[t] = reinterpret_cast<const uint4 *>(&c_orient)[t];
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
    const int np = min(min((int)npieces[i], TPL_MAX_PIECES), pstride);
    pack_queue(pieces + i * pstride, np, e.q);
    e.lines = lines ? (uint32_t)lines[i] : 0u;
    e.moves = moves ? (uint32_t)moves[i] : 0u;
    e.state = st ? (uint32_t)st[i] : 0u;
    e.head = head ? head[i] : 0u;
    e.npieces = (uint32_t)np;
    e.qblock = 0u;
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
             const uint8_t *__restrict__ mask, int mode, uint32_t *episode, uint32_t *tstep, uint64_t seed, uint64_t env_base, int gen_count) {
    const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    if (mode == TPL_RESET_MASK && !mask[i]) return;
    uint32_t ep = episode ? episode[i] : 0u;
    if (mode == TPL_RESET_DONE) {
        const uint4 d = st[3 * stride + i];
        const uint32_t state = d.w & 0xFFu, head = (d.w >> 8) & 0xFFu, np = (d.w >> 16) & 0xFFu;
        if (state == S_RUNNING && head < np) return;
        if (gen_count > QUEUE_PIECES && state == S_RUNNING) {        // a running env on an empty queue: next block of its sequence
            Env e; load_env(st, stride, i, e);
            if (refill_queue(e, seed, env_base + (uint64_t)i, ep, gen_count)) {
                st[2 * stride + i] = make_uint4(e.col[8], e.col[9], e.q[0], e.q[1]);
                st[3 * stride + i] = pack_meta(e);
                return;
            }
        }
    }
    // a masked or auto reset WITHOUT explicit indices starts a new episode: the counter is bumped before the draw, so the env
    // does not get the config (and the action stream) of the episode it just finished again
    if (mode == TPL_RESET_DONE || (mode == TPL_RESET_MASK && !idx)) {
        ep += 1;
        if (episode) episode[i] = ep;
    }
    if (tstep) tstep[i] = 0u;                           // every reset path restarts the rollouts' per-episode action stream
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
    TPL_SCRATCH;
    pdl_trigger();
    load_table(s_tab);
    pdl_wait();
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // software pipeline: the next env's record and action are in flight while the current move is computed
    const int64_t stepn = (int64_t)gridDim.x * THREADS;
    int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a, d = a; uint32_t ar = 0, al = 0;
    if (i < n) { a = st[i]; b = st[stride + i]; c = st[2 * stride + i]; d = st[3 * stride + i]; ar = rot[i]; al = loc[i]; }
    for (; i < n; i += stepn) {
        const int64_t i2 = i + stepn;
        uint4 na = a, nb = b, nc = c, nd = d; uint32_t nr = 0, nl = 0;
        if (i2 < n) { na = st[i2]; nb = st[stride + i2]; nc = st[2 * stride + i2]; nd = st[3 * stride + i2]; nr = rot[i2]; nl = loc[i2]; }
        Env e; unpack_env(a, b, c, d, e);
        const uint32_t was = e.state;
        int k; bool changed;
        const uint32_t fl = step_env(e, s_tab, scr, THREADS, ar, al, L, M, k, changed);
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
        a = na; b = nb; c = nc; d = nd; ar = nr; al = nl;
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
// P32: the whole output array lies inside one 4 GB-aligned window (checked on the host), so advancing the store pointer by
// one slot row is a 32-bit add on its low word -- one instruction instead of the IADD3 + IMAD.X pair of a 64-bit add.
template <int MODE, bool P32 = false>
struct GlobalSink {
    static constexpr bool PACKED = (MODE == 0), RAGGED = false;
    uint32_t *words; uint8_t *flags; float4 *ff;      // already offset by the env index
    uint32_t n;
    uint32_t one;                                     // the value 1 from a kernel parameter, i.e. opaque to ptxas (see next_col)
    uint32_t *pcol;                                   // compact form: this env's word of the current (rotation, column)
    __device__ __forceinline__ void begin_rotation(int r) {
#ifdef __CUDA_ARCH__
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(pcol) : "r"((uint32_t)(r * 10) * 4u), "r"(n), "l"(words));
#endif
    }
    __device__ __forceinline__ void put_packed_col(int, uint32_t packed) { *pcol = packed; }
    // The same address through an opaque copy of the pointer.  `if (!a) *p = x; if (b) *p = y;` on one visible address is fused
    // into a select, a combined predicate and ONE store -- two instructions on the ALU pipe, the pipe that bounds the
    // enumeration -- while two predicated stores cost one more issue slot on the load/store pipe, which has room.
    __device__ __forceinline__ void put_packed_col_again(int, uint32_t packed) {
        uint32_t *p2 = pcol;
#ifdef __CUDA_ARCH__
        asm("" : "+l"(p2));
#endif
        *p2 = packed;
    }
    // pcol += n words.  64-bit form: ONE multiply-add by `one` (with an immediate multiplier ptxas strength-reduces it to a
    // LEA / LEA.HI.X pair on the ALU pipe); ptxas turns it into IADD3 + IMAD.X.
    __device__ __forceinline__ void next_col() {
#ifdef __CUDA_ARCH__
        if constexpr (P32) {
            uint32_t lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(pcol));
            lo = mad_fma_pipe(n * 4u, one, lo);              // (pinned to the FMA pipe: left to itself ptxas sometimes picks an ALU-pipe add)
            asm("mov.b64 %0, {%1, %2};" : "=l"(pcol) : "r"(lo), "r"(hi));
        } else {
            asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(pcol) : "r"(n * 4u), "r"(one));
        }
#endif
    }
    __device__ __forceinline__ void put_packed(int slot, uint32_t packed) { words[(uint32_t)slot * n] = packed; }
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

// ---- distinct-placements ("ragged") output form (MODE 4) -------------------------------------------------------------------
// Only the placements that differ are written: rotations r < n_rot (game/tetris.py:61) and columns c <= 10 - w (:364), in
// rotation-major order -- 9 (O), 17 (I, S, Z) or 34 (L, J, T) words per env, 23.1 on average instead of 40.  The runs of a
// 32-env tile are packed back to back: each lane writes its words into the warp's staging area in shared memory at the
// exclusive prefix of the run lengths (a 32-bit shared-memory address + an immediate per column: no pointer arithmetic per
// slot), the warp reserves `total` words of the output array with ONE atomicAdd on the cursor (issued before the
// enumeration, consumed after it), and the tile leaves the SM as contiguous 16-byte stores.  Where an env's run starts is
// reported per env (`runs[i] = word offset | piece << 29`), so the order of the tiles in the array does not matter.
constexpr int RAG_STAGE_WORDS = DISTINCT_MAX * 32 + 16;      // + a 10-word dummy row for the lanes whose rotation is an alias
constexpr int RAG_DUMMY_WORD = DISTINCT_MAX * 32;
constexpr int RAG_SMEM_BYTES = RAG_STAGE_WORDS * 4 * (THREADS / 32);

struct RaggedStageSink {
    static constexpr bool PACKED = true, RAGGED = true;
    uint32_t run_saddr;        // shared-space byte address of this env's run in the staging area
    uint32_t dummy_saddr;      // ... of the warp's dummy row
    uint32_t prow;             // ... of placement (current rotation, column 0)
    __device__ __forceinline__ void begin_rotation_ragged(bool canon, uint32_t rot_base) {
        prow = canon ? run_saddr + rot_base * 4u : dummy_saddr;
    }
    template <int C> __device__ __forceinline__ void put_col(uint32_t w) {
#ifdef __CUDA_ARCH__
        asm volatile("st.shared.b32 [%0+%1], %2;" :: "r"(prow), "n"(C * 4), "r"(w) : "memory");
#endif
    }
    __device__ __forceinline__ void put_canon(int idx, uint32_t w) {           // (a deferred slot resolved in place: unused by the kernels)
#ifdef __CUDA_ARCH__
        asm volatile("st.shared.b32 [%0], %1;" :: "r"(run_saddr + (uint32_t)idx * 4u), "r"(w) : "memory");
#endif
    }
};
struct RaggedGlobalSink {                                    // deferred slots go straight to the env's run in global memory
    static constexpr bool PACKED = true, RAGGED = true;
    uint32_t *run;
    __device__ __forceinline__ void put_canon(int idx, uint32_t w) { run[idx] = w; }
};
template <int MODE> struct ResolveSink { using type = GlobalSink<MODE>; };
template <> struct ResolveSink<4> { using type = RaggedGlobalSink; };
// `at` = the env index (40-slot forms) or the word offset of the env's run (distinct-placements form)
template <int MODE>
__device__ __forceinline__ typename ResolveSink<MODE>::type make_resolve_sink(uint32_t at, uint32_t n, uint32_t *words, uint8_t *flags,
                                                                              float4 *ff, uint32_t one) {
    if constexpr (MODE == 4) return RaggedGlobalSink{words + at};
    else return GlobalSink<MODE>{words + at, flags + at, ff + at, n, one, nullptr};
}

// ---- warp-private queues of deferred (row-completing) slots ---------------------------------------------------------
// Only ~2 lanes of a warp have such a slot in any given tile, so resolving them in place runs the ~300-instruction general
// move at a few % lane utilisation, and pooling them per CTA (the previous design) costs two block-wide barriers per tile:
// ncu attributed 27 % of all warp samples to barrier stalls, because the short resolve phase keeps one warp busy while
// three wait.  Here every warp owns a small FIFO in shared memory: lanes append their deferred slots (and park the ten
// columns of their env once), the warp keeps walking through its tiles, and whenever 32 slots have accumulated they are
// resolved by a full warp.  No __syncthreads in the tile loop at all: the 16 resident warps of an SM drift freely and
// cover each other's load latency.
constexpr int WQ_ENVS = 64;          // env entries: < 32 live after a resolve + <= 32 new per tile
#ifndef TPL_WQ_ITEMS
#define TPL_WQ_ITEMS 128             // (-DTPL_WQ_ITEMS=72 makes the wait-for-room path of wq_publish run often: scripts/overflow_check.sh;
                                     //  128 instead of 512 keeps the distinct-form kernels at four CTAs per SM next to the pool-record stage)
#endif
constexpr int WQ_ITEMS = TPL_WQ_ITEMS;   // deferred slots; a lane whose slots do not fit waits for the next round of wq_publish (never seen in practice)
static_assert(WQ_ITEMS >= 72, "after a resolve (< 32 slots live) the 40 slots of any one lane must fit");
struct WarpQueue {
    uint32_t env[13 * WQ_ENVS];      // [k][entry]: 10 columns, piece | cells << 8 | fl_noclear << 16, lines, env index
    uint16_t items[WQ_ITEMS];        // entry | slot << 6
};
struct WqPos { uint32_t head, tail, ehead, etail; };      // warp-uniform, monotonically increasing (indices are taken modulo)

__device__ __forceinline__ void rag_wait(bool complete);

// Exclusive prefix sum over the warp of small counts (cnt <= 63; top = their warp maximum): one ballot per bit that is set in
// any count -- independent instructions -- instead of the five dependent shuffle + select + add rounds of the classic scan
// (ncu: the scan alone held 2 % of all warp time in short-scoreboard stalls; the counts here are almost always 0, 1 or 2).
__device__ __forceinline__ uint32_t warp_prefix_small(uint32_t cnt, uint32_t top) {
    const uint32_t lt = (1u << (threadIdx.x & 31u)) - 1u;
    uint32_t excl = 0u;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        if ((top >> b) == 0u) break;                                   // warp-uniform
        excl += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, (cnt >> b) & 1u) & lt) << b;
    }
    return excl;
}

// Every lane of the warp calls this once per tile (cx.mask == 0: nothing to defer); `i` = env index, or in the
// distinct-placements form the word offset of the env's run.  Publish, then resolve while 32 slots are queued.  `flush`: resolve
// whatever is queued (the kernels run one extra, tile-less round of their loop for that, so that the ~350 instructions of the
// resolver exist ONCE in the kernel instead of three times -- they used to be a fifth of its code).  A lane whose slots do not
// fit the ring waits for the next round of the loop below: after a resolve fewer than 32 slots are live, so at least
// WQ_ITEMS - 31 >= 40 fit, i.e. any lane's.
template <int MODE>
__device__ __forceinline__ void wq_publish(WarpQueue &q, WqPos &p, const Env &e, const PendingCtx &cx, uint32_t i, bool flush,
                                           const uint4 *s_tab, uint32_t *scr, uint32_t n, uint32_t *words, uint8_t *flags,
                                           float4 *ff, int L, uint32_t one) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t cnt = (uint32_t)__popcll(cx.mask);                        // this lane's slots not yet in the queue
    for (;;) {
        const uint32_t top = __reduce_max_sync(0xFFFFFFFFu, cnt);
        if (top != 0u) {
            const uint32_t incl = warp_prefix_small(cnt, top) + cnt;
            const uint32_t room = WQ_ITEMS - (p.tail - p.head);
            const bool fits = cnt != 0u && incl <= room;               // monotone in the lane index
            const unsigned fit = __ballot_sync(0xFFFFFFFFu, fits);
            if (fits) {
                const uint32_t entry = (p.etail + (uint32_t)__popc(fit & ((1u << lane) - 1u))) % WQ_ENVS;
#pragma unroll
                for (int k = 0; k < COLS; ++k) q.env[k * WQ_ENVS + entry] = e.col[k];
                q.env[10 * WQ_ENVS + entry] = cx.piece | (cx.cells << 8) | (cx.fl_noclear << 16);
                q.env[11 * WQ_ENVS + entry] = cx.lines;
                q.env[12 * WQ_ENVS + entry] = i;
                unsigned long long m = cx.mask;
                uint32_t at = p.tail + incl - cnt;
                while (m) {
                    const int sl = __ffsll((long long)m) - 1;
                    m &= m - 1ull;
                    q.items[at++ % WQ_ITEMS] = (uint16_t)(entry | ((uint32_t)sl << 6));
                }
                cnt = 0u;
            }
            if (fit) {
                p.tail += __shfl_sync(0xFFFFFFFFu, incl, 31 - __clz(fit));      // inclusive count at the last lane that fits
                p.etail += (uint32_t)__popc(fit);
            }
            __syncwarp();
        }
        const bool more = top != 0u && __any_sync(0xFFFFFFFFu, cnt != 0u);      // (some lane is still waiting for room)
        while (p.tail - p.head >= 32u || ((flush || more) && p.tail != p.head)) {
            if constexpr (MODE == 4) rag_wait(true);
            const uint32_t live = p.tail - p.head, take = live < 32u ? live : 32u;
            if (lane < take) {
                const uint32_t item = q.items[(p.head + lane) % WQ_ITEMS], entry = item & 63u, slot = item >> 6;
                uint32_t cols[COLS];
#pragma unroll
                for (int j = 0; j < COLS; ++j) cols[j] = q.env[j * WQ_ENVS + entry];
                const uint32_t m0 = q.env[10 * WQ_ENVS + entry];
                const PendingCtx c2{0ull, m0 & 0xFFu, (m0 >> 8) & 0xFFu, q.env[11 * WQ_ENVS + entry], m0 >> 16};
                auto sink = make_resolve_sink<MODE>(q.env[12 * WQ_ENVS + entry], n, words, flags, ff, one);
                resolve_slot(cols, c2, (int)slot, s_tab, scr, THREADS, L, sink);
            }
            __syncwarp();
            p.head += take;
            if (p.head != p.tail) p.ehead += ((q.items[p.head % WQ_ITEMS] & 63u) - p.ehead) & 63u;      // entry of the first slot left
            else p.ehead = p.etail;
        }
        if (!more) break;
    }
}

// ---- distinct-placements form: the two warp-collective halves around the enumeration of a tile
struct RagTile { uint32_t excl, total, tb; };
// cnt = this lane's run length (0 / 9 / 17 / 34): exclusive prefix over the warp, and one atomicAdd reserving the tile's
// words (rounded up to 4 so that every tile starts 16-byte aligned); the returned base is only needed after the enumeration
__device__ __forceinline__ RagTile rag_reserve(uint32_t cnt, uint32_t *cursor, uint32_t one) {
    const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    // the run lengths are 0, 9, 17 or 34: the prefix is three ballots (independent) instead of a five-round shuffle scan
    const unsigned b9 = __ballot_sync(0xFFFFFFFFu, cnt == 9u), b17 = __ballot_sync(0xFFFFFFFFu, cnt == 17u), b34 = __ballot_sync(0xFFFFFFFFu, cnt == 34u);
    RagTile t;
    t.excl = 9u * (uint32_t)__popc(b9 & lt) + 17u * (uint32_t)__popc(b17 & lt) + 34u * (uint32_t)__popc(b34 & lt);
    t.total = 9u * (uint32_t)__popc(b9) + 17u * (uint32_t)__popc(b17) + 34u * (uint32_t)__popc(b34);
    t.tb = 0u;
    // The address is made to LOOK thread-dependent ((one - 1) * threadIdx.x = 0; `one` is the opaque kernel parameter 1): on a provably
    // warp-uniform address ptxas wraps the atomic -- also an inline-PTX one -- in its warp-aggregation pattern, whose result
    // broadcast (a shuffle right behind the atomic) made every tile wait for the contended counter HERE, 7 % of the
    // distinct-form step's warp time, instead of after the enumeration, where the base is first needed.
    if (lane == 0u && t.total) t.tb = atomicAdd(cursor + (one - 1u) * threadIdx.x, (t.total + 3u) & ~3u);
    return t;
}
// staging area -> rows[tile base ...] with 16-byte stores; returns the tile base (word offset).  The trailing __syncwarp orders
// these stores before the deferred-slot resolver's (other lanes of this warp write single words of the same runs later)
// and lets the next tile reuse the staging area.
#ifndef TPL_RAG_TMA
#define TPL_RAG_TMA 1               // the staged tile leaves with ONE bulk copy (cp.async.bulk, TMA engine) instead of a store loop:
#endif                              // 0.1048 -> 0.1027 ms per 2^20-env step (-DTPL_RAG_TMA=0 keeps the plain loop for comparison)
__device__ __forceinline__ uint32_t rag_flush(uint32_t *stage, const RagTile &t, uint32_t *rows) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tb = __shfl_sync(0xFFFFFFFFu, t.tb, 0);
    const uint32_t ta = (t.total + 3u) & ~3u;
    if (lane < ta - t.total) stage[t.total + lane] = 0u;                     // the alignment gap holds zeros, not stale words
#if TPL_RAG_TMA
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");             // this lane's staging writes -> visible to the bulk engine
    __syncwarp();
    if (lane == 0u && ta) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(stage);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(rows + tb), "r"(sa), "r"(ta * 4u) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
#else
    __syncwarp();
    for (uint32_t k = lane * 4u; k < ta; k += 128u)
        *reinterpret_cast<uint4 *>(rows + tb + k) = *reinterpret_cast<const uint4 *>(stage + k);
    __syncwarp();
#endif
    return tb;
}
// bulk-copy variant only: the staging area may be rewritten once the engine has READ it; the deferred-slot resolver may
// write single words of a flushed run once the copy is COMPLETE (both are waits of the issuing lane, then a warp sync)
__device__ __forceinline__ void rag_wait(bool complete) {
#if TPL_RAG_TMA
    if ((threadIdx.x & 31u) == 0u) {
        if (complete) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
#endif
}

__device__ __forceinline__ void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" :: "l"(ptr)); }
__device__ __forceinline__ void prefetch_l1(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" :: "l"(ptr)); }

// ---- record pipeline: the four 16-byte chunks of the NEXT tile's records travel global -> shared memory with cp.async
// (LDGSTS: no registers held while in flight) during the ~2 300 instructions the warp spends on the current tile, so a
// tile starts with four LDS.128 instead of four exposed DRAM round trips (ncu: long-scoreboard was the top stall reason).
struct RecordStage { uint4 chunk[4][32]; };            // one per warp
__device__ __forceinline__ void stage_issue(RecordStage &rs, const uint4 *st, int64_t stride, int64_t i, int n) {
    const uint32_t lane = threadIdx.x & 31u;
    const int64_t ic = i < n ? i : (int64_t)n - 1;      // out-of-range lanes copy a valid record they never use
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&rs.chunk[j][lane]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(st + j * stride + ic) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_take(RecordStage &rs, Env &e) {
    const uint32_t lane = threadIdx.x & 31u;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const uint4 a = rs.chunk[0][lane], b = rs.chunk[1][lane], c = rs.chunk[2][lane], d = rs.chunk[3][lane];   // own copies only
    unpack_env(a, b, c, d, e);
}

// ---- pool-record pipeline: the record a reset of the NEXT tile would install -- a pure function of (seed, env, episode + 1) --
// is copied global -> shared memory with cp.async while the current tile is enumerated, like the env records themselves.
// Why: the gather is L2-resident, yet ncu showed ~1000 cycles of long-scoreboard stall per tile where the record was first used
// (6.8 % of all warp time) although it had been prefetched towards L1 at the top of the tile; loading it into registers ahead
// of the move shortened the stall but not the step (the loads queue behind the tile's DRAM traffic in the in-order L1, and a
// dead word of the record made ptxas reuse its register, which then waited for the load: write-after-write).  A whole
// enumeration (~8000 cycles) of distance costs no registers and removes the stall.
__device__ __forceinline__ void pool_issue(RecordStage &ps, const uint4 *pool, uint32_t k) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint4 *p = pool + 4 * (size_t)k;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&ps.chunk[j][lane]);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(p + j) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

#ifndef TPL_AS_MINBLOCKS
#define TPL_AS_MINBLOCKS 4          // 128 registers per thread: fewer re-materialised operands than the default choice of 96
#endif
#ifdef TPL_AS_MAXNREG               // tuning knob: an explicit register cap instead of the one the launch bounds imply
#define TPL_AS_BOUNDS __maxnreg__(TPL_AS_MAXNREG)
#else
#define TPL_AS_BOUNDS __launch_bounds__(THREADS, TPL_AS_MINBLOCKS)
#endif
// piece and run length (distinct placements) of the env's current piece; 7 / 0 when the queue is empty
__device__ __forceinline__ void current_run(const Env &e, const uint4 *tab, uint32_t &piece, uint32_t &cnt) {
    piece = 7u; cnt = 0u;
    if (e.head < e.npieces) { piece = queue_piece(e.q, e.head); cnt = orient_run_len(tab[piece * 8 + 1]); }
}

template <int MODE, bool P32 = false>
__global__ void TPL_AS_BOUNDS
afterstates_kernel(const uint4 *__restrict__ st, int64_t stride, int n, uint32_t *__restrict__ words,
                   uint8_t *__restrict__ flags, float4 *__restrict__ ff, int L, int M, uint32_t one,
                   uint32_t *__restrict__ runs, uint32_t run_base, uint32_t *cursor, uint32_t *cursor_clear) {
    constexpr bool RAG = (MODE == 4);
    extern __shared__ __align__(16) uint32_t s_rag[];                 // RAG: one staging area per warp (RAG_SMEM_BYTES)
    __shared__ uint4 s_tab[TAB_WORDS4];
    __shared__ WarpQueue s_wq[THREADS / 32];
    TPL_SCRATCH;
    pdl_trigger();
    load_table(s_tab);
    __shared__ RecordStage s_stage[THREADS / 32];
    WarpQueue &q = s_wq[threadIdx.x >> 5];
    RecordStage &rs = s_stage[threadIdx.x >> 5];
    uint32_t *stage = s_rag + (threadIdx.x >> 5) * RAG_STAGE_WORDS;
    pdl_wait();                                                      // nothing above reads or writes what another kernel produces
    if (RAG && cursor_clear && blockIdx.x == 0 && threadIdx.x == 0) *cursor_clear = 0u;      // the NEXT call's counter
    WqPos qp{0u, 0u, 0u, 0u};
    const int wtiles = (n + 31) / 32, wstep = (int)gridDim.x * (THREADS / 32);
    int t = (int)blockIdx.x * (THREADS / 32) + (int)(threadIdx.x >> 5);
    if (t < wtiles) stage_issue(rs, st, stride, (int64_t)t * 32 + (threadIdx.x & 31), n);
    // (warp-uniform trip count; one extra round after the warp's last tile drains its queue of deferred slots: see wq_publish)
    for (;; t += wstep) {
        const bool drain = t >= wtiles;
        const int64_t i = (int64_t)t * 32 + (threadIdx.x & 31);
        PendingCtx cx; cx.mask = 0ull;
        Env e;
        uint32_t at = 0u;                                // what the resolver addresses the env's slots by
        if (!drain) {
            stage_take(rs, e);                           // each lane reads back only what it copied itself: no warp sync needed
            if (t + wstep < wtiles) stage_issue(rs, st, stride, i + (int64_t)wstep * 32, n);
            if constexpr (RAG) {
                uint32_t piece = 7u, cnt = 0u;
                if (i < n) current_run(e, s_tab, piece, cnt);
                const RagTile rt = rag_reserve(cnt, cursor, one);
                rag_wait(false);
                if (i < n) {
                    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage);
                    RaggedStageSink sink{sbase + rt.excl * 4u, sbase + RAG_DUMMY_WORD * 4u, 0u};
                    afterstates_env(e, s_tab, scr, THREADS, L, M, sink, 0, 4, &cx, one);
                }
                at = rag_flush(stage, rt, words) + rt.excl;
                if (i < n) runs[i] = (at + run_base) | (piece << 29);         // (the resolver addresses `words + at`: call-local)
            } else {
                if (i < n) {
                    GlobalSink<MODE, P32> sink{words + i, flags + i, ff + i, (uint32_t)n, one, nullptr};
                    afterstates_env(e, s_tab, scr, THREADS, L, M, sink, 0, 4, &cx, one);
                }
                at = (uint32_t)i;
            }
        }
        wq_publish<MODE>(q, qp, e, cx, at, drain, s_tab, scr, (uint32_t)n, words, flags, ff, L, one);
        if (drain) break;
    }
}

// Small batches (BASELINE configs[1]: 4096 envs = 32 CTAs on 148 SMs) are latency-bound: one warp per scheduler runs a
// 2 000-instruction dependent stream.  This variant gives each (env, rotation) pair its own thread: 4x the threads, a
// 4x shorter critical path, at the price of repeating the per-env set-up in the four threads of an env.
template <int MODE>
__global__ void __launch_bounds__(THREADS)
afterstates_split_kernel(const uint4 *__restrict__ st, int64_t stride, int n, uint32_t *__restrict__ words,
                         uint8_t *__restrict__ flags, float4 *__restrict__ ff, int L, int M, uint32_t one) {
    __shared__ uint4 s_tab[TAB_WORDS4];
    TPL_SCRATCH;
    load_table(s_tab);
    const int64_t idx = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    const int64_t i = idx >> 2;
    const int r = (int)(idx & 3);
    if (i >= n) return;
    Env e; load_env(st, stride, i, e);
    GlobalSink<MODE> sink{words + i, flags + i, ff + i, (uint32_t)n, one, nullptr};
    afterstates_env(e, s_tab, scr, THREADS, L, M, sink, r, r + 1);
}

// =================================================================================================
// afterstates, piece-sorted tiles (compact output form).
//
// The plain kernel above enumerates all 40 slots of every env because the lanes of a warp hold different pieces:
// rotations >= n_rot and columns > 10 - w are aliases for one lane but real placements for its neighbour.  Here a
// CTA of 128 threads takes a tile of 256 consecutive envs, counting-sorts their ids by current piece in shared memory
// (match.any + one shared atomic per warp and piece) and enumerates them in sorted order, two per thread: warps
// become (almost) piece-uniform, alias rotations/columns are skipped warp-wide (23 of 40 slots are distinct on
// average) and filled by copies.  Sorted positions are dealt out serpentine (t and 255 - t), so every warp gets a light
// and a heavy half and the warps of a CTA finish together.  Results go into a [40][256] shared-memory tile at the env's
// ORIGINAL position, so the tile leaves the SM as 40 contiguous 1 KB rows through the TMA bulk-copy engine
// (cp.async.bulk, shared -> global): no per-thread store addressing at all.
// =================================================================================================
constexpr int ST = 128;                                   // threads per CTA
constexpr int TILE = 2 * ST;                              // envs per tile
constexpr int SORT_SMEM_BYTES = 40 * TILE * 4 + SCR_ROWS * ST * 4 + TAB_WORDS4 * 16 + TILE * 2 + 64;

struct TileSink {                                          // packed words at the env's original column of the tile
    static constexpr bool PACKED = true, RAGGED = false;
    uint32_t *out;                                         // s_out + original local id
    uint32_t *orot;
    __device__ __forceinline__ void begin_rotation(int r) { orot = out + r * 10 * TILE; }
    __device__ __forceinline__ void put_packed_col(int c, uint32_t packed) { orot[c * TILE] = packed; }
    __device__ __forceinline__ void put_packed_col_again(int c, uint32_t packed) { orot[c * TILE] = packed; }
    __device__ __forceinline__ void next_col() {}
    __device__ __forceinline__ void put_packed(int slot, uint32_t packed) { out[slot * TILE] = packed; }
    __device__ __forceinline__ void put(int slot, uint32_t word, uint32_t fl) { out[slot * TILE] = word | (fl << 3); }
    __device__ __forceinline__ void copy(int dst, int src, uint32_t extra) { out[dst * TILE] = out[src * TILE] | (extra << 3); }
};

__device__ __forceinline__ void bulk_store_row(void *gdst, const void *ssrc, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
}

// sort key of an env from its chunks 2 and 3: pieces with the same rotation count adjacent (O | I S Z | L J T), no piece last
__device__ __forceinline__ uint32_t sort_key(const uint4 &c, const uint4 &d) {
    const uint32_t head = (d.w >> 8) & 0xFFu, np = (d.w >> 16) & 0xFFu;
    if (head >= np) return 7u;
    const uint32_t q[4] = {c.z, c.w, d.x, d.y};
    return (0x0326541u >> (4 * queue_piece(q, head))) & 7u;                          // I L J T S Z O -> 1 4 5 6 2 3 0
}

__global__ void __launch_bounds__(ST, 4)
afterstates_sorted_kernel(const uint4 *__restrict__ st, int64_t stride, int n, uint32_t *__restrict__ words, int L, int M) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *s_out = reinterpret_cast<uint32_t *>(smem_raw);                       // [40][TILE]
    uint32_t *s_scr = s_out + 40 * TILE;                                            // [SCR_ROWS][ST]
    uint4 *s_tab = reinterpret_cast<uint4 *>(s_scr + SCR_ROWS * ST);                // [TAB_WORDS4]
    uint16_t *s_orig = reinterpret_cast<uint16_t *>(s_tab + TAB_WORDS4);            // [TILE] local id of each sorted position
    int *s_cnt = reinterpret_cast<int *>(s_orig + TILE);                            // [8] counts, [8] bases

    const int tid = threadIdx.x, lane = tid & 31;
    for (int t = tid; t < TAB_WORDS4; t += ST) s_tab[t] = reinterpret_cast<const uint4 *>(&c_orient)[t];
    const int ntiles = (n + TILE - 1) / TILE;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = (int64_t)tile * TILE;
        const int cnt = min(TILE, (int)(n - base));
        // ---- keys of this thread's two envs (local ids tid and tid + ST)
        uint32_t key[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int lid = tid + h * ST;
            key[h] = 7u;
            if (lid < cnt) key[h] = sort_key(st[2 * stride + base + lid], st[3 * stride + base + lid]);
        }
        if (tid < 8) s_cnt[tid] = 0;
        __syncthreads();
        int wbase[2], rank[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned peers = __match_any_sync(0xFFFFFFFFu, key[h]);
            const int leader = __ffs(peers) - 1;
            rank[h] = __popc(peers & ((1u << lane) - 1u));
            int wb = 0;
            if (lane == leader) wb = atomicAdd(&s_cnt[key[h]], __popc(peers));
            wbase[h] = __shfl_sync(0xFFFFFFFFu, wb, leader);
        }
        __syncthreads();
        if (tid < 32) {                                                             // exclusive prefix over the 8 keys
            const int v = lane < 8 ? s_cnt[lane] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) { const int u = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += u; }
            if (lane < 8) s_cnt[8 + lane] = incl - v;
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) s_orig[s_cnt[8 + key[h]] + wbase[h] + rank[h]] = (uint16_t)(tid + h * ST);
        __syncthreads();
        {   // pull the next tile's key chunks towards L2 while this tile is enumerated
            const int64_t nb = base + (int64_t)gridDim.x * TILE + tid;
            if (nb < n) {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(st + 2 * stride + nb));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(st + 3 * stride + nb));
            }
            if (nb + ST < n) {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(st + 2 * stride + nb + ST));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(st + 3 * stride + nb + ST));
            }
        }
        // ---- enumerate sorted positions tid and TILE-1-tid (light + heavy half for every warp)
        {   // both records of this thread: start them towards L1 now, so the second gather overlaps the first enumeration
            const int l0 = s_orig[tid], l1 = s_orig[TILE - 1 - tid];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (l0 < cnt) asm volatile("prefetch.global.L1 [%0];" :: "l"(st + j * stride + base + l0));
                if (l1 < cnt) asm volatile("prefetch.global.L1 [%0];" :: "l"(st + j * stride + base + l1));
            }
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int lid = s_orig[h == 0 ? tid : TILE - 1 - tid];
            uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a, d = a;
            if (lid < cnt) {
                const int64_t i = base + lid;
                a = st[i]; b = st[stride + i]; c = st[2 * stride + i]; d = st[3 * stride + i];
            }
            Env e; unpack_env(a, b, c, d, e);
            TileSink sink{s_out + lid, nullptr};
            afterstates_env_impl<true>(e, s_tab, s_scr + tid, ST, L, M, sink);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");               // generic-proxy writes -> visible to the bulk engine
        __syncthreads();
        if (tid < 40) {                                                             // 40 rows of TILE words per tile
            for (int row = tid; row < 40; row += ST)
                bulk_store_row(words + (size_t)row * (size_t)n + (size_t)base, s_out + row * TILE, (uint32_t)cnt * 4u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");        // smem may be reused once it has been read
        }
        __syncthreads();
    }
}

// =================================================================================================
// fused hot-path step: move -> auto-reset of finished envs -> afterstates of the resulting state.
// One launch reads each 64-byte record once and writes it once; the memory time of the move hides under the
// integer work of the 40-slot enumeration (the three separate kernels read the state 2.25 times and write it twice).
// =================================================================================================
template <int MODE, bool P32 = false>
__global__ void TPL_AS_BOUNDS      // 128 registers: unconstrained, ptxas takes 166 and occupancy drops to 12 warps/SM
step_observe_kernel(uint4 *st, int64_t stride, int n, const uint8_t *__restrict__ rot, const uint8_t *__restrict__ loc,
                    int8_t *dlines, uint8_t *flags, int8_t *sto, unsigned long long *stats,
                    const uint4 *__restrict__ pool, int K, uint32_t *episode, uint32_t *tstep, uint64_t seed, uint64_t env_base, int gen_count,
                    uint32_t *__restrict__ words, uint8_t *__restrict__ aflags, float4 *__restrict__ ff, int L, int M, uint32_t one,
                    uint32_t *__restrict__ runs, uint32_t run_base, uint32_t *cursor, uint32_t *cursor_clear) {
    constexpr bool RAG = (MODE == 4);
    extern __shared__ __align__(16) uint32_t s_rag[];                 // RAG: one staging area per warp (RAG_SMEM_BYTES)
    __shared__ uint4 s_tab[TAB_WORDS4];
    __shared__ WarpQueue s_wq[THREADS / 32];
    TPL_SCRATCH;
    pdl_trigger();
    load_table(s_tab);
    __shared__ RecordStage s_stage[THREADS / 32], s_pstage[THREADS / 32];
    WarpQueue &q = s_wq[threadIdx.x >> 5];
    RecordStage &rs = s_stage[threadIdx.x >> 5], &pstage = s_pstage[threadIdx.x >> 5];
    uint32_t *stage = s_rag + (threadIdx.x >> 5) * RAG_STAGE_WORDS;
    const int wtiles = (n + 31) / 32, wstep = (int)gridDim.x * (THREADS / 32);
    int t = (int)blockIdx.x * (THREADS / 32) + (int)(threadIdx.x >> 5);
#ifndef TPL_NO_PDL_PREFETCH
    // Hints only (no data is read): the lines of the warp's first tile are asked into L2 while the previous kernel is still
    // finishing, so that the first loads after pdl_wait -- every warp of the grid issues them at the same moment, with nothing
    // else to run -- find them there.  (L2 is the point of coherence: a line the previous kernel still writes stays current.)
    if (t < wtiles) {
        const int64_t i0 = (int64_t)t * 32 + (threadIdx.x & 31);
        if (i0 < n) {
#pragma unroll
            for (int j = 0; j < 4; ++j) prefetch_l2(st + j * stride + i0);
            if ((threadIdx.x & 31) == 0) { prefetch_l2(rot + i0); prefetch_l2(loc + i0); }
            if (episode && (threadIdx.x & 7) == 0) prefetch_l2(episode + i0);
        }
    }
#endif
    pdl_wait();                                                      // nothing above reads or writes what another kernel produces
    if (RAG && cursor_clear && blockIdx.x == 0 && threadIdx.x == 0) *cursor_clear = 0u;      // the NEXT call's counter
    WqPos qp{0u, 0u, 0u, 0u};
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // The next tile's action and episode number ride in three registers, its records in shared memory (stage_issue); the episode
    // number of the tile after that rides in a fourth, because the pool record of the next tile is requested in the middle of this
    // one (pool_issue) and the number it is drawn from must have arrived by then.
    uint32_t nrot = 0, nloc = 0, nep = 0, nep2 = 0;
    if (t < wtiles) {
        const int64_t i0 = (int64_t)t * 32 + (threadIdx.x & 31), i1 = i0 + (int64_t)wstep * 32;
        stage_issue(rs, st, stride, i0, n);
        if (i0 < n) { nrot = rot[i0]; nloc = loc[i0]; nep = episode ? episode[i0] : 0u; }
        if (i1 < n) nep2 = episode ? episode[i1] : 0u;
        if (pool) pool_issue(pstage, pool, config_index(seed, env_base + (uint64_t)(i0 < n ? i0 : (int64_t)n - 1), nep + 1u, K));
    }
    // (warp-uniform trip count; one extra round after the warp's last tile drains its queue of deferred slots: see wq_publish)
    for (;; t += wstep) {
        const bool drain = t >= wtiles;
        const int64_t i = (int64_t)t * 32 + (threadIdx.x & 31), i2 = i + (int64_t)wstep * 32;
        PendingCtx cx; cx.mask = 0ull;
        Env e;
        uint32_t at = 0u;                                // what the resolver addresses the env's slots by
        if (!drain) {
            stage_take(rs, e);                           // each lane reads back only what it copied itself: no warp sync needed
            const uint32_t arot = nrot, aloc = nloc, ep1 = nep + 1u, ep1_next = nep2 + 1u;
            {
                const int64_t i3 = i2 + (int64_t)wstep * 32;
                if (t + wstep < wtiles) stage_issue(rs, st, stride, i2, n);
                if (i2 < n) { nrot = rot[i2]; nloc = loc[i2]; }
                nep = nep2;
                if (i3 < n) nep2 = episode ? episode[i3] : 0u;
            }
            if (i < n) {
                const uint32_t was = e.state;
                int k; bool changed;
                const uint32_t fl = step_env(e, s_tab, scr, THREADS, arot, aloc, L, M, k, changed);
                if (dlines) dlines[i] = (int8_t)k;
                if (flags) flags[i] = (uint8_t)fl;
                if (sto) sto[i] = (int8_t)e.state;
                acc[6] += 1; acc[4] += (uint32_t)k; acc[5] += changed ? 1u : 0u;
                if (was == S_RUNNING && e.state != S_RUNNING) {
                    acc[0] += 1;
                    if (fl & F_WIN) acc[1] += 1; else if (fl & F_TOPOUT) acc[2] += 1; else acc[3] += 1;
                }
                if (refill_queue(e, seed, env_base + (uint64_t)i, ep1 - 1u, gen_count)) changed = true;
                else if (pool && (e.state != S_RUNNING || e.head >= e.npieces)) {       // TPL_RESET_DONE semantics
                    if (episode) episode[i] = ep1;
                    if (tstep) tstep[i] = 0u;                                           // a new episode: the action stream of the rollouts restarts
                    const uint32_t lane = threadIdx.x & 31u;                            // (the record landed with this tile's env records)
                    install_record(e, pstage.chunk[0][lane], pstage.chunk[1][lane], pstage.chunk[2][lane], pstage.chunk[3][lane], seed,
                                   env_base + (uint64_t)i, ep1, gen_count);
                    acc[7] += 1;
                    changed = true;
                }
                if (changed) {
                    st[i] = make_uint4(e.col[0], e.col[1], e.col[2], e.col[3]);
                    st[stride + i] = make_uint4(e.col[4], e.col[5], e.col[6], e.col[7]);
                    st[2 * stride + i] = make_uint4(e.col[8], e.col[9], e.q[0], e.q[1]);
                }
                st[3 * stride + i] = pack_meta(e);
            }
            // the record the next tile's reset would install: on its way during the enumeration below
            if (pool && t + wstep < wtiles)
                pool_issue(pstage, pool, config_index(seed, env_base + (uint64_t)(i2 < n ? i2 : (int64_t)n - 1), ep1_next, K));
            if constexpr (RAG) {
                uint32_t piece = 7u, cnt = 0u;
                if (i < n) current_run(e, s_tab, piece, cnt);
                const RagTile rt = rag_reserve(cnt, cursor, one);
                rag_wait(false);
                if (i < n) {
                    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage);
                    RaggedStageSink sink{sbase + rt.excl * 4u, sbase + RAG_DUMMY_WORD * 4u, 0u};
                    afterstates_env(e, s_tab, scr, THREADS, L, M, sink, 0, 4, &cx, one);
                }
                at = rag_flush(stage, rt, words) + rt.excl;
                if (i < n) runs[i] = (at + run_base) | (piece << 29);         // (the resolver addresses `words + at`: call-local)
            } else {
                if (i < n) {
                    GlobalSink<MODE, P32> sink{words + i, aflags + i, ff + i, (uint32_t)n, one, nullptr};
                    afterstates_env(e, s_tab, scr, THREADS, L, M, sink, 0, 4, &cx, one);
                }
                at = (uint32_t)i;
            }
        }
        wq_publish<MODE>(q, qp, e, cx, at, drain, s_tab, scr, (uint32_t)n, words, aflags, ff, L, one);
        if (drain) break;
    }
    if (stats) flush_stats(acc, stats);
}

// =================================================================================================
// distinct placements -> the 40-slot grid (compact form): slot (r, c) of env i is placement
// (r % n_rot, min(c, 10 - w)) of its run (game/tetris.py:61, :364), with TPL_FLAG_ALIAS set when that differs from (r, c)
// =================================================================================================
__global__ void __launch_bounds__(THREADS)
expand_distinct_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ runs, int n, uint32_t *__restrict__ words) {
    __shared__ uint4 s_tab[TAB_WORDS4];
    load_table(s_tab);
    const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (i >= n) return;
    const uint32_t d = runs[i], piece = d >> 29;
    const uint32_t *run = rows + (d & 0x1FFFFFFFu);
    for (int r = 0; r < 4; ++r) {
        uint32_t rb = 0, cmax = 0, nrot = 1;
        if (piece < 7u) {
            nrot = (uint32_t)orient_nrot(s_tab[piece * 8]);
            const uint4 o = s_tab[(piece * 4 + (uint32_t)r % nrot) * 2], ob = s_tab[(piece * 4 + (uint32_t)r % nrot) * 2 + 1];
            rb = orient_rot_base(ob); cmax = (uint32_t)(COLS - orient_w(o));
        }
        for (int c = 0; c < COLS; ++c) {
            uint32_t w = F_NOPIECE << 3;
            if (piece < 7u) {
                const uint32_t cc = (uint32_t)c < cmax ? (uint32_t)c : cmax;
                w = run[rb + cc] | (((uint32_t)r >= nrot || cc != (uint32_t)c) ? (F_ALIAS << 3) : 0u);
            }
            words[(size_t)(r * 10 + c) * (size_t)n + (size_t)i] = w;
        }
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
    for (int base = 0, block = 0; base < count; base += QUEUE_PIECES, ++block) {
        uint32_t q[4];
        const int c = min(QUEUE_PIECES, count - base);
        gen_queue(seed, env_base + (uint64_t)i, episode ? episode[i] : episode0, c, q, (uint32_t)block);
        for (int p = 0; p < c; ++p) out[i * count + base + p] = (uint8_t)queue_piece(q, p);
    }
}

// =================================================================================================
// fused rollouts: state stays in registers for `steps` moves
// =================================================================================================
// ---- greedy policy: the deferred (row-completing) slots of a step are pooled over the warp ---------------------------------
// The arg-max needs every slot's value before the move, so the deferred slots cannot wait for later tiles as in the
// afterstate kernels.  Resolved in place, the warp runs the general move once per deferred slot of its BUSIEST lane while
// the other lanes idle; pooled, the (lane, slot) items of the whole warp are dealt out to consecutive lanes and one round
// usually covers them all.  Results come back to the owner through a 64-bit atomicMax on (value, lowest slot).
struct GreedyPool {
    uint32_t env[12 * 32];                         // [k][lane]: 10 columns, piece | cells << 8 | fl_noclear << 16, lines
    uint16_t items[40 * 32];                       // owner lane | slot << 5
    unsigned long long key[32];
};
__device__ __forceinline__ unsigned long long greedy_key(int v, int slot) {
    return ((unsigned long long)((uint32_t)v ^ 0x80000000u) << 8) | (unsigned long long)(255 - slot);
}

template <bool W16>
__device__ __forceinline__ void rollout_greedy_step_pooled(Env &e, uint32_t &ep, uint32_t &t, uint32_t (&acc)[8], const uint4 *tab,
                                                           uint32_t *scr, const uint4 *__restrict__ pool, int K, uint64_t seed, uint64_t env,
                                                           int gen_count, int L, int M, const GreedyWeights &gw, GreedyPool &gp, bool valid) {
    const uint32_t lane = threadIdx.x & 31u;
    if (valid && !refill_queue(e, seed, env, ep, gen_count) && (e.state != S_RUNNING || e.head >= e.npieces)) {
        ep += 1; t = 0; acc[7] += 1;
        install_config(e, pool, config_index(seed, env, ep, K), seed, env, ep, gen_count);
    }
    GreedySinkT<W16, true> sink{gw.w[0], gw.w[1], gw.w[2], gw.w[3], gw.w[4], gw.w[5], (int)0x80000000, 40};   // (deferred slots: pooled below)
    PendingCtx cx; cx.mask = 0ull;
    if (valid) afterstates_env(e, tab, scr, THREADS, L, M, sink, 0, 4, &cx);
    const uint32_t cnt = (uint32_t)__popcll(cx.mask);
    const uint32_t top = __reduce_max_sync(0xFFFFFFFFu, cnt);
    if (top) {
        const uint32_t incl = warp_prefix_small(cnt, top) + cnt;
        const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, cnt);
        gp.key[lane] = 0ull;
        if (cnt) {
#pragma unroll
            for (int k = 0; k < COLS; ++k) gp.env[k * 32 + lane] = e.col[k];
            gp.env[10 * 32 + lane] = cx.piece | (cx.cells << 8) | (cx.fl_noclear << 16);
            gp.env[11 * 32 + lane] = cx.lines;
            unsigned long long m = cx.mask;
            uint32_t at = incl - cnt;
            while (m) {
                const int sl = __ffsll((long long)m) - 1;
                m &= m - 1ull;
                gp.items[at++] = (uint16_t)(lane | ((uint32_t)sl << 5));
            }
        }
        __syncwarp();
        for (uint32_t base = 0; base < total; base += 32u) {
            const uint32_t k = base + lane;
            if (k < total) {
                const uint32_t item = gp.items[k], owner = item & 31u, slot = item >> 5;
                uint32_t cols[COLS];
#pragma unroll
                for (int j = 0; j < COLS; ++j) cols[j] = gp.env[j * 32 + owner];
                const uint32_t m0 = gp.env[10 * 32 + owner];
                const PendingCtx c2{0ull, m0 & 0xFFu, (m0 >> 8) & 0xFFu, gp.env[11 * 32 + owner], m0 >> 16};
                FirstPutSink fs{0u, 0u, false};
                resolve_slot(cols, c2, (int)slot, tab, scr, THREADS, L, fs);
                const int v = greedy_value<W16>(gw.w[0], gw.w[1], gw.w[2], gw.w[3], gw.w[4], gw.w[5], fs.word, fs.fl);
                atomicMax(&gp.key[owner], greedy_key(v, (int)slot));
            }
        }
        __syncwarp();
        const unsigned long long mine = gp.key[lane];
        if (mine > greedy_key(sink.best, sink.best_slot)) sink.best_slot = 255 - (int)(mine & 255ull);
        __syncwarp();                                  // the pool is reused by the next step
    }
    if (valid) {
        const int slot = sink.best_slot < 40 ? sink.best_slot : 0;
        const uint32_t rot = (uint32_t)(slot / 10), loc = (uint32_t)(slot - 10 * (slot / 10));
        int k; bool changed;
        const uint32_t fl = step_env(e, tab, scr, THREADS, rot, loc, L, M, k, changed);
        t += 1;
        acc[6] += 1; acc[4] += (uint32_t)k; acc[5] += changed ? 1u : 0u;
        if (e.state != S_RUNNING) {
            acc[0] += 1;
            if (fl & F_WIN) acc[1] += 1; else if (fl & F_TOPOUT) acc[2] += 1; else acc[3] += 1;
        }
    }
}

#ifndef TPL_RO_MINBLOCKS
#define TPL_RO_MINBLOCKS 4          // 128 registers (unconstrained the greedy variant takes 161: 12 warps per SM)
#endif
template <int POLICY>               // 0 random agent, 1 greedy (int32 weights), 2 greedy with int16 feature weights (dot-product form)
__global__ void __launch_bounds__(THREADS, POLICY == 0 ? 7 : TPL_RO_MINBLOCKS)      // the random agent needs 66 registers only
rollout_kernel(uint4 *st, int64_t stride, int n, const uint4 *__restrict__ pool, int K, uint32_t *episode,
               uint32_t *tstep, unsigned long long *stats, int steps, uint64_t seed, uint64_t env_base,
               int gen_count, int L, int M, GreedyWeights gw) {
    __shared__ uint4 s_tab[TAB_WORDS4];
    __shared__ GreedyPool s_gp[POLICY == 0 ? 1 : THREADS / 32];
    TPL_SCRATCH;
    load_table(s_tab);
    const int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    const bool valid = i < n;
    const int64_t ic = valid ? i : (int64_t)n - 1;                       // lanes past the end shadow the last env and store nothing
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint64_t env = env_base + (uint64_t)ic;
    Env e; load_env(st, stride, ic, e);
    uint32_t ep = episode[ic], t = tstep[ic];
    for (int s = 0; s < steps; ++s) {
        if (POLICY == 2) rollout_greedy_step_pooled<true>(e, ep, t, acc, s_tab, scr, pool, K, seed, env, gen_count, L, M, gw, s_gp[threadIdx.x >> 5], valid);
        else if (POLICY == 1) rollout_greedy_step_pooled<false>(e, ep, t, acc, s_tab, scr, pool, K, seed, env, gen_count, L, M, gw, s_gp[threadIdx.x >> 5], valid);
        else if (valid) rollout_random_step(e, ep, t, acc, s_tab, scr, THREADS, pool, K, seed, env, gen_count, L, M);
    }
    if (valid) {
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
// Cached per device: the launches below go to the CUDA *current* device, which need not be the one the first call saw.
struct DevInfo { int sms, occ_step, occ_as, occ_so, occ_sod, occ_asd; };
static DevInfo &dev_info() {
    static DevInfo info[64] = {};
    int dev = 0; cudaGetDevice(&dev);
    return info[(dev >= 0 && dev < 64) ? dev : 0];
}
static int sm_count() {
    DevInfo &d = dev_info();
    if (!d.sms) {
        int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || d.sms <= 0) d.sms = 148;
    }
    return d.sms;
}

// The piece-sorted afterstate kernel (opt-in: TPL_SORTED_AFTERSTATES=1) moves its output tile with 16-byte-granular
// bulk copies.  It executes 30 % fewer instructions than the plain kernel (58 M vs 82 M warp-instructions at 2^20 envs)
// but exposes the record-gather latency and four CTA barriers per tile, and ends up at the same 0.12 ms
// (profiles/r01_ncu_full_v3_afterstates_sorted.txt); it stays off until its loads are software-pipelined.
static bool sorted_path_ok(int n, const void *feats) {
    static int enabled = -1;
    if (enabled < 0) { const char *v = getenv("TPL_SORTED_AFTERSTATES"); enabled = (v && v[0] == '1') ? 1 : 0; }
    return enabled && n >= 2 * TILE && (n % 4) == 0 && ((uintptr_t)feats % 16) == 0;
}

// the compact output array [40][n] words does not cross a 4 GB-aligned address boundary (GlobalSink<.., P32>)
// (TPL_NO_P32=1, read once per process, forces the 64-bit pointer chain: lets the tests cover that path too)
static bool one_window(const void *feats, int n) {
    static int disabled = -1;
    if (disabled < 0) { const char *v = getenv("TPL_NO_P32"); disabled = (v && v[0] == '1') ? 1 : 0; }
    const uintptr_t a = (uintptr_t)feats, b = a + (uintptr_t)160 * (uintptr_t)n - 1;
    return !disabled && (a >> 32) == (b >> 32);
}

// Launch with programmatic dependent launch allowed (see pdl_trigger / pdl_wait): the prologue of this kernel may overlap the
// tail of the previous kernel in the stream.  TPL_NO_PDL=1 (read once per process) launches plainly, for comparisons.
template <class... KArgs, class... Args>
static void launch_pdl(void (*kernel)(KArgs...), unsigned grid, size_t dyn_smem, cudaStream_t s, Args... args) {
    static int disabled = -1;
    if (disabled < 0) { const char *v = getenv("TPL_NO_PDL"); disabled = (v && v[0] == '1') ? 1 : 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = dyn_smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = disabled ? 0u : 1u;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);       // (errors surface through check_launch -> cudaGetLastError)
}

static unsigned grid_persistent(int n, int blocks_per_sm) {
    const int sms = sm_count();
    if (blocks_per_sm <= 0) blocks_per_sm = 16;
    const unsigned need = grid_for(n), cap = (unsigned)(sms * blocks_per_sm);
    return need < cap ? need : cap;
}

// distinct-placements argument checks shared by the two entry points
static int check_distinct(const char *who, int n, const uint32_t *rows, int64_t rows_capacity, const uint32_t *runs, const uint32_t *cursor2,
                          int phase) {
    if (!rows || !runs || !cursor2) return fail(TPL_EINVAL, "%s: rows / runs / cursor2 must not be null", who);
    if (phase != 0 && phase != 1) return fail(TPL_EINVAL, "%s: phase must be 0 or 1", who);
    if (n > (1 << 23)) return fail(TPL_ERANGE, "%s: at most 2^23 envs per call (29-bit run offsets)", who);
    if (rows_capacity < TPL_DISTINCT_CAPACITY(n)) return fail(TPL_ERANGE, "%s: rows_capacity %lld < TPL_DISTINCT_CAPACITY(n) = %lld words", who,
                                                             (long long)rows_capacity, (long long)TPL_DISTINCT_CAPACITY(n));
    if (((uintptr_t)rows & 15u) != 0) return fail(TPL_EINVAL, "%s: rows must be 16-byte aligned", who);
    return 0;
}
template <class KernelT>
static int allow_rag_smem(KernelT kernel, const char *who) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RAG_SMEM_BYTES);
    if (e != cudaSuccess) return fail((int)e, "%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
    return 0;
}

static int step_observe_impl(const char *who, void *state, int64_t plane_stride, int n, const uint8_t *rot, const uint8_t *loc, int8_t *dlines,
                             uint8_t *flags, int8_t *st, long long *stats, const void *pool, int K, uint32_t *episode, uint32_t *tstep,
                             uint64_t seed, uint64_t env_base, int gen_count, uint8_t *feats, uint8_t *aflags, float *feats_f32,
                             uint32_t *rows, uint32_t *runs, uint32_t run_base, uint32_t *cursor2, int phase, int L, int M, void *stream) {
    if (n < 0 || !state || !rot || !loc) return fail(TPL_EINVAL, "%s: null argument", who);
    if (plane_stride < n) return fail(TPL_ERANGE, "%s: plane_stride < n", who);
    if (n > (1 << 25)) return fail(TPL_ERANGE, "%s: at most 2^25 envs per call (32-bit output offsets)", who);
    if (pool && K <= 0) return fail(TPL_EINVAL, "%s: pool given but K <= 0", who);
    if (gen_count < 0 || gen_count > GEN_MAX) return fail(TPL_ERANGE, "%s: gen_count %d > %d", who, gen_count, GEN_MAX);
    if (L < 0 || M < 0 || M > 65535 || L > 65535) return fail(TPL_ERANGE, "%s: L/M out of range", who);
    if (n == 0) return 0;
    const cudaStream_t s = (cudaStream_t)stream;
    uint4 *sp = (uint4 *)state; const uint4 *pp = (const uint4 *)pool; uint32_t *w = (uint32_t *)feats; float4 *f = (float4 *)feats_f32;
    unsigned long long *sq = (unsigned long long *)stats;
    // exactly one resident wave of persistent CTAs (4 per SM at 128 registers): every warp walks through its share of the
    // 32-env tiles, and the partly filled last round of its deferred-slot queue is paid once per warp
    int &so_blocks_per_sm = rows ? dev_info().occ_sod : dev_info().occ_so;
    if (rows) {
        if (!so_blocks_per_sm) {
            int rc = allow_rag_smem(step_observe_kernel<4>, who); if (rc) return rc;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&so_blocks_per_sm, step_observe_kernel<4>, THREADS, RAG_SMEM_BYTES) != cudaSuccess || so_blocks_per_sm <= 0)
                so_blocks_per_sm = 4;
        }
        launch_pdl(step_observe_kernel<4>, grid_persistent(n, so_blocks_per_sm), RAG_SMEM_BYTES, s,
                   sp, plane_stride, n, rot, loc, dlines, flags, st, sq, pp, K, episode, tstep, seed, env_base, gen_count, rows, nullptr, nullptr, L, M, 1u,
                   runs, run_base, cursor2 + phase, cursor2 + (phase ^ 1));
        return check_launch(who);
    }
    if (!so_blocks_per_sm &&
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&so_blocks_per_sm, step_observe_kernel<0>, THREADS, 0) != cudaSuccess || so_blocks_per_sm <= 0))
        so_blocks_per_sm = 4;
    const unsigned g = grid_persistent(n, so_blocks_per_sm);
#define TPL_SO(...) launch_pdl(step_observe_kernel<__VA_ARGS__>, g, 0, s, sp, plane_stride, n, rot, loc, dlines, flags, st, sq, pp, K, episode, tstep, \
                               seed, env_base, gen_count, w, aflags, f, L, M, 1u, nullptr, 0u, nullptr, nullptr)
    if (feats && !aflags && one_window(feats, n)) TPL_SO(0, true);
    else if (feats && !aflags) TPL_SO(0);
    else if (feats && !feats_f32) TPL_SO(1);
    else if (!feats) TPL_SO(2);
    else TPL_SO(3);
#undef TPL_SO
    return check_launch(who);
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
                        const uint8_t *mask, int mode, uint32_t *episode, uint32_t *tstep, uint64_t seed, uint64_t env_base, int gen_count,
                        void *stream) {
    if (n < 0 || !state || !pool || K <= 0) return fail(TPL_EINVAL, "tpl_reset_from_pool: null state/pool or K <= 0");
    if (plane_stride < n) return fail(TPL_ERANGE, "tpl_reset_from_pool: plane_stride < n");
    if (mode < TPL_RESET_ALL || mode > TPL_RESET_DONE) return fail(TPL_EINVAL, "tpl_reset_from_pool: bad mode %d", mode);
    if (mode == TPL_RESET_MASK && !mask) return fail(TPL_EINVAL, "tpl_reset_from_pool: mode MASK needs a mask");
    if (gen_count < 0 || gen_count > GEN_MAX) return fail(TPL_ERANGE, "tpl_reset_from_pool: gen_count %d > %d", gen_count, GEN_MAX);
    if (n == 0) return 0;
    reset_kernel<<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>((uint4 *)state, plane_stride, n, (const uint4 *)pool, K, idx,
                                                                     mask, mode, episode, tstep, seed, env_base, gen_count);
    return check_launch("tpl_reset_from_pool");
}

int tpl_step(void *state, int64_t plane_stride, int n, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags,
             int8_t *st, long long *stats, int L, int M, void *stream) {
    if (n < 0 || !state || !rot || !loc) return fail(TPL_EINVAL, "tpl_step: null argument");
    if (plane_stride < n) return fail(TPL_ERANGE, "tpl_step: plane_stride < n");
    if (L < 0 || M < 0 || M > 65535 || L > 65535) return fail(TPL_ERANGE, "tpl_step: L/M out of range");
    if (n == 0) return 0;
    int &step_blocks_per_sm = dev_info().occ_step;     // exactly one resident wave: no partial second wave at the tail
    if (!step_blocks_per_sm &&
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&step_blocks_per_sm, step_kernel, THREADS, 0) != cudaSuccess || step_blocks_per_sm <= 0))
        step_blocks_per_sm = 8;
    launch_pdl(step_kernel, grid_persistent(n, step_blocks_per_sm), 0, (cudaStream_t)stream, (uint4 *)state, plane_stride, n, rot, loc, dlines, flags, st,
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
    int &as_blocks_per_sm = dev_info().occ_as;         // one resident wave, as for the fused kernel
    if (!as_blocks_per_sm &&
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&as_blocks_per_sm, afterstates_kernel<0>, THREADS, 0) != cudaSuccess || as_blocks_per_sm <= 0))
        as_blocks_per_sm = 4;
    const unsigned g = grid_persistent(n, as_blocks_per_sm);
    if (feats && !flags && !feats_f32 && sorted_path_ok(n, feats)) {
        {   // per device and cheap; this path is an opt-in experiment
            cudaError_t e = cudaFuncSetAttribute(afterstates_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM_BYTES);
            if (e != cudaSuccess) return fail((int)e, "tpl_afterstates: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        }
        const int ntiles = (n + TILE - 1) / TILE;
        const unsigned gs = (unsigned)(ntiles < 4 * sm_count() ? ntiles : 4 * sm_count());
        afterstates_sorted_kernel<<<gs, ST, SORT_SMEM_BYTES, s>>>(st, plane_stride, n, w, L, M);
        return check_launch("tpl_afterstates(sorted)");
    }
    if (grid_for(n) < 2u * (unsigned)sm_count()) {                 // fewer than two CTAs per SM: split envs over 4 threads
        const unsigned g4 = (unsigned)(((int64_t)n * 4 + THREADS - 1) / THREADS);
        if (feats && !flags) afterstates_split_kernel<0><<<g4, THREADS, 0, s>>>(st, plane_stride, n, w, flags, f, L, M, 1u);
        else if (feats && !feats_f32) afterstates_split_kernel<1><<<g4, THREADS, 0, s>>>(st, plane_stride, n, w, flags, f, L, M, 1u);
        else if (!feats) afterstates_split_kernel<2><<<g4, THREADS, 0, s>>>(st, plane_stride, n, w, flags, f, L, M, 1u);
        else afterstates_split_kernel<3><<<g4, THREADS, 0, s>>>(st, plane_stride, n, w, flags, f, L, M, 1u);
        return check_launch("tpl_afterstates(split)");
    }
    if (feats && !flags && one_window(feats, n)) launch_pdl(afterstates_kernel<0, true>, g, 0, s, st, plane_stride, n, w, flags, f, L, M, 1u, nullptr, 0u, nullptr, nullptr);
    else if (feats && !flags) launch_pdl(afterstates_kernel<0>, g, 0, s, st, plane_stride, n, w, flags, f, L, M, 1u, nullptr, 0u, nullptr, nullptr);
    else if (feats && !feats_f32) launch_pdl(afterstates_kernel<1>, g, 0, s, st, plane_stride, n, w, flags, f, L, M, 1u, nullptr, 0u, nullptr, nullptr);
    else if (!feats) launch_pdl(afterstates_kernel<2>, g, 0, s, st, plane_stride, n, w, flags, f, L, M, 1u, nullptr, 0u, nullptr, nullptr);
    else launch_pdl(afterstates_kernel<3>, g, 0, s, st, plane_stride, n, w, flags, f, L, M, 1u, nullptr, 0u, nullptr, nullptr);
    return check_launch("tpl_afterstates");
}

int tpl_step_observe(void *state, int64_t plane_stride, int n, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags,
                     int8_t *st, long long *stats, const void *pool, int K, uint32_t *episode, uint32_t *tstep, uint64_t seed,
                     uint64_t env_base, int gen_count, uint8_t *feats, uint8_t *aflags, float *feats_f32, int L, int M, void *stream) {
    if (!feats && !feats_f32) return fail(TPL_EINVAL, "tpl_step_observe: no feature output requested");
    if (feats_f32 && !aflags) return fail(TPL_EINVAL, "tpl_step_observe: the float form needs the flags array");
    return step_observe_impl("tpl_step_observe", state, plane_stride, n, rot, loc, dlines, flags, st, stats, pool, K, episode, tstep, seed,
                             env_base, gen_count, feats, aflags, feats_f32, nullptr, nullptr, 0u, nullptr, 0, L, M, stream);
}

int tpl_step_observe_distinct(void *state, int64_t plane_stride, int n, const uint8_t *rot, const uint8_t *loc, int8_t *dlines,
                              uint8_t *flags, int8_t *st, long long *stats, const void *pool, int K, uint32_t *episode, uint32_t *tstep,
                              uint64_t seed, uint64_t env_base, int gen_count, uint32_t *rows, int64_t rows_capacity, uint32_t *runs,
                              uint32_t run_base, uint32_t *cursor2, int phase, int L, int M, void *stream) {
    int rc = check_distinct("tpl_step_observe_distinct", n, rows, rows_capacity, runs, cursor2, phase); if (rc) return rc;
    return step_observe_impl("tpl_step_observe_distinct", state, plane_stride, n, rot, loc, dlines, flags, st, stats, pool, K, episode, tstep,
                             seed, env_base, gen_count, nullptr, nullptr, nullptr, rows, runs, run_base, cursor2, phase, L, M, stream);
}

int tpl_afterstates_distinct(const void *state, int64_t plane_stride, int n, uint32_t *rows, int64_t rows_capacity, uint32_t *runs,
                             uint32_t run_base, uint32_t *cursor2, int phase, int L, int M, void *stream) {
    if (n < 0 || !state) return fail(TPL_EINVAL, "tpl_afterstates_distinct: null state");
    if (plane_stride < n) return fail(TPL_ERANGE, "tpl_afterstates_distinct: plane_stride < n");
    int rc = check_distinct("tpl_afterstates_distinct", n, rows, rows_capacity, runs, cursor2, phase); if (rc) return rc;
    if (n == 0) return 0;
    int &bps = dev_info().occ_asd;
    if (!bps) {
        rc = allow_rag_smem(afterstates_kernel<4>, "tpl_afterstates_distinct"); if (rc) return rc;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, afterstates_kernel<4>, THREADS, RAG_SMEM_BYTES) != cudaSuccess || bps <= 0) bps = 4;
    }
    launch_pdl(afterstates_kernel<4>, grid_persistent(n, bps), RAG_SMEM_BYTES, (cudaStream_t)stream,
               (const uint4 *)state, plane_stride, n, rows, nullptr, nullptr, L, M, 1u, runs, run_base, cursor2 + phase, cursor2 + (phase ^ 1));
    return check_launch("tpl_afterstates_distinct");
}

void tpl_distinct_tables(uint8_t *count, uint8_t *slot_of, uint8_t *canon_of) {
    constexpr OrientTable t = make_table();
    for (int p = 0; p < 7; ++p) {
        const int nrot = (int)((t.e[p * 4].ax >> 24) & 3u) + 1;
        if (count) count[p] = (uint8_t)((t.e[p * 4].bw >> 8) & 0xFFu);
        if (slot_of) for (int j = 0; j < TPL_DISTINCT_MAX; ++j) slot_of[p * TPL_DISTINCT_MAX + j] = 255;
        for (int r = 0; r < 4; ++r) {
            const OrientEntry &e = t.e[p * 4 + r % nrot];
            const int w = (int)((e.ax >> 16) & 7u), rb = (int)(e.bw & 0xFFu);
            for (int c = 0; c < COLS; ++c) {
                const int cc = c < COLS - w ? c : COLS - w;
                if (canon_of) canon_of[p * 40 + r * 10 + c] = (uint8_t)(rb + cc);
                if (slot_of && r < nrot && c == cc) slot_of[p * TPL_DISTINCT_MAX + rb + c] = (uint8_t)(r * 10 + c);
            }
        }
    }
}

int tpl_expand_distinct(const uint32_t *rows, const uint32_t *runs, int n, uint8_t *feats, void *stream) {
    if (n < 0 || !rows || !runs || !feats) return fail(TPL_EINVAL, "tpl_expand_distinct: null argument");
    if (n == 0) return 0;
    expand_distinct_kernel<<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>(rows, runs, n, (uint32_t *)feats);
    return check_launch("tpl_expand_distinct");
}

int tpl_gen_pieces(uint8_t *out, int n, int count, uint64_t seed, uint64_t env_base, const uint32_t *episode, uint32_t episode0,
                   void *stream) {
    if (n < 0 || !out) return fail(TPL_EINVAL, "tpl_gen_pieces: null output");
    if (count < 0 || count > GEN_MAX) return fail(TPL_ERANGE, "tpl_gen_pieces: count %d outside 0..%d", count, GEN_MAX);
    if (n == 0 || count == 0) return 0;
    gen_pieces_kernel<<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>(out, n, count, seed, env_base, episode, episode0);
    return check_launch("tpl_gen_pieces");
}

static int rollout_common(const char *who, bool greedy, void *state, int64_t plane_stride, int n, const void *pool, int K,
                          uint32_t *episode, uint32_t *tstep, long long *stats, int steps, const int32_t *w6, uint64_t seed,
                          uint64_t env_base, int gen_count, int L, int M, void *stream) {
    if (n < 0 || !state || !pool || K <= 0 || !episode || !tstep || !stats) return fail(TPL_EINVAL, "%s: null argument or K <= 0", who);
    if (plane_stride < n) return fail(TPL_ERANGE, "%s: plane_stride < n", who);
    if (gen_count < 0 || gen_count > GEN_MAX) return fail(TPL_ERANGE, "%s: gen_count %d > %d", who, gen_count, GEN_MAX);
    if (greedy && !w6) return fail(TPL_EINVAL, "%s: null weights", who);
    if (n == 0 || steps <= 0) return 0;
    GreedyWeights gw{};
    if (greedy) for (int q = 0; q < 6; ++q) gw.w[q] = w6[q];
    bool w16 = greedy;
    for (int q = 0; greedy && q < 4; ++q) w16 = w16 && gw.w[q] >= -32768 && gw.w[q] <= 32767;
#define TPL_RO(P) rollout_kernel<P><<<grid_for(n), THREADS, 0, (cudaStream_t)stream>>>((uint4 *)state, plane_stride, n, (const uint4 *)pool, K, \
                                episode, tstep, (unsigned long long *)stats, steps, seed, env_base, gen_count, L, M, gw)
    if (w16) TPL_RO(2);
    else if (greedy) TPL_RO(1);
    else TPL_RO(0);
#undef TPL_RO
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
