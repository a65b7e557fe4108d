// piclim_host_api.cu -- host-buffer half of the C ABI (tpl_env_*): an opaque handle owns the device state,
// a stream and device-side staging; callers pass HOST pointers, exactly what a ctypes/cffi binding inside the
// reference's game/tetris.py would hold (numpy arrays).  Built on the device-pointer entry points only.
#include "../../include/tetris_piclim.h"

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace tpl { int fail(int code, const char *fmt, ...); }
using tpl::fail;

#ifndef TPL_MAX_CHUNKS
#define TPL_MAX_CHUNKS 8
#endif
struct tpl_env {
    int n = 0, L = 0, M = 0, device = 0;
    uint64_t seed = 0, env_base = 0;
    cudaStream_t stream = nullptr;
    void *state = nullptr; int64_t stride = 0;
    void *pool = nullptr; int K = 0;
    uint32_t *episode = nullptr;
    // device staging for the host-facing calls
    uint8_t *d_rot = nullptr, *d_loc = nullptr, *d_flags = nullptr;
    int8_t *d_dlines = nullptr, *d_st = nullptr;
    uint8_t *d_feats = nullptr, *d_aflags = nullptr;                 // [40][n][4], [40][n]
    uint16_t *d_rows = nullptr; uint8_t *d_cur = nullptr, *d_next = nullptr, *d_head = nullptr, *d_np = nullptr, *d_queue = nullptr;
    int32_t *d_lines = nullptr, *d_moves = nullptr;
    void *d_scratch = nullptr; size_t scratch_bytes = 0;            // uploads for load / set_pool / reset
    // pipelined rollout step (tpl_env_step_observe*): the envs are cut into chunks; the kernel of chunk c + 1 runs on `stream`
    // while the results of chunk c travel to the host on `copy_stream`
    int nchunks = 1; int chunk_envs = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_done[TPL_MAX_CHUNKS] = {};
    uint32_t *d_tstep = nullptr;
    uint32_t *d_drows = nullptr, *d_runs = nullptr, *d_cursor2 = nullptr;   // distinct-placements form: rows regions per chunk, descriptors, 2 counters per chunk
    uint32_t *h_cursor = nullptr;                                    // pinned: words each chunk produced
    int phase = 0;
};

#define CU(call)                                                                                     \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail((int)e_, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
#define RC(call) do { int r_ = (call); if (r_) return r_; } while (0)

static int ensure(void **p, size_t bytes) {
    if (*p) return 0;
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail((int)e, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return 0;
}
// the afterstate words: an allocation that does not cross a 4 GB-aligned address boundary lets the kernels advance their
// store pointer with a 32-bit add (GlobalSink<.., P32>); a straddling one (rare) is swapped for another
static int ensure_feats(tpl_env *e, size_t bytes) {
    if (e->d_feats) return 0;
    void *held[4]; int nheld = 0;
    int rc = 0;
    for (;;) {
        void *p = nullptr;
        rc = ensure(&p, bytes);
        if (rc) break;
        const uintptr_t a = (uintptr_t)p;
        if ((a >> 32) == ((a + bytes - 1) >> 32) || nheld == 4) { e->d_feats = (uint8_t *)p; break; }
        held[nheld++] = p;
    }
    for (int k = 0; k < nheld; ++k) cudaFree(held[k]);
    return rc;
}
static int ensure_scratch(tpl_env *e, size_t bytes) {
    if (e->scratch_bytes >= bytes) return 0;
    if (e->d_scratch) cudaFree(e->d_scratch);
    e->d_scratch = nullptr; e->scratch_bytes = 0;
    RC(ensure(&e->d_scratch, bytes));
    e->scratch_bytes = bytes;
    return 0;
}
static inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

extern "C" {

void *tpl_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void tpl_host_free(void *p) { if (p) cudaFreeHost(p); }

int tpl_env_create(tpl_env **out, int n, int L, int M, int device, uint64_t seed, uint64_t env_base) {
    if (!out || n <= 0) return fail(TPL_EINVAL, "tpl_env_create: n must be positive");
    if (L < 0 || M < 0 || L > 65535 || M > 65535) return fail(TPL_ERANGE, "tpl_env_create: L/M out of range");
    CU(cudaSetDevice(device));
    tpl_env *e = new (std::nothrow) tpl_env();
    if (!e) return fail(TPL_ENOMEM, "tpl_env_create: out of host memory");
    e->n = n; e->L = L; e->M = M; e->device = device; e->seed = seed; e->env_base = env_base;
    e->stride = ((int64_t)n + 31) / 32 * 32;
    // chunks of the pipelined step: multiples of 32 envs (a chunk starts on a tile boundary of the state planes), at least
    // 256 Ki envs each, two by default -- measured at 2^20 envs (profiles/r02_e2e_chunks.txt): 1 chunk 2.054 ms, 2: 2.047,
    // 4: 2.104, 8: 2.204 per distinct-form step; the transfer is 1.83 ms of it at the link's rate, so there is little
    // to overlap and every extra chunk costs five small copies (TPL_ENV_CHUNKS overrides, up to 8)
    e->nchunks = n / 262144; if (e->nchunks < 1) e->nchunks = 1; if (e->nchunks > 2) e->nchunks = 2;
    if (const char *v = getenv("TPL_ENV_CHUNKS")) { int c = atoi(v); if (c >= 1 && c <= TPL_MAX_CHUNKS) e->nchunks = c; }
    e->chunk_envs = (int)((((int64_t)n + e->nchunks - 1) / e->nchunks + 31) / 32 * 32);
    e->nchunks = (n + e->chunk_envs - 1) / e->chunk_envs;
    cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking);
    for (int c = 0; c < e->nchunks && err == cudaSuccess; ++c) err = cudaEventCreateWithFlags(&e->ev_done[c], cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaMalloc(&e->state, (size_t)e->stride * 64);
    if (err == cudaSuccess) err = cudaMemsetAsync(e->state, 0, (size_t)e->stride * 64, e->stream);
    if (err == cudaSuccess) err = cudaMalloc((void **)&e->episode, (size_t)n * 4);
    if (err == cudaSuccess) err = cudaMemsetAsync(e->episode, 0, (size_t)n * 4, e->stream);
    if (err == cudaSuccess) err = cudaMalloc((void **)&e->d_tstep, (size_t)n * 4);
    if (err == cudaSuccess) err = cudaMemsetAsync(e->d_tstep, 0, (size_t)n * 4, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) { int r = fail((int)err, "tpl_env_create: %s", cudaGetErrorString(err)); tpl_env_destroy(e); return r; }
    *out = e;
    return 0;
}

void tpl_env_destroy(tpl_env *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
    for (int c = 0; c < TPL_MAX_CHUNKS; ++c) if (e->ev_done[c]) cudaEventDestroy(e->ev_done[c]);
    if (e->h_cursor) cudaFreeHost(e->h_cursor);
    void *ptrs[] = {e->d_tstep, e->d_drows, e->d_runs, e->d_cursor2, e->state, e->pool, e->episode, e->d_rot, e->d_loc, e->d_flags, e->d_dlines, e->d_st, e->d_feats, e->d_aflags,
                    e->d_rows, e->d_cur, e->d_next, e->d_head, e->d_np, e->d_queue, e->d_lines, e->d_moves, e->d_scratch};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int tpl_env_set_limits(tpl_env *e, int L, int M) {
    if (!e) return fail(TPL_EINVAL, "tpl_env_set_limits: null handle");
    if (L < 0 || M < 0 || L > 65535 || M > 65535) return fail(TPL_ERANGE, "tpl_env_set_limits: L/M out of range");
    e->L = L; e->M = M;
    return 0;
}

void *tpl_env_state_ptr(tpl_env *e, int64_t *plane_stride) { if (plane_stride) *plane_stride = e->stride; return e->state; }
void *tpl_env_stream(tpl_env *e) { return (void *)e->stream; }

// upload (rows, pieces, npieces [, lines, moves, st, head]) for m items into scratch and pack them
static int upload_pack(tpl_env *e, void *out, int64_t stride, int aos, int m, const uint16_t *rows, const uint8_t *pieces,
                       int pstride, const uint8_t *npieces, const int32_t *lines, const int32_t *moves, const int8_t *st,
                       const uint8_t *head) {
    if (!rows || !pieces || !npieces || pstride <= 0) return fail(TPL_EINVAL, "upload: rows/pieces/npieces required");
    for (int i = 0; i < m; ++i) {
        if (npieces[i] > TPL_MAX_PIECES || npieces[i] > pstride)
            return fail(TPL_ERANGE, "config %d has %d pieces (> min(42, pieces_stride = %d))", i, npieces[i], pstride);
        for (int q = 0; q < npieces[i]; ++q)
            if (pieces[(size_t)i * pstride + q] > 6) return fail(TPL_ERANGE, "config %d: piece id %d at position %d (ids are 0..6)", i, pieces[(size_t)i * pstride + q], q);
    }
    const size_t b_rows = al16((size_t)m * 40), b_p = al16((size_t)m * pstride), b_np = al16((size_t)m), b_i32 = al16((size_t)m * 4);
    RC(ensure_scratch(e, b_rows + b_p + 3 * b_np + 2 * b_i32));
    char *base = (char *)e->d_scratch; size_t off = 0;
    uint16_t *d_rows = (uint16_t *)(base + off); off += b_rows;
    uint8_t *d_p = (uint8_t *)(base + off); off += b_p;
    uint8_t *d_np = (uint8_t *)(base + off); off += b_np;
    int32_t *d_lines = (int32_t *)(base + off); off += b_i32;
    int32_t *d_moves = (int32_t *)(base + off); off += b_i32;
    int8_t *d_st = (int8_t *)(base + off); off += b_np;
    uint8_t *d_head = (uint8_t *)(base + off);
    CU(cudaMemcpyAsync(d_rows, rows, (size_t)m * 40, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_p, pieces, (size_t)m * pstride, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(d_np, npieces, (size_t)m, cudaMemcpyHostToDevice, e->stream));
    if (lines) CU(cudaMemcpyAsync(d_lines, lines, (size_t)m * 4, cudaMemcpyHostToDevice, e->stream));
    if (moves) CU(cudaMemcpyAsync(d_moves, moves, (size_t)m * 4, cudaMemcpyHostToDevice, e->stream));
    if (st) CU(cudaMemcpyAsync(d_st, st, (size_t)m, cudaMemcpyHostToDevice, e->stream));
    if (head) CU(cudaMemcpyAsync(d_head, head, (size_t)m, cudaMemcpyHostToDevice, e->stream));
    RC(tpl_pack(out, stride, aos, m, d_rows, d_p, pstride, d_np, lines ? d_lines : nullptr, moves ? d_moves : nullptr,
                st ? d_st : nullptr, head ? d_head : nullptr, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

int tpl_env_set_pool(tpl_env *e, int K, const uint16_t *rows, const uint8_t *pieces, int pieces_stride, const uint8_t *npieces) {
    if (!e || K <= 0) return fail(TPL_EINVAL, "tpl_env_set_pool: K must be positive");
    CU(cudaSetDevice(e->device));
    if (e->pool) { cudaFree(e->pool); e->pool = nullptr; e->K = 0; }
    RC(ensure(&e->pool, (size_t)K * 64));
    RC(upload_pack(e, e->pool, 0, 1, K, rows, pieces, pieces_stride, npieces, nullptr, nullptr, nullptr, nullptr));
    e->K = K;
    return 0;
}

int tpl_env_load(tpl_env *e, const uint16_t *rows, const uint8_t *pieces, int pieces_stride, const uint8_t *npieces,
                 const int32_t *lines, const int32_t *moves, const int8_t *st, const uint8_t *head) {
    if (!e) return fail(TPL_EINVAL, "tpl_env_load: null handle");
    CU(cudaSetDevice(e->device));
    return upload_pack(e, e->state, e->stride, 0, e->n, rows, pieces, pieces_stride, npieces, lines, moves, st, head);
}

int tpl_env_reset(tpl_env *e, const int32_t *idx, const uint8_t *mask, int mode, int gen_count) {
    if (!e) return fail(TPL_EINVAL, "tpl_env_reset: null handle");
    if (!e->pool) return fail(TPL_EINVAL, "tpl_env_reset: no config pool (call tpl_env_set_pool first)");
    CU(cudaSetDevice(e->device));
    const size_t b_idx = al16((size_t)e->n * 4);
    RC(ensure_scratch(e, b_idx + al16((size_t)e->n)));
    int32_t *d_idx = (int32_t *)e->d_scratch; uint8_t *d_mask = (uint8_t *)e->d_scratch + b_idx;
    if (idx) CU(cudaMemcpyAsync(d_idx, idx, (size_t)e->n * 4, cudaMemcpyHostToDevice, e->stream));
    if (mask) CU(cudaMemcpyAsync(d_mask, mask, (size_t)e->n, cudaMemcpyHostToDevice, e->stream));
    if (mode == TPL_RESET_ALL) CU(cudaMemsetAsync(e->episode, 0, (size_t)e->n * 4, e->stream));
    RC(tpl_reset_from_pool(e->state, e->stride, e->n, e->pool, e->K, idx ? d_idx : nullptr, mask ? d_mask : nullptr, mode,
                           e->episode, e->d_tstep, e->seed, e->env_base, gen_count, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

static int ensure_move_bufs(tpl_env *e) {
    const size_t n = (size_t)e->n;
    RC(ensure((void **)&e->d_rot, n)); RC(ensure((void **)&e->d_loc, n)); RC(ensure((void **)&e->d_flags, n));
    RC(ensure((void **)&e->d_dlines, n)); RC(ensure((void **)&e->d_st, n));
    return 0;
}

int tpl_env_move(tpl_env *e, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags, int8_t *st) {
    if (!e || !rot || !loc) return fail(TPL_EINVAL, "tpl_env_move: null argument");
    CU(cudaSetDevice(e->device));
    RC(ensure_move_bufs(e));
    const size_t n = (size_t)e->n;
    CU(cudaMemcpyAsync(e->d_rot, rot, n, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_loc, loc, n, cudaMemcpyHostToDevice, e->stream));
    RC(tpl_step(e->state, e->stride, e->n, e->d_rot, e->d_loc, e->d_dlines, e->d_flags, e->d_st, nullptr, e->L, e->M, e->stream));
    if (dlines) CU(cudaMemcpyAsync(dlines, e->d_dlines, n, cudaMemcpyDeviceToHost, e->stream));
    if (flags) CU(cudaMemcpyAsync(flags, e->d_flags, n, cudaMemcpyDeviceToHost, e->stream));
    if (st) CU(cudaMemcpyAsync(st, e->d_st, n, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

int tpl_env_get_state(tpl_env *e, uint16_t *rows, uint8_t *cur, uint8_t *next, int32_t *lines, int32_t *moves, int8_t *st,
                      uint8_t *head, uint8_t *npieces, uint8_t *queue) {
    if (!e) return fail(TPL_EINVAL, "tpl_env_get_state: null handle");
    CU(cudaSetDevice(e->device));
    const size_t n = (size_t)e->n;
    if (rows) RC(ensure((void **)&e->d_rows, n * 40));
    if (cur) RC(ensure((void **)&e->d_cur, n));
    if (next) RC(ensure((void **)&e->d_next, n));
    if (lines) RC(ensure((void **)&e->d_lines, n * 4));
    if (moves) RC(ensure((void **)&e->d_moves, n * 4));
    if (st) RC(ensure((void **)&e->d_st, n));
    if (head) RC(ensure((void **)&e->d_head, n));
    if (npieces) RC(ensure((void **)&e->d_np, n));
    if (queue) RC(ensure((void **)&e->d_queue, n * TPL_MAX_PIECES));
    RC(tpl_unpack(e->state, e->stride, e->n, rows ? e->d_rows : nullptr, cur ? e->d_cur : nullptr, next ? e->d_next : nullptr,
                  lines ? e->d_lines : nullptr, moves ? e->d_moves : nullptr, st ? e->d_st : nullptr, head ? e->d_head : nullptr,
                  npieces ? e->d_np : nullptr, queue ? e->d_queue : nullptr, e->stream));
    if (rows) CU(cudaMemcpyAsync(rows, e->d_rows, n * 40, cudaMemcpyDeviceToHost, e->stream));
    if (cur) CU(cudaMemcpyAsync(cur, e->d_cur, n, cudaMemcpyDeviceToHost, e->stream));
    if (next) CU(cudaMemcpyAsync(next, e->d_next, n, cudaMemcpyDeviceToHost, e->stream));
    if (lines) CU(cudaMemcpyAsync(lines, e->d_lines, n * 4, cudaMemcpyDeviceToHost, e->stream));
    if (moves) CU(cudaMemcpyAsync(moves, e->d_moves, n * 4, cudaMemcpyDeviceToHost, e->stream));
    if (st) CU(cudaMemcpyAsync(st, e->d_st, n, cudaMemcpyDeviceToHost, e->stream));
    if (head) CU(cudaMemcpyAsync(head, e->d_head, n, cudaMemcpyDeviceToHost, e->stream));
    if (npieces) CU(cudaMemcpyAsync(npieces, e->d_np, n, cudaMemcpyDeviceToHost, e->stream));
    if (queue) CU(cudaMemcpyAsync(queue, e->d_queue, n * TPL_MAX_PIECES, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

static int afterstates_to_host(tpl_env *e, uint8_t *feats, uint8_t *aflags) {
    const size_t n = (size_t)e->n;
    if (feats) RC(ensure_feats(e, n * 160));
    if (aflags) RC(ensure((void **)&e->d_aflags, n * 40));
    RC(tpl_afterstates(e->state, e->stride, e->n, feats ? e->d_feats : nullptr, aflags ? e->d_aflags : nullptr, nullptr, e->L, e->M,
                       e->stream));
    if (feats) CU(cudaMemcpyAsync(feats, e->d_feats, n * 160, cudaMemcpyDeviceToHost, e->stream));
    if (aflags) CU(cudaMemcpyAsync(aflags, e->d_aflags, n * 40, cudaMemcpyDeviceToHost, e->stream));
    return 0;
}

int tpl_env_afterstates(tpl_env *e, uint8_t *feats, uint8_t *flags) {
    if (!e || (!feats && !flags)) return fail(TPL_EINVAL, "tpl_env_afterstates: no output requested");
    CU(cudaSetDevice(e->device));
    RC(afterstates_to_host(e, feats, flags));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

// One rollout step through host buffers, pipelined over env chunks: on `stream`, for each chunk, H2D of its actions and the
// fused kernel; on `copy_stream`, as soon as a chunk's kernel is done, D2H of its results -- so the transfer of chunk c
// overlaps the kernel (and the action upload) of chunk c + 1, and the PCIe link never waits for the whole batch.
// 40-slot forms: a chunk's kernel writes chunk-local slot-major arrays [40][nc]; one strided 2-D copy per array puts them at
// the chunk's columns of the caller's [40][n] host arrays.
int tpl_env_step_observe(tpl_env *e, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags, int8_t *st,
                         uint8_t *feats, uint8_t *aflags) {
    if (!e || !rot || !loc) return fail(TPL_EINVAL, "tpl_env_step_observe: null argument");
    if (!e->pool) return fail(TPL_EINVAL, "tpl_env_step_observe: no config pool (call tpl_env_set_pool first)");
    if (!feats && aflags) return fail(TPL_EINVAL, "tpl_env_step_observe: aflags without feats");
    CU(cudaSetDevice(e->device));
    RC(ensure_move_bufs(e));
    const size_t n = (size_t)e->n;
    RC(ensure_feats(e, n * 160));
    if (aflags) RC(ensure((void **)&e->d_aflags, n * 40));
    // (features left on the device: one chunk, so that tpl_env_feats_ptr is one [40][n] array -- nothing big to overlap anyway)
    const int nchunks = feats ? e->nchunks : 1;
    const size_t chunk_envs = feats ? (size_t)e->chunk_envs : n;
    for (int c = 0; c < nchunks; ++c) {
        const size_t c0 = (size_t)c * chunk_envs, nc = (c0 + chunk_envs <= n) ? chunk_envs : n - c0;
        CU(cudaMemcpyAsync(e->d_rot + c0, rot + c0, nc, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(e->d_loc + c0, loc + c0, nc, cudaMemcpyHostToDevice, e->stream));
        uint8_t *cf = e->d_feats + c0 * 160, *ca = aflags ? e->d_aflags + c0 * 40 : nullptr;
        RC(tpl_step_observe((uint4 *)e->state + c0, e->stride, (int)nc, e->d_rot + c0, e->d_loc + c0, e->d_dlines + c0, e->d_flags + c0,
                            e->d_st + c0, nullptr, e->pool, e->K, e->episode + c0, e->d_tstep + c0, e->seed, e->env_base + c0, 0, cf, ca,
                            nullptr, e->L, e->M, e->stream));
        CU(cudaEventRecord(e->ev_done[c], e->stream));
        CU(cudaStreamWaitEvent(e->copy_stream, e->ev_done[c], 0));
        if (dlines) CU(cudaMemcpyAsync(dlines + c0, e->d_dlines + c0, nc, cudaMemcpyDeviceToHost, e->copy_stream));
        if (flags) CU(cudaMemcpyAsync(flags + c0, e->d_flags + c0, nc, cudaMemcpyDeviceToHost, e->copy_stream));
        if (st) CU(cudaMemcpyAsync(st + c0, e->d_st + c0, nc, cudaMemcpyDeviceToHost, e->copy_stream));
        // feats == NULL: the 40-slot features stay in HBM (compact form, tpl_env_feats_ptr) for a policy that runs on the GPU;
        // only the move's results travel back
        if (feats) CU(cudaMemcpy2DAsync(feats + c0 * 4, n * 4, cf, nc * 4, nc * 4, 40, cudaMemcpyDeviceToHost, e->copy_stream));
        if (aflags) CU(cudaMemcpy2DAsync(aflags + c0, n, ca, nc, nc, 40, cudaMemcpyDeviceToHost, e->copy_stream));
    }
    CU(cudaStreamSynchronize(e->copy_stream));
    return 0;
}

// The same step in the distinct-placements form (tpl_step_observe_distinct): only the placements that differ cross PCIe.
// rows (host, >= TPL_DISTINCT_CAPACITY(n) + 4 * chunks words): chunk c writes into its own region of the array; runs (host,
// u32[n]) says where each env's run starts in it.  How many words a chunk produced is only known after its kernel, so the
// host reads the chunk's counter (4 bytes, behind the kernel on `stream`) and then queues exactly that many words on
// `copy_stream` -- while the later chunks' kernels are already running.  *words_copied = total rows words transferred.
int tpl_env_step_observe_distinct(tpl_env *e, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags, int8_t *st,
                                  uint32_t *rows, int64_t rows_capacity, uint32_t *runs, int64_t *words_copied) {
    if (!e || !rot || !loc || !rows || !runs) return fail(TPL_EINVAL, "tpl_env_step_observe_distinct: null argument");
    if (!e->pool) return fail(TPL_EINVAL, "tpl_env_step_observe_distinct: no config pool (call tpl_env_set_pool first)");
    const int64_t region = TPL_DISTINCT_CAPACITY(e->chunk_envs);                 // words per chunk region (device and host)
    if (rows_capacity < region * e->nchunks)
        return fail(TPL_ERANGE, "tpl_env_step_observe_distinct: rows_capacity %lld < %lld words (tpl_env_distinct_capacity)",
                    (long long)rows_capacity, (long long)(region * e->nchunks));
    CU(cudaSetDevice(e->device));
    RC(ensure_move_bufs(e));
    const size_t n = (size_t)e->n;
    RC(ensure((void **)&e->d_drows, (size_t)region * e->nchunks * 4));
    RC(ensure((void **)&e->d_runs, n * 4));
    if (!e->d_cursor2) {
        RC(ensure((void **)&e->d_cursor2, 2 * TPL_MAX_CHUNKS * 4));
        CU(cudaMemsetAsync(e->d_cursor2, 0, 2 * TPL_MAX_CHUNKS * 4, e->stream));
        CU(cudaHostAlloc((void **)&e->h_cursor, TPL_MAX_CHUNKS * 4, cudaHostAllocDefault));
    }
    const int phase = e->phase; e->phase ^= 1;
    for (int c = 0; c < e->nchunks; ++c) {
        const size_t c0 = (size_t)c * e->chunk_envs, nc = (c0 + e->chunk_envs <= n) ? (size_t)e->chunk_envs : n - c0;
        CU(cudaMemcpyAsync(e->d_rot + c0, rot + c0, nc, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(e->d_loc + c0, loc + c0, nc, cudaMemcpyHostToDevice, e->stream));
        RC(tpl_step_observe_distinct((uint4 *)e->state + c0, e->stride, (int)nc, e->d_rot + c0, e->d_loc + c0, e->d_dlines + c0,
                                     e->d_flags + c0, e->d_st + c0, nullptr, e->pool, e->K, e->episode + c0, e->d_tstep + c0, e->seed,
                                     e->env_base + c0, 0, e->d_drows + region * c, region, e->d_runs + c0, (uint32_t)(region * c),
                                     e->d_cursor2 + 2 * c, phase, e->L, e->M, e->stream));
        CU(cudaMemcpyAsync(e->h_cursor + c, e->d_cursor2 + 2 * c + phase, 4, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaEventRecord(e->ev_done[c], e->stream));
    }
    int64_t total = 0;
    for (int c = 0; c < e->nchunks; ++c) {
        const size_t c0 = (size_t)c * e->chunk_envs, nc = (c0 + e->chunk_envs <= n) ? (size_t)e->chunk_envs : n - c0;
        CU(cudaEventSynchronize(e->ev_done[c]));
        const size_t words = e->h_cursor[c];
        if ((int64_t)words > region) return fail(TPL_ERANGE, "tpl_env_step_observe_distinct: chunk %d produced %zu words > region", c, words);
        total += (int64_t)words;
        CU(cudaMemcpyAsync(rows + region * c, e->d_drows + region * c, words * 4, cudaMemcpyDeviceToHost, e->copy_stream));
        CU(cudaMemcpyAsync(runs + c0, e->d_runs + c0, nc * 4, cudaMemcpyDeviceToHost, e->copy_stream));
        if (dlines) CU(cudaMemcpyAsync(dlines + c0, e->d_dlines + c0, nc, cudaMemcpyDeviceToHost, e->copy_stream));
        if (flags) CU(cudaMemcpyAsync(flags + c0, e->d_flags + c0, nc, cudaMemcpyDeviceToHost, e->copy_stream));
        if (st) CU(cudaMemcpyAsync(st + c0, e->d_st + c0, nc, cudaMemcpyDeviceToHost, e->copy_stream));
    }
    CU(cudaStreamSynchronize(e->copy_stream));
    if (words_copied) *words_copied = total;
    return 0;
}

int64_t tpl_env_distinct_capacity(tpl_env *e) { return e ? TPL_DISTINCT_CAPACITY(e->chunk_envs) * e->nchunks : 0; }
int tpl_env_chunks(tpl_env *e) { return e ? e->nchunks : 0; }

void *tpl_env_feats_ptr(tpl_env *e) { return e ? (void *)e->d_feats : nullptr; }

}  // extern "C"
