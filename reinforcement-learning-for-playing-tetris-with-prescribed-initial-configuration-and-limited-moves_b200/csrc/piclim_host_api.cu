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
    cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaMalloc(&e->state, (size_t)e->stride * 64);
    if (err == cudaSuccess) err = cudaMemsetAsync(e->state, 0, (size_t)e->stride * 64, e->stream);
    if (err == cudaSuccess) err = cudaMalloc((void **)&e->episode, (size_t)n * 4);
    if (err == cudaSuccess) err = cudaMemsetAsync(e->episode, 0, (size_t)n * 4, e->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) { int r = fail((int)err, "tpl_env_create: %s", cudaGetErrorString(err)); tpl_env_destroy(e); return r; }
    *out = e;
    return 0;
}

void tpl_env_destroy(tpl_env *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    void *ptrs[] = {e->state, e->pool, e->episode, e->d_rot, e->d_loc, e->d_flags, e->d_dlines, e->d_st, e->d_feats, e->d_aflags,
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
    for (int i = 0; i < m; ++i) if (npieces[i] > TPL_MAX_PIECES) return fail(TPL_ERANGE, "config %d has %d pieces (> 42)", i, npieces[i]);
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
                           e->episode, e->seed, e->env_base, gen_count, e->stream));
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

int tpl_env_step_observe(tpl_env *e, const uint8_t *rot, const uint8_t *loc, int8_t *dlines, uint8_t *flags, int8_t *st,
                         uint8_t *feats, uint8_t *aflags) {
    if (!e || !rot || !loc) return fail(TPL_EINVAL, "tpl_env_step_observe: null argument");
    if (!e->pool) return fail(TPL_EINVAL, "tpl_env_step_observe: no config pool (call tpl_env_set_pool first)");
    CU(cudaSetDevice(e->device));
    RC(ensure_move_bufs(e));
    const size_t n = (size_t)e->n;
    CU(cudaMemcpyAsync(e->d_rot, rot, n, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->d_loc, loc, n, cudaMemcpyHostToDevice, e->stream));
    if (!feats && aflags) return fail(TPL_EINVAL, "tpl_env_step_observe: aflags without feats");
    RC(ensure_feats(e, n * 160));
    if (aflags) RC(ensure((void **)&e->d_aflags, n * 40));
    RC(tpl_step_observe(e->state, e->stride, e->n, e->d_rot, e->d_loc, e->d_dlines, e->d_flags, e->d_st, nullptr, e->pool, e->K,
                        e->episode, e->seed, e->env_base, 0, e->d_feats, aflags ? e->d_aflags : nullptr, nullptr, e->L, e->M, e->stream));
    if (dlines) CU(cudaMemcpyAsync(dlines, e->d_dlines, n, cudaMemcpyDeviceToHost, e->stream));
    if (flags) CU(cudaMemcpyAsync(flags, e->d_flags, n, cudaMemcpyDeviceToHost, e->stream));
    if (st) CU(cudaMemcpyAsync(st, e->d_st, n, cudaMemcpyDeviceToHost, e->stream));
    // feats == NULL: the 40-slot features stay in HBM (compact form, tpl_env_feats_ptr) for a policy that runs on the GPU;
    // only the move's results travel back
    if (feats) CU(cudaMemcpyAsync(feats, e->d_feats, n * 160, cudaMemcpyDeviceToHost, e->stream));
    if (aflags) CU(cudaMemcpyAsync(aflags, e->d_aflags, n * 40, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return 0;
}

void *tpl_env_feats_ptr(tpl_env *e) { return e ? (void *)e->d_feats : nullptr; }

}  // extern "C"
