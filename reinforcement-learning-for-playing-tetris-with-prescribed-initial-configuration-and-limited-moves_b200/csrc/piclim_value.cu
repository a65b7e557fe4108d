// piclim_value.cu -- inference-only value net for RANKING afterstates, one fused kernel on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in tensor memory), plus the per-env action selection that consumes it.
//
// The reference's network is five Linear layers, in -> 128 -> 128 -> 128 -> 128 -> out with ReLU between them
// (model/model.py:9-20); as the afterstate value net it is 4 -> 128 -> 128 -> 128 -> 128 -> 1 on the four features the env
// kernel emits.  Training stays in PyTorch (north star); what a rollout needs every step is only the forward pass over every
// distinct placement of every env -- 23 N rows -- and PyTorch runs that as five library GEMMs that write the [rows, 128]
// activations to HBM four times.  Here a CTA keeps all weights in shared memory and the activations on chip:
//   rows (u32 feature words, the distinct-placements form)  ->  bf16 A tile [128 x 16] in shared memory (features * scale, a
//   constant 1 that carries the first bias)  ->  UMMA 128x128x16  ->  TMEM  ->  registers: ReLU + bf16 in one cvt per two
//   elements  ->  shared memory  ->  UMMA 128x128x144 (three times; the ninth K-step multiplies a constant-1 column with the
//   bias row, so the epilogue has no bias add)  ->  UMMA 128x16x144 (the last layer, N padded)  ->  one TMEM load per row  ->
//   values f32[rows].
// bf16 operands, fp32 accumulation: the numerics of ValueNet.rank_bf16.  Two 128-row tiles are in flight per CTA (two groups of
// four warps, each with its own accumulator columns, activation buffer and mbarrier), so one group's epilogue overlaps the
// other's MMAs.  Operand tiles use the canonical K-major no-swizzle UMMA layout: 8-row x 16-byte core matrices, the two
// K-halves of an instruction 128 bytes apart (LBO), 8-row groups SBO apart.
#include "../../include/tetris_piclim.h"
#include "piclim_core.cuh"

#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace tpl {
int fail(int code, const char *fmt, ...);
int check_launch(const char *what);
}

namespace tplv {

constexpr int HID = 128;                         // hidden width (model/model.py:9-13)
constexpr int KAUG = HID + 16;                   // K of the hidden layers: 128 activations + one K-step whose first column is the constant 1
                                                 // that carries the bias (the other 15 are zero) -- no bias add in the epilogue
constexpr int NOUT = 16;                         // the last layer (128 -> 1) as an MMA too: N padded to the smallest UMMA N for M = 128
constexpr int TILE_ROWS = 128;                   // UMMA M
constexpr int GROUPS = 2;                        // tiles in flight per CTA
constexpr int VTHREADS = GROUPS * 128;
// ---- packed weight blob (device memory, copied to shared memory by every CTA); byte offsets
constexpr int OFF_W1 = 0;                        // [128 n][16 k] bf16: k < 4 weights, k == 4 bias, rest 0                       (4 KB)
constexpr int W_BYTES = HID * KAUG * 2;          // [128 n][144 k] bf16: k < 128 weights, k == 128 bias, rest 0                  (36 KB)
constexpr int OFF_W2 = 4096;                     // three of them
constexpr int OFF_W5 = OFF_W2 + 3 * W_BYTES;     // [16 n][144 k] bf16: row 0 = last layer's weights and bias, rows 1..15 zero  (4.5 KB)
constexpr int OFF_B = OFF_W5 + NOUT * KAUG * 2;  // f32: scale[4], pad[4]
constexpr int BLOB_BYTES = OFF_B + 32;
static_assert(BLOB_BYTES % 16 == 0 && BLOB_BYTES == TPL_VALUE_BLOB_BYTES, "blob size is part of the ABI");
// ---- shared memory
constexpr int OFF_A0 = BLOB_BYTES;               // per group: input tile [128 x 16] bf16 (4 KB), then activations [128 x 144] bf16 (36 KB)
constexpr int ACT_BYTES = TILE_ROWS * KAUG * 2;
constexpr int GROUP_BYTES = 4096 + ACT_BYTES;
constexpr int OFF_BAR = OFF_A0 + GROUPS * GROUP_BYTES;       // 2 mbarriers + the TMEM base address
constexpr int SMEM_BYTES = OFF_BAR + 32;
constexpr int TMEM_COLS = 512;                   // two groups at a 256-column pitch:
constexpr int TM_D = 0;                          //   128 accumulator columns
constexpr int TM_A = 128;                        //   the next layer's A operand: 144 bf16 per row = 72 columns (TPL_VALUE_A_TMEM)
constexpr int TM_OUT = 200;                      //   16 columns for the last layer
#ifndef TPL_VALUE_A_TMEM
#define TPL_VALUE_A_TMEM 1                       // activations go back into tensor memory (the MMA's A operand may live there), not through
#endif                                           // shared memory: the epilogue's 32 KB of shared-memory stores and the MMA's re-read of them drop out

// canonical K-major, no swizzle: element (row, k) of a [rows x K] bf16 tile
__host__ __device__ constexpr uint32_t canon_off(uint32_t row, uint32_t k, uint32_t K) {
    return (row >> 3) * (K * 16u) + (k >> 3) * 128u + (row & 7u) * 16u + (k & 7u) * 2u;
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// fp32 parameters (device pointers, PyTorch layout: weight [out][in]) -> blob
__global__ void pack_kernel(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3, const float *b3,
                            const float *w4, const float *b4, const float *w5, const float *b5, float s0, float s1, float s2, float s3,
                            uint8_t *blob) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    __nv_bfloat16 *W1 = reinterpret_cast<__nv_bfloat16 *>(blob + OFF_W1);
    for (int e = tid; e < HID * 16; e += nth) {
        const int n = e >> 4, k = e & 15;
        const float v = k < 4 ? w1[n * 4 + k] : (k == 4 ? b1[n] : 0.0f);
        W1[canon_off(n, k, 16) / 2] = __float2bfloat16_rn(v);
    }
    const float *ws[3] = {w2, w3, w4}, *bs[3] = {b2, b3, b4};
    for (int l = 0; l < 3; ++l) {
        __nv_bfloat16 *W = reinterpret_cast<__nv_bfloat16 *>(blob + OFF_W2 + l * W_BYTES);
        for (int e = tid; e < HID * KAUG; e += nth) {
            const int n = e / KAUG, k = e - n * KAUG;
            const float v = k < HID ? ws[l][n * HID + k] : (k == HID ? bs[l][n] : 0.0f);
            W[canon_off(n, k, KAUG) / 2] = __float2bfloat16_rn(v);
        }
    }
    __nv_bfloat16 *W5 = reinterpret_cast<__nv_bfloat16 *>(blob + OFF_W5);
    for (int e = tid; e < NOUT * KAUG; e += nth) {
        const int n = e / KAUG, k = e - n * KAUG;
        const float v = n == 0 ? (k < HID ? w5[k] : (k == HID ? b5[0] : 0.0f)) : 0.0f;
        W5[canon_off(n, k, KAUG) / 2] = __float2bfloat16_rn(v);
    }
    float *B = reinterpret_cast<float *>(blob + OFF_B);
    if (tid == 0) { B[0] = s0; B[1] = s1; B[2] = s2; B[3] = s3; B[4] = B[5] = B[6] = B[7] = 0.0f; }
}

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers (the strings follow CUTLASS's cute/arch/*sm100* headers)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: start address, leading (K-half) and stride (8-row group) byte offsets in 16-byte units,
// descriptor version 1 (Blackwell), no swizzle
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t idesc_for(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_ROWS >> 4) << 24); }
constexpr uint32_t IDESC_HID = idesc_for(HID), IDESC_OUT = idesc_for(NOUT);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with the A operand in tensor memory (lane = row, two bf16 per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar_saddr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar_saddr), "r"(count) : "memory");
}
// Wait for the phase with the given parity.  A wait that lasts a second can only mean a lost arrival (a mis-built descriptor,
// a fault in the MMA): trap -- the launch then fails loudly instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar_saddr, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar_saddr), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 2000000000ll) __trap();
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" :: "r"(1 + g) : "memory"); }

// 32 consecutive accumulator columns of this thread's TMEM lane (warp-collective); issue only -- tmem_ld_wait() completes it
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
// All tensor-memory loads of this thread have landed.  The empty volatile statements pin every destination register behind the
// wait: the compiler orders volatile asm statements, so no use of v[] can be scheduled before the data is there.
template <int N>
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[N]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < N; ++j) asm volatile("" : "+r"(v[j]));
}
// registers -> 32 / 8 consecutive columns of this thread's TMEM lane (warp-collective); tmem_st_wait() completes them
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                    "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                    "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&p);
}
// max(x, 0) and the bf16 rounding of two accumulators in ONE instruction
__device__ __forceinline__ uint32_t relu_pack_bf16x2(uint32_t lo_f32, uint32_t hi_f32) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_f32)), "f"(__uint_as_float(lo_f32)));
    return d;
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;\n" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// values[r] = V(features of rows[r]) for r < *count (or < nrows when count == nullptr)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VTHREADS, 1)
value_rows_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ count, uint32_t nrows_max, const uint8_t *__restrict__ blob,
                  float *__restrict__ values) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, g = tid >> 7, tg = tid & 127, wq = (tid >> 5) & 3;
    {   // weights: one coalesced copy per CTA (L2-resident after the first CTA)
        const uint4 *src = reinterpret_cast<const uint4 *>(blob);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int k = tid; k < BLOB_BYTES / 16; k += VTHREADS) dst[k] = src[k];
    }
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 16);
    const uint32_t bar = smem_u32(smem + OFF_BAR + 8 * g);
    if (tid == 0) {
        mbar_init(smem_u32(smem + OFF_BAR), 1); mbar_init(smem_u32(smem + OFF_BAR + 8), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {                               // warp 0 owns the tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    uint8_t *a0 = smem + OFF_A0 + g * GROUP_BYTES, *act = a0 + 4096;
    const uint32_t my_a0 = canon_off((uint32_t)tg, 0, 16), my_act = canon_off((uint32_t)tg, 0, KAUG);
#if !TPL_VALUE_A_TMEM
    // the bias K-step of this row of the activation tile, written once: (1, 0, ..., 0)
    *reinterpret_cast<uint4 *>(act + my_act + (HID / 8) * 128u) = make_uint4(0x00003F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(act + my_act + (HID / 8 + 1) * 128u) = make_uint4(0u, 0u, 0u, 0u);
#endif
    proxy_fence();                                // weights (and bias columns) were written through the generic proxy, the MMAs read through the async one
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot + (uint32_t)g * 256u;              // this group's tensor-memory columns (see TM_*)
    const uint32_t tmem_mine = tmem_d + ((uint32_t)(wq * 32) << 16);      // the 32 lanes this warp may read and write
#if TPL_VALUE_A_TMEM
    {   // the bias K-step of this row of the A operand, written once: (1, 0, ..., 0) as bf16 pairs
        const uint32_t ones[8] = {0x00003F80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(tmem_mine + TM_A + HID / 2, ones);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
#endif
    const float *B = reinterpret_cast<const float *>(smem + OFF_B);
    const float sc0 = B[0], sc1 = B[1], sc2 = B[2], sc3 = B[3];
    const uint32_t nrows = count ? min(*count, nrows_max) : nrows_max;
    const uint32_t ntiles = (nrows + TILE_ROWS - 1) / TILE_ROWS;
    uint32_t parity = 0;
    const uint32_t abase = smem_u32(act);
#ifdef TPL_VALUE_TRACE          // developer aid: cycle stamps of CTA 0 / group 0 / thread 0 into the unused tail of `values`
    float *trace = values + (nrows_max - 4096u);
    int tr = 0;
    long long tlast = clock64();
#define TPL_TR(tag) if (blockIdx.x == 0 && tid == 0 && tr < 4000) { const long long now_ = clock64(); trace[tr++] = (float)(tag); trace[tr++] = (float)(now_ - tlast); tlast = now_; }
#else
#define TPL_TR(tag)
#endif

    const uint32_t tile0 = blockIdx.x * GROUPS + g, tstep = gridDim.x * GROUPS;
    // this thread's feature word of the NEXT tile is loaded a tile ahead: the DRAM round trip is off the per-tile critical path
    uint32_t w_next = (tile0 < ntiles && tile0 * TILE_ROWS + (uint32_t)tg < nrows) ? rows[tile0 * TILE_ROWS + (uint32_t)tg] : 0u;
    for (uint32_t tile = tile0; tile < ntiles; tile += tstep) {
        const uint32_t row = tile * TILE_ROWS + (uint32_t)tg;
        {   // input tile: (rows cleared, holes, bumpiness, aggregate height) * scale, then the constant 1 of the first bias
            const uint32_t w = w_next;
            const uint32_t row2 = row + tstep * TILE_ROWS;
            w_next = (tile + tstep < ntiles && row2 < nrows) ? rows[row2] : 0u;
            const float f0 = (float)(w & 7u) * sc0, f1 = (float)((w >> 8) & 0xFFu) * sc1, f2 = (float)((w >> 16) & 0xFFu) * sc2,
                        f3 = (float)(w >> 24) * sc3;
            *reinterpret_cast<uint4 *>(a0 + my_a0) = make_uint4(pack_bf16x2(f0, f1), pack_bf16x2(f2, f3), pack_bf16x2(1.0f, 0.0f), 0u);
            *reinterpret_cast<uint4 *>(a0 + my_a0 + 128) = make_uint4(0u, 0u, 0u, 0u);
        }
        proxy_fence();
        group_sync(g);
        if (tg == 0) {                            // layer 1: one 128 x 128 x 16 instruction
            tc_fence_after();
            umma_f16(tmem_d + TM_D, umma_desc(smem_u32(a0), 128, 256), umma_desc(smem_u32(smem + OFF_W1), 128, 256), IDESC_HID, 0u);
            umma_commit(bar);
        }
#pragma unroll 1
        for (int layer = 1; layer <= 4; ++layer) {
            TPL_TR(1)
            mbar_wait(bar, parity); parity ^= 1u;
            TPL_TR(2)
            tc_fence_after();
            // epilogue of a hidden layer: accumulator -> ReLU -> bf16 -> the next MMA's A operand (canonical layout: 8 columns =
            // one 16-byte chunk); the bias came in through the constant-1 column, so there is one instruction per two elements
            {
                uint32_t v[HID];                  // the whole row of the accumulator: four loads in flight, ONE wait
#pragma unroll
                for (int c0 = 0; c0 < HID; c0 += 32) tmem_ld32_issue(tmem_mine + TM_D + (uint32_t)c0, v + c0);
                tmem_ld_wait(v);
                TPL_TR(3)
#if TPL_VALUE_A_TMEM
                uint32_t a[HID / 2];              // ReLU + bf16, two activations per 32-bit tensor-memory column
#pragma unroll
                for (int q = 0; q < HID / 2; ++q) a[q] = relu_pack_bf16x2(v[2 * q], v[2 * q + 1]);
                tmem_st32(tmem_mine + TM_A, a);
                tmem_st32(tmem_mine + TM_A + 32, a + 32);
                tmem_st_wait();
                TPL_TR(4)
#else
#pragma unroll
                for (int q = 0; q < HID / 8; ++q)
                    *reinterpret_cast<uint4 *>(act + my_act + (uint32_t)q * 128u) =
                        make_uint4(relu_pack_bf16x2(v[8 * q], v[8 * q + 1]), relu_pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                   relu_pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), relu_pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
                proxy_fence();                    // activations visible to the tensor core's shared-memory reads
#endif
            }
            tc_fence_before();                    // this thread's tensor-memory accesses are done before the next MMAs touch the same columns
            group_sync(g);
            TPL_TR(5)
            if (tg == 0) {
                tc_fence_after();
                // hidden layer: 128 x 128 x 144 = nine K-steps (the ninth adds the bias); last layer: 128 x 16 x 144 into its own columns
                const uint32_t wbase = smem_u32(smem + (layer < 4 ? OFF_W2 + (layer - 1) * W_BYTES : OFF_W5));
                const uint32_t dcol = tmem_d + (layer < 4 ? TM_D : TM_OUT), idesc = layer < 4 ? IDESC_HID : IDESC_OUT;
#pragma unroll
                for (uint32_t k = 0; k < KAUG / 16; ++k) {
#if TPL_VALUE_A_TMEM
                    umma_f16_ts(dcol, tmem_d + TM_A + k * 8u, umma_desc(wbase + k * 256u, 128, KAUG * 16), idesc, k);
#else
                    umma_f16(dcol, umma_desc(abase + k * 256u, 128, KAUG * 16), umma_desc(wbase + k * 256u, 128, KAUG * 16), idesc, k);
#endif
                }
                umma_commit(bar);
            }
            TPL_TR(6)
        }
        mbar_wait(bar, parity); parity ^= 1u;
        TPL_TR(7)
        tc_fence_after();
        const float value = __uint_as_float(tmem_ld1(tmem_mine + TM_OUT));       // column 0 of the last layer's accumulator
        if (row < nrows) values[row] = value;
        tc_fence_before();
        group_sync(g);                            // every lane has read its value before the next tile's MMAs run
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(*tmem_slot), "r"(TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// per-env action selection over the distinct placements: q = reward(placement) + gamma * V, arg-max with lowest-index tie-break,
// epsilon-greedy on the counter RNG (stream 3, keyed by (seed, env, step)); writes the action and the chosen placement's word
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint32_t STREAM_EXPLORE = 3;

__global__ void __launch_bounds__(128)
select_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ runs, const float *__restrict__ values, int n, float gamma,
              float r_win, float r_lose, float eps, uint64_t seed, uint64_t env_base, uint32_t step, uint8_t *rot, uint8_t *loc,
              uint32_t *chosen, float *chosen_q) {
    __shared__ uint4 s_tab[tpl::TAB_COMPACT4];
    for (int t = threadIdx.x; t < tpl::TAB_COMPACT4; t += blockDim.x) s_tab[t] = reinterpret_cast<const uint4 *>(&tpl::c_orient)[t];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t d = runs[i], piece = d >> 29, off = d & 0x1FFFFFFFu;
    uint32_t r_out = 0, c_out = 0, w_out = 0; float q_out = 0.0f;
    if (piece < 7u) {
        const uint32_t cnt = tpl::orient_run_len(s_tab[piece * 8 + 1]);
        int best = 0; float bestq = -3.0e38f;
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t w = rows[off + j], fl = (w & 0xFFu) >> 3;
            const float q = (float)(w & 7u) + ((fl & tpl::F_WIN) ? r_win : 0.0f) + ((fl & (tpl::F_LOSE | tpl::F_TOPOUT)) ? r_lose : 0.0f) +
                            gamma * values[off + j];
            if (q > bestq) { bestq = q; best = (int)j; }
        }
        const uint4 u = tpl::rng_words(seed, env_base + (uint64_t)i, step, STREAM_EXPLORE, 0);
        if ((float)(u.x >> 8) * (1.0f / 16777216.0f) < eps) best = (int)__umulhi(u.y, cnt);
        // placement index -> (rot, loc): rotations in order, rot_base[r] <= j < rot_base[r] + 11 - w
        for (uint32_t r = 0; r < 4; ++r) {
            const uint4 o = s_tab[(piece * 4 + r) * 2], ob = s_tab[(piece * 4 + r) * 2 + 1];
            const uint32_t rb = tpl::orient_rot_base(ob);
            if (rb != 0xFFu && (uint32_t)best >= rb && (uint32_t)best < rb + 11u - (uint32_t)tpl::orient_w(o)) { r_out = r; c_out = (uint32_t)best - rb; }
        }
        w_out = rows[off + (uint32_t)best];
        q_out = bestq;
    }
    rot[i] = (uint8_t)r_out; loc[i] = (uint8_t)c_out;
    if (chosen) chosen[i] = w_out;
    if (chosen_q) chosen_q[i] = q_out;
}

// ---------------------------------------------------------------------------------------------------------------------
// replay push: one transition per env straight from the step's outputs into the ring buffers of a DQN replay memory, in the
// form the TD target consumes without further element-wise work:
//   x        chosen placement's feature word, flags cleared (bytes = rows cleared, holes, bumpiness, aggregate height)
//   reward   rows cleared + r_win / r_lose by the move's flags;   live = 0 if the episode ended (or nothing is left to place), else 1
//   next_w   the new state's distinct placements, flags cleared, zero padded to DISTINCT_MAX
//   next_r   reward of each of them (-inf past the run's end, so a max over the row ignores the padding)
//   next_g   gamma where the placement does not end the episode, else 0:  Q(placement) = next_r + next_g * V(placement)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
replay_push_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ runs, const int8_t *__restrict__ dlines,
                   const uint8_t *__restrict__ mflags, const int8_t *__restrict__ state, const uint32_t *__restrict__ chosen, int n,
                   int64_t pos, int64_t capacity, float gamma, float r_win, float r_lose, uint32_t *x, float *reward, float *live,
                   uint32_t *next_w, float *next_r, float *next_g) {
    __shared__ uint4 s_tab[tpl::TAB_COMPACT4];
    for (int t = threadIdx.x; t < tpl::TAB_COMPACT4; t += blockDim.x) s_tab[t] = reinterpret_cast<const uint4 *>(&tpl::c_orient)[t];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t slot = (pos + i) % capacity;
    const uint32_t fl = mflags[i];
    x[slot] = chosen[i] & ~0xF8u;
    reward[slot] = (float)dlines[i] + ((fl & tpl::F_WIN) ? r_win : 0.0f) + ((fl & (tpl::F_LOSE | tpl::F_TOPOUT)) ? r_lose : 0.0f);
    const uint32_t d = runs[i], piece = d >> 29;
    const uint32_t cnt = piece < 7u ? tpl::orient_run_len(s_tab[piece * 8 + 1]) : 0u;
    live[slot] = (state[i] != 0 || cnt == 0u) ? 0.0f : 1.0f;
    const uint32_t *run = rows + (d & 0x1FFFFFFFu);
    uint32_t *dw = next_w + slot * tpl::DISTINCT_MAX;
    float *dr = next_r + slot * tpl::DISTINCT_MAX, *dg = next_g + slot * tpl::DISTINCT_MAX;
    for (uint32_t j = 0; j < (uint32_t)tpl::DISTINCT_MAX; ++j) {
        uint32_t w = 0u; float r = (j == 0u && cnt == 0u) ? 0.0f : -INFINITY, g = 0.0f;
        if (j < cnt) {
            w = run[j];
            const uint32_t f = (w & 0xFFu) >> 3;
            r = (float)(w & 7u) + ((f & tpl::F_WIN) ? r_win : 0.0f) + ((f & (tpl::F_LOSE | tpl::F_TOPOUT)) ? r_lose : 0.0f);
            g = (f & (tpl::F_WIN | tpl::F_LOSE | tpl::F_TOPOUT)) ? 0.0f : gamma;
            w &= ~0xF8u;
        }
        dw[j] = w; dr[j] = r; dg[j] = g;
    }
}

}  // namespace tplv

using namespace tplv;

extern "C" {

int tpl_value_pack(const float *w1, const float *b1, const float *w2, const float *b2, const float *w3, const float *b3, const float *w4,
                   const float *b4, const float *w5, const float *b5, const float *scale4_host, void *blob, void *stream) {
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w4 || !b4 || !w5 || !b5 || !scale4_host || !blob)
        return tpl::fail(TPL_EINVAL, "tpl_value_pack: null argument");
    if (((uintptr_t)blob & 15u) != 0) return tpl::fail(TPL_EINVAL, "tpl_value_pack: blob must be 16-byte aligned");
    pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(w1, b1, w2, b2, w3, b3, w4, b4, w5, b5, scale4_host[0], scale4_host[1], scale4_host[2],
                                                      scale4_host[3], (uint8_t *)blob);
    return tpl::check_launch("tpl_value_pack");
}

int tpl_value_rows(const uint32_t *rows, const uint32_t *count, int64_t nrows_max, const void *blob, float *values, void *stream) {
    if (!rows || !blob || !values || nrows_max < 0) return tpl::fail(TPL_EINVAL, "tpl_value_rows: null argument");
    if (nrows_max > 0xFFFFFF00ll) return tpl::fail(TPL_ERANGE, "tpl_value_rows: at most 2^32 - 256 rows per call");
    if (nrows_max == 0) return 0;
    static bool attr[64] = {};
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(value_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return tpl::fail((int)e, "tpl_value_rows: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr[dev] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (nrows_max + TILE_ROWS - 1) / TILE_ROWS;
    const int64_t want = (tiles + GROUPS - 1) / GROUPS;
    const unsigned grid = (unsigned)(want < sms ? want : sms);                    // persistent: one CTA per SM, two tiles in flight each
    value_rows_kernel<<<grid, VTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(rows, count, (uint32_t)nrows_max, (const uint8_t *)blob, values);
    return tpl::check_launch("tpl_value_rows");
}

int tpl_select_action(const uint32_t *rows, const uint32_t *runs, const float *values, int n, float gamma, float reward_win,
                      float reward_lose, float eps, uint64_t seed, uint64_t env_base, uint32_t step, uint8_t *rot, uint8_t *loc,
                      uint32_t *chosen, float *chosen_q, void *stream) {
    if (!rows || !runs || !values || !rot || !loc || n < 0) return tpl::fail(TPL_EINVAL, "tpl_select_action: null argument");
    if (n == 0) return 0;
    select_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(rows, runs, values, n, gamma, reward_win, reward_lose, eps, seed,
                                                                                 env_base, step, rot, loc, chosen, chosen_q);
    return tpl::check_launch("tpl_select_action");
}

int tpl_replay_push(const uint32_t *rows, const uint32_t *runs, const int8_t *dlines, const uint8_t *mflags, const int8_t *state,
                    const uint32_t *chosen, int n, int64_t pos, int64_t capacity, float gamma, float reward_win, float reward_lose,
                    uint32_t *x, float *reward, float *live, uint32_t *next_w, float *next_r, float *next_g, void *stream) {
    if (!rows || !runs || !dlines || !mflags || !state || !chosen || !x || !reward || !live || !next_w || !next_r || !next_g)
        return tpl::fail(TPL_EINVAL, "tpl_replay_push: null argument");
    if (n < 0 || capacity <= 0 || pos < 0 || n > capacity) return tpl::fail(TPL_ERANGE, "tpl_replay_push: need 0 <= n <= capacity, pos >= 0");
    if (n == 0) return 0;
    replay_push_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(rows, runs, dlines, mflags, state, chosen, n, pos, capacity,
                                                                                      gamma, reward_win, reward_lose, x, reward, live, next_w,
                                                                                      next_r, next_g);
    return tpl::check_launch("tpl_replay_push");
}

}  // extern "C"
