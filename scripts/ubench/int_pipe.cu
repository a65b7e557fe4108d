// Integer-pipe microbenchmark for B200 (sm_100a): warp-instruction throughput of the integer ops the
// Tetris-piclim kernels are made of.  8 independent dependency chains per thread, 1024 threads per CTA,
// one CTA per SM.  Prints lane-ops / clk / SM for each op (peak INT32 ALU = 64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe int_pipe.cu && ./int_pipe
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c) {
    if (OP == 0) return (a & b) | c;                                  // LOP3
    if (OP == 1) return a + b + c;                                    // IADD3
    if (OP == 2) return a << (b & 31);                                // SHF (variable)
    if (OP == 3) return a * b + c;                                    // IMAD
    if (OP == 4) return __sad((int)a, (int)b, c);                     // VABSDIFF
    if (OP == 5) return __vsadu4(a, b) + c;                           // VABSDIFF4.U8.ACC
    if (OP == 6) return (unsigned)__viaddmax_s32((int)a, (int)b, (int)c);   // VIADDMNMX
    if (OP == 7) return __viaddmax_s16x2(a, b, c);                    // VIADDMNMX.S16x2
    if (OP == 8) return __byte_perm(a, b, c);                         // PRMT (variable selector)
    if (OP == 9) return __popc(a) + b;                                // POPC (+ add)
    if (OP == 10) return __clz(a) + b;                                // FLO (+ add)
    if (OP == 11) return (unsigned)abs((int)a - (int)b);              // IADD + IABS
    if (OP == 12) return a > b ? c : a;                               // ISETP + SEL
    if (OP == 13) return (unsigned)max((int)a, (int)b);               // VIMNMX
    if (OP == 14) return __vabsdiffu4(a, b);                          // VABSDIFF4.U8
    if (OP == 15) return __byte_perm(a, b, 0x6420);                   // PRMT (immediate selector)
    if (OP == 16) return __funnelshift_r(a, b, c);                    // SHF.R funnel
    return a;
}

template <int OP>
__global__ void __launch_bounds__(1024) bench(unsigned *out, unsigned seed, long long *cycles) {
    unsigned v[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) v[k] = seed + threadIdx.x * 17 + k * 101;
    unsigned b = seed ^ 0x5bd1e995u, c = seed * 3 + 7;
    long long t0 = clock64();
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) v[k] = op<OP>(v[k], b, c);
        b += 0;   // keep b, c loop-invariant registers
    }
    long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s ^= v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, unsigned *out, long long *cyc) {
    int sms = 148;
    bench<OP><<<sms, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<sms, 1024>>>(out, 12345u, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; ++i) c += (double)h[i]; c /= sms;
    double laneops = 1024.0 * ITERS * CHAINS;
    printf("%-28s %7.2f lane-ops/clk/SM   (%.0f cycles, %.3f ms, %.2f Tops/s chip)\n", name, laneops / c, c, ms,
           laneops * sms / (ms * 1e-3) / 1e12);
}

int main() {
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("LOP3", out, cyc); run<1>("IADD3", out, cyc); run<2>("SHF.L var", out, cyc); run<3>("IMAD", out, cyc);
    run<4>("VABSDIFF (sad)", out, cyc); run<5>("VABSDIFF4.U8.ACC (+IADD)", out, cyc); run<6>("VIADDMNMX", out, cyc);
    run<7>("VIADDMNMX.S16x2", out, cyc); run<8>("PRMT var", out, cyc); run<9>("POPC (+IADD)", out, cyc);
    run<10>("FLO/clz (+IADD)", out, cyc); run<11>("IADD+IABS", out, cyc); run<12>("ISETP+SEL", out, cyc);
    run<13>("VIMNMX", out, cyc); run<14>("VABSDIFF4.U8", out, cyc); run<15>("PRMT imm", out, cyc); run<16>("SHF funnel", out, cyc);
    return 0;
}
