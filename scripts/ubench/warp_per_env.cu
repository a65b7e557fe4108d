// warp_per_env.cu -- a number behind one design decision ("Why bit-columns ... one thread per env").
//
// The survey's first proposal for this path was: boards as 20 x uint16 bitrows, ONE WARP PER ENV, lanes enumerating the placements,
// __ballot / __shfl for column heights.  This microbenchmark runs the core of the enumeration -- the hard-drop row of all 40
// (rotation, column) slots of the current piece (game/tetris.py:424-433) -- both ways on the same random boards and compares
// their results and their time:
//   A  "warp per env":   lane r < 20 loads bitrow r; ten ballots transpose the board into bit-columns (every lane gets all ten);
//                        lane k < 10 turns column k into a height; lane s computes slot s (and slot s + 32 in a second round)
//                        from four shuffled heights.
//   B  "thread per env": the layout the library uses -- ten bit-columns per env in one thread's registers, ten bit-scans, then the
//                        40 slots from registers (the same table, the same max).
// Both write sum over the 40 slots of (y + 1) per env; the program checks that A == B for every env and prints ns per env.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_per_env warp_per_env.cu && ./warp_per_env
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

// bottom offsets bo[piece][rot][j] (rows between the shape's bottom row and the lowest cell of column j; 64 = column not covered)
// and widths, from the row masks at game/tetris.py:25-55 (I L J T S Z O; rot >= n_rot repeats rot % n_rot, :61)
__constant__ int8_t c_bo[7][4][4] = {
    {{0, 0, 0, 0}, {0, 64, 64, 64}, {0, 0, 0, 0}, {0, 64, 64, 64}},
    {{0, 0, 0, 64}, {2, 0, 64, 64}, {0, 1, 1, 64}, {0, 0, 64, 64}},       // L: 4/7 ; 3,2,2 ; 7,1 ; 1,1,3
    {{0, 0, 0, 64}, {0, 0, 64, 64}, {1, 1, 0, 64}, {0, 2, 64, 64}},       // J: 1/7 ; 2,2,3 ; 7,4 ; 3,1,1
    {{0, 0, 0, 64}, {1, 0, 64, 64}, {1, 0, 1, 64}, {0, 1, 64, 64}},       // T: 2/7 ; 2,3,2 ; 7,2 ; 1,3,1
    {{0, 0, 1, 64}, {1, 0, 64, 64}, {0, 0, 1, 64}, {1, 0, 64, 64}},       // S: 6/3 ; 1,3,2
    {{1, 0, 0, 64}, {0, 1, 64, 64}, {1, 0, 0, 64}, {0, 1, 64, 64}},       // Z: 3/6 ; 2,3,1
    {{0, 0, 64, 64}, {0, 0, 64, 64}, {0, 0, 64, 64}, {0, 0, 64, 64}},     // O
};
__constant__ int8_t c_w[7][4] = {{4, 1, 4, 1}, {3, 2, 3, 2}, {3, 2, 3, 2}, {3, 2, 3, 2}, {3, 2, 3, 2}, {3, 2, 3, 2}, {2, 2, 2, 2}};

__device__ __forceinline__ int drop_y(int h0, int h1, int h2, int h3, int piece, int rot) {
    const int8_t *bo = c_bo[piece][rot];
    return max(max(h0 - bo[0], h1 - bo[1]), max(h2 - bo[2], h3 - bo[3]));
}

// A: one warp per env, 20 x u16 bitrows (row 0 = top, bit c = column c)
__global__ void __launch_bounds__(128) warp_per_env(const uint16_t *__restrict__ rows, const uint8_t *__restrict__ piece, int n, int *out) {
    const int lane = threadIdx.x & 31;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += warps) {
        const uint32_t row = lane < 20 ? rows[(size_t)e * 20 + lane] : 0u;
        // ballot transposition: bit r of m = row r has a cell in column c; the lane that owns column c keeps it
        uint32_t mine = 0u;
#pragma unroll
        for (int c = 0; c < 10; ++c) {
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, (row >> c) & 1u);
            if (lane == c) mine = m;
        }
        const int h = mine ? 20 - (__ffs(mine) - 1) : 0;           // lane k < 10: height of column k (lanes >= 10: 0)
        const int p = piece[e];
        int sum = 0;
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            const int s = lane + 32 * round;                        // slot = rot * 10 + col
            const int rot = s / 10, col = s - 10 * rot;
            const int cc = min(col, 10 - c_w[p][rot & 3]);          // loc clamps to 10 - w (:364)
            const int h0 = __shfl_sync(0xFFFFFFFFu, h, cc), h1 = __shfl_sync(0xFFFFFFFFu, h, cc + 1);
            const int h2 = __shfl_sync(0xFFFFFFFFu, h, cc + 2), h3 = __shfl_sync(0xFFFFFFFFu, h, cc + 3);
            if (s < 40) sum += drop_y(h0, h1, h2, h3, p, rot) + 1;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        if (lane == 0) out[e] = sum;
    }
}

// B: one thread per env, ten bit-columns (bit b = cell in row 19 - b), planes of 16-byte chunks as in the library
__global__ void __launch_bounds__(128) thread_per_env(const uint4 *__restrict__ cols, int64_t stride, const uint8_t *__restrict__ piece, int n, int *out) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint4 a = cols[e], b = cols[stride + e], c = cols[2 * stride + e];
        const uint32_t col[10] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y};
        int H[14];
#pragma unroll
        for (int k = 0; k < 10; ++k) H[k] = 32 - __clz(col[k]);
#pragma unroll
        for (int k = 10; k < 14; ++k) H[k] = 0;
        const int p = piece[e];
        int sum = 0;
#pragma unroll
        for (int rot = 0; rot < 4; ++rot) {
            const int8_t *bo = c_bo[p][rot];
            const int b0 = bo[0], b1 = bo[1], b2 = bo[2], b3 = bo[3], cmax = 10 - c_w[p][rot];
            int last = 0;
#pragma unroll
            for (int cc = 0; cc < 10; ++cc) {
                const int y = max(max(H[cc] - b0, H[cc + 1] - b1), max(H[cc + 2] - b2, H[cc + 3] - b3));
                last = cc <= cmax ? y : last;                       // loc clamps to 10 - w (:364)
                sum += last + 1;
            }
        }
        out[e] = sum;
    }
}

int main() {
    const int n = 1 << 20;
    std::vector<uint16_t> rows((size_t)n * 20);
    std::vector<uint32_t> cols((size_t)n * 12, 0u);
    std::vector<uint8_t> piece(n);
    uint64_t s = 12345;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); };
    for (int e = 0; e < n; ++e) {
        piece[e] = rnd() % 7;
        const int top = rnd() % 18;                                  // rows above `top` stay empty
        for (int r = 0; r < 20; ++r) rows[(size_t)e * 20 + r] = r < top ? 0 : (uint16_t)(rnd() & rnd() & 0x3FF);
        for (int c = 0; c < 10; ++c) {
            uint32_t v = 0;
            for (int r = 0; r < 20; ++r) v |= ((rows[(size_t)e * 20 + r] >> c) & 1u) << (19 - r);
            cols[(size_t)(c / 4) * n * 4 + (size_t)e * 4 + (c % 4)] = v;      // plane c/4, chunk e, word c%4
        }
    }
    uint16_t *d_rows; uint4 *d_cols; uint8_t *d_piece; int *d_a, *d_b;
    cudaMalloc(&d_rows, rows.size() * 2); cudaMalloc(&d_cols, cols.size() * 4); cudaMalloc(&d_piece, n);
    cudaMalloc(&d_a, n * 4); cudaMalloc(&d_b, n * 4);
    cudaMemcpy(d_rows, rows.data(), rows.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d_cols, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_piece, piece.data(), n, cudaMemcpyHostToDevice);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float msa = 0, msb = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        for (int k = 0; k < 10; ++k) warp_per_env<<<sms * 16, 128>>>(d_rows, d_piece, n, d_a);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&msa, e0, e1);
        cudaEventRecord(e0);
        for (int k = 0; k < 10; ++k) thread_per_env<<<sms * 16, 128>>>(d_cols, n, d_piece, n, d_b);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&msb, e0, e1);
    }
    std::vector<int> ha(n), hb(n);
    cudaMemcpy(ha.data(), d_a, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), d_b, n * 4, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (int e = 0; e < n; ++e) bad += ha[e] != hb[e];
    cudaError_t err = cudaGetLastError();
    printf("hard-drop rows of all 40 slots, %d envs (%s, %zu mismatches between the two)\n", n, cudaGetErrorString(err), bad);
    printf("A  warp per env   (20 x u16 bitrows, ballot transposition, lanes = slots): %8.4f ms per pass = %6.3f ns per env\n", msa / 10, msa / 10 * 1e6 / n);
    printf("B  thread per env (10 bit-columns in registers, this library's layout)    : %8.4f ms per pass = %6.3f ns per env\n", msb / 10, msb / 10 * 1e6 / n);
    printf("A / B = %.1f\n", msa / msb);
    return bad != 0;
}
