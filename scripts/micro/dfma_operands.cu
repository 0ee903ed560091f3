// Microbenchmark (developer experiment): FP64 pipe throughput by operand form and instruction mix, 148 x 8 CTAs x 256 threads.
//   0: a = fma(a, C, C)  (two constant operands; kite_fp64_peak)     1: a = fma(a, rx, ry)  (three register operands, two shared)
//   2: a_i = fma(a_i, b_i, c_i) (all distinct registers)              3: mix per 8 ops: 5 DFMA + 2 DMUL + 1 DADD, register operands
//   4: as 2 with only 4 independent chains (ILP 4)                    5: as 2 at 12 warps/SM (168-register occupancy: 3 CTAs x 128 threads)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double s0, double s1) {
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; ++i) { a[i] = s0 + i + threadIdx.x; b[i] = 0.9999999 + 1e-9 * (i + threadIdx.x * s1); c[i] = 1e-7 * (i + 1) * s1 + 1e-13 * threadIdx.x; }
    const double rx = b[0], ry = c[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = fma(a[i], 0.9999999, 1e-7);
                else if (MODE == 1) a[i] = fma(a[i], rx, ry);
                else if (MODE == 2 || MODE == 5) a[i] = fma(a[i], b[i], c[i]);
                else if (MODE == 4) a[i & 3] = fma(a[i & 3], b[i], c[i]);
                else { if (i < 5) a[i] = fma(a[i], b[i], c[i]); else if (i < 7) a[i] = a[i] * b[i]; else a[i] = a[i] + c[i]; }
            }
        }
    }
    double t = 0; for (int i = 0; i < 8; ++i) t += a[i];
    out[(long)blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE> void run(double* d, const char* name, int blocks, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    k<MODE><<<blocks, threads>>>(d, 100, 1.0, 1.0);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters, 1.0, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = 64.0 * iters * blocks * threads;           // FP64 instructions (thread level)
    printf("%-62s %.3f ms  %.2f T inst/s  (%.2f TFLOP/s if all were FMAs)\n", name, ms, inst / ms / 1e9, 2 * inst / ms / 1e9);
}
int main() {
    double* d; cudaMalloc(&d, 8 * 148 * 8 * 256);
    run<0>(d, "0 fma(a, C, C)", 148 * 8, 256);
    run<1>(d, "1 fma(a, rx, ry)", 148 * 8, 256);
    run<2>(d, "2 fma(a_i, b_i, c_i)", 148 * 8, 256);
    run<3>(d, "3 5 DFMA + 2 DMUL + 1 DADD, registers", 148 * 8, 256);
    run<4>(d, "4 fma(a_i, b_i, c_i), 4 chains", 148 * 8, 256);
    run<5>(d, "5 fma(a_i, b_i, c_i), 12 warps per SM (3 x 128)", 148 * 3, 128);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
