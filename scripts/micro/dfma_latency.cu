// Microbenchmark (developer tool): dependent-DFMA latency and per-SM throughput vs resident warps and ILP on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double seed) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + i;
    const double m = 0.9999999, b = 1e-7;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = double(t1 - t0) / (double(iters) * 16 * ILP);
}
template <int ILP>
void run(int warps_per_sm, int sms, double* d) {
    int iters = 4096;
    k<ILP><<<sms, warps_per_sm * 32>>>(d, iters, 1.0);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ILP><<<sms, warps_per_sm * 32>>>(d, iters, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc; cudaMemcpy(&cyc, d + (size_t)sms * warps_per_sm * 32, 8, cudaMemcpyDeviceToHost);
    double tf = 2.0 * iters * 16.0 * ILP * sms * warps_per_sm * 32 / (ms * 1e-3) / 1e12;
    printf("ILP=%d warps/SM=%2d: %.2f cycles per DFMA per warp (issue interval), %.2f TFLOP/s\n", ILP, warps_per_sm, cyc, tf);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* d; cudaMalloc(&d, sizeof(double) * ((size_t)sms * 1024 + 8));
    for (int w : {4, 8, 12, 16, 32}) { run<1>(w, sms, d); run<2>(w, sms, d); run<4>(w, sms, d); run<8>(w, sms, d); }
    return 0;
}
