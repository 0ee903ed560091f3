// Microbenchmark (developer experiment): shared-memory wavefronts of octet-broadcast loads.
//   mode 0: LDS.64,  4 distinct 8-byte addresses per warp (lane >> 3), contiguous 32 B
//   mode 1: LDS.128, 4 distinct 16-byte addresses per warp (lane >> 3), contiguous 64 B
//   mode 2: LDS.128, 8 distinct 16-byte addresses per warp (lane >> 2), contiguous 128 B
//   mode 3: LDS.64,  8 distinct 8-byte addresses per warp (lane >> 2), contiguous 64 B
// Run under: ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum ./lds_broadcast
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters) {
    __shared__ __align__(16) double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 16; ++s) {
            if (MODE == 0) acc += sm[(s + it) % 64 * 4 + (lane >> 3)];
            else if (MODE == 3) acc += sm[(s + it) % 64 * 8 + (lane >> 2)];
            else if (MODE == 1) { double2 v = *reinterpret_cast<const double2*>(&sm[(s + it) % 64 * 8 + (lane >> 3) * 2]); acc += v.x + v.y; }
            else { double2 v = *reinterpret_cast<const double2*>(&sm[(s + it) % 64 * 16 + (lane >> 2) * 2]); acc += v.x + v.y; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    double* d; cudaMalloc(&d, 8 * 148 * 256);
    k<0><<<148, 256>>>(d, 1000); k<1><<<148, 256>>>(d, 1000); k<2><<<148, 256>>>(d, 1000); k<3><<<148, 256>>>(d, 1000);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
