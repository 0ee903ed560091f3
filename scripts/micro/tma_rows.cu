// Microbenchmark (developer experiment, not part of the product): what does the TMA unit of an SM sustain on boxes of short
// rows?  The EKF predict moves one [169][8 filters] FP64 box (169 rows of 64 bytes, row pitch B * 8 bytes) in and one out per
// round; this kernel does only that traffic, with the same double buffering and per-warp mbarriers, for rows of 64 / 128 / 256
// bytes, and reports the time per 1 M filters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rows tma_rows.cu && ./tma_rows
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Args { CUtensorMap in, out; long rounds; int F; unsigned box_bytes; int mode; const double* lin_in; double* lin_out; };   // mode 1: load, 2: store, 3: both; lin_*: tiled layout [rounds][169][F], one contiguous 1-D bulk copy per box

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1) k_tma(const __grid_constant__ Args a, double* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const size_t BOX = (a.box_bytes + 1023) / 1024 * 1024;
    unsigned char* wb = smem + (size_t)warp * 2 * BOX;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)W * 2 * BOX) + warp * 2;
    if (lane < 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bars + lane)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const long stride = (long)gridDim.x * W, first = (long)blockIdx.x * W + warp;
    auto load = [&](unsigned t, long r) {
        if (lane == 0) {
            const unsigned bar = smem_u32(bars + (t & 1));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(a.box_bytes) : "memory");
            if (a.lin_in)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(wb + (t & 1) * BOX)), "l"(a.lin_in + r * (long)(a.box_bytes / 8)), "r"(a.box_bytes), "r"(bar) : "memory");
            else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(smem_u32(wb + (t & 1) * BOX)), "l"(reinterpret_cast<unsigned long long>(&a.in)), "r"((int)(r * a.F)), "r"(0), "r"(bar) : "memory");
        }
    };
    unsigned t = 0;
    double acc = 0.0;
    if ((a.mode & 1) && first < a.rounds) load(0, first);
    for (long r = first; r < a.rounds; r += stride, ++t) {
        unsigned char* box = wb + (t & 1) * BOX;
        if (a.mode & 1) {
            // wait for this round's box
            const unsigned bar = smem_u32(bars + (t & 1)), par = (t >> 1) & 1;
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" :: "r"(bar), "r"(par) : "memory");
            acc += reinterpret_cast<double*>(box)[lane];
        }
        // the other buffer: its store (issued last round) must have been read out before the next load lands in it
        if (lane == 0) { if (a.mode == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
        __syncwarp();
        if ((a.mode & 1) && r + stride < a.rounds) load(t + 1, r + stride);
        if (a.mode & 2) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (a.lin_out)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(a.lin_out + r * (long)(a.box_bytes / 8)), "r"(smem_u32(box)), "r"(a.box_bytes) : "memory");
                else
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             :: "l"(reinterpret_cast<unsigned long long>(&a.out)), "r"((int)(r * a.F)), "r"(0), "r"(smem_u32(box)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 12345.678) sink[0] = acc;
}

int main(int argc, char** argv) {
    const long B = argc > 1 ? atol(argv[1]) : 1048576;
    const long ld = B + (argc > 2 ? atol(argv[2]) : 0);      // row pitch in doubles (B + padding)
    printf("B %ld, row pitch %ld doubles\n", B, ld);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaFree(0));
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    encode_tiled_fn enc = (encode_tiled_fn)p;
    double *P, *Pn, *sink;
    CK(cudaMalloc(&P, sizeof(double) * 169 * ld)); CK(cudaMalloc(&Pn, sizeof(double) * 169 * ld)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(P, 0, sizeof(double) * 169 * ld));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int Fs[3] = {8, 16, 32};
    for (int fi = 0; fi < 3; ++fi) {
        const int F = Fs[fi];
        for (int sw = 0; sw < 2; ++sw) {
            if (sw && F == 32) continue;
            Args a; a.F = F; a.box_bytes = 169u * F * 8u; a.rounds = B / F; a.lin_in = nullptr; a.lin_out = nullptr;
            const cuuint64_t dims[2] = {(cuuint64_t)B, 169};
            const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
            const cuuint32_t box[2] = {(cuuint32_t)F, 169};
            const cuuint32_t estr[2] = {1, 1};
            const CUtensorMapSwizzle swz = !sw ? CU_TENSOR_MAP_SWIZZLE_NONE : (F == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
            if (enc(&a.in, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, P, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
                enc(&a.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, Pn, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed F=%d\n", F); continue; }
            const size_t BOX = (a.box_bytes + 1023) / 1024 * 1024;
            for (int W = 1; W <= 8; ++W) {
                const size_t smem = (size_t)W * 2 * BOX + 16 * W;
                if (smem > 227 * 1024) break;
                for (int mode = 1; mode <= 3; ++mode) {
                    a.mode = mode;
                    float best = 1e9f;
                    for (int rep = 0; rep < 4; ++rep) {
                        cudaEventRecord(e0);
                        k_tma<<<sms, W * 32, smem>>>(a, sink);
                        cudaEventRecord(e1);
                        CK(cudaEventSynchronize(e1));
                        float ms; cudaEventElapsedTime(&ms, e0, e1);
                        if (rep && ms < best) best = ms;
                    }
                    const double bytes = 169.0 * 8 * B * ((mode & 1) + (mode >> 1));
                    printf("rows %3d B swizzle %d  warps/SM %d  %s  %.3f ms per %ld filters  %.2f TB/s  %.1f cycles per row request per SM (1.965 GHz)\n",
                           F * 8, sw, W, mode == 1 ? "load " : mode == 2 ? "store" : "both ", best, B, bytes / best / 1e9,
                           best * 1e-3 * 1.965e9 / (169.0 * (B / F) / sms * ((mode & 1) + (mode >> 1))));
                }
            }
        }
    }
    // tiled layout: the same bytes as [rounds][169][8]: one contiguous 10.8 KB bulk copy per box
    {
        Args a; a.F = 8; a.box_bytes = 169u * 8u * 8u; a.rounds = B / 8; a.lin_in = P; a.lin_out = Pn;
        const size_t BOX = (a.box_bytes + 1023) / 1024 * 1024;
        for (int W = 2; W <= 8; W += 2) {
            const size_t smem = (size_t)W * 2 * BOX + 16 * W;
            for (int mode = 1; mode <= 3; ++mode) {
                a.mode = mode;
                float best = 1e9f;
                for (int rep = 0; rep < 4; ++rep) {
                    cudaEventRecord(e0);
                    k_tma<<<sms, W * 32, smem>>>(a, sink);
                    cudaEventRecord(e1);
                    CK(cudaEventSynchronize(e1));
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (rep && ms < best) best = ms;
                }
                const double bytes = 169.0 * 8 * B * ((mode & 1) + (mode >> 1));
                printf("tiled [B/8][169][8], 1-D bulk copies of %u B  warps/SM %d  %s  %.3f ms per %ld filters  %.2f TB/s\n",
                       a.box_bytes, W, mode == 1 ? "load " : mode == 2 ? "store" : "both ", best, B, bytes / best / 1e9);
            }
        }
    }
    CK(cudaGetLastError());
    return 0;
}
