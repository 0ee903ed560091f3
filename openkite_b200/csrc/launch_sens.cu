#include <cstdint>

#include "kite_launch.h"
namespace kite {
template <bool ARM, bool RIGID, bool TMA_OUT>
static void go_fused(const SensArgs& a, cudaStream_t s) {
    // the opt-in is a per-device function attribute and a host may own contexts on several devices: set it on every
    // launch (a few hundred nanoseconds) instead of caching a per-process flag
    cudaFuncSetAttribute(k_sens_fused<ARM, RIGID, TMA_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SfCfg<ARM>::SMEM);
    // one warp per group at most: the steps of a group depend on each other, and the scratch (kite_rk4_sens_work_bytes)
    // holds one line per resident warp of a grid sized by the GROUP count
    const long ngroups = (a.B + 31) / 32;
    constexpr int W = SfCfg<ARM>::WARPS;
    const long want = (ngroups + W - 1) / W;
    const long sms = current_device_sms();
    const unsigned grid = (unsigned)(want < sms ? want : sms);          // persistent: one CTA per SM
    k_sens_fused<ARM, RIGID, TMA_OUT><<<grid, W * 32, SfCfg<ARM>::SMEM, s>>>(a);
}
int current_device_sms() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        return 0;
    return sms;
}
// Upper bound of the resident warps of the persistent kernel on ANY visible device (the scratch-size query has no
// context argument, and a host may own contexts on several devices).
long sens_fused_max_warps() {
    int ndev = 0, best = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) ndev = 0;
    for (int d = 0; d < ndev; ++d) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d) == cudaSuccess && sms > best) best = sms;
    }
    if (best <= 0) best = 256;                                          // no device visible: a safe upper bound for sizing
    return (long)best * SF_WARPS;
}
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn encode_tiled() {
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    }
    return fn;
}
// Tensor map of an [N][rows][B] FP64 array with row pitch ld doubles, box = [1][rows][8 units], 64-byte swizzle in shared
// memory.  False when the layout does not meet the TMA constraints (16-byte aligned base and pitch; an even B, because the
// TMA clips out-of-range columns in 16-byte units and an odd B would spill one double into the padding): the caller then
// uses the direct-store kernel.
bool sens_make_tensor_map(CUtensorMap* tm, double* base, long B, long ld, int rows, long N) {
    encode_tiled_fn enc = encode_tiled();
    if (!enc || !base || B <= 0 || N <= 0 || (B & 1) || B >= (1L << 31) || N >= (1L << 31) || ((uintptr_t)base & 15) || ((ld * 8) & 15) ||
        (double)ld * 8 * rows * (N > 1 ? 1 : 0) >= (double)(1L << 40) || ld * 8 >= (1L << 40))
        return false;
    const cuuint64_t dims[3] = {(cuuint64_t)B, (cuuint64_t)rows, (cuuint64_t)N};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)ld * 8 * (cuuint64_t)rows};
    const cuuint32_t box[3] = {8, (cuuint32_t)rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
void launch_sens_fused(const SensArgs& a, bool rigid, bool arm, bool tma_out, cudaStream_t s) {
    if (rigid) go_fused<false, true, false>(a, s);
    else if (arm) { if (tma_out) go_fused<true, false, true>(a, s); else go_fused<true, false, false>(a, s); }
    else { if (tma_out) go_fused<false, false, true>(a, s); else go_fused<false, false, false>(a, s); }
}
}  // namespace kite
