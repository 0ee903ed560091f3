#include "kite_launch.h"
namespace kite {
template <bool ARM, bool RIGID>
static void go_fused(const SensArgs& a, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_sens_fused<ARM, RIGID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SfCfg<ARM>::SMEM);
        configured = true;
    }
    const long ngroups = (a.B + 31) / 32;
    constexpr int W = SfCfg<ARM>::WARPS;
    const long want = (ngroups + W - 1) / W;
    const long sms = sens_fused_max_warps() / SF_WARPS;
    const unsigned grid = (unsigned)(want < sms ? want : sms);          // persistent: one CTA per SM
    k_sens_fused<ARM, RIGID><<<grid, W * 32, SfCfg<ARM>::SMEM, s>>>(a);
}
long sens_fused_max_warps() {
    static long warps = 0;
    if (!warps) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            return 256L * SF_WARPS;                                     // no device visible: a safe upper bound for sizing
        warps = (long)sms * SF_WARPS;
    }
    return warps;
}
void launch_sens_fused(const SensArgs& a, bool rigid, bool arm, cudaStream_t s) {
    if (rigid) go_fused<false, true>(a, s);
    else if (arm) go_fused<true, false>(a, s);
    else go_fused<false, false>(a, s);
}
}  // namespace kite
