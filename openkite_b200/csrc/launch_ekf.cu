#include "kite_launch.h"
namespace kite {
template <bool ARM, bool RIGID>
static void go_predict(const EkfArgs& a, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_ekf_predict<ARM, RIGID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EfCfg<ARM>::SMEM);
        configured = true;
    }
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const long ngroups = (a.B + 31) / 32;
    const long want = (ngroups + EfCfg<ARM>::WARPS - 1) / EfCfg<ARM>::WARPS;
    const unsigned grid = (unsigned)(want < sms ? want : sms);          // persistent: one CTA per SM
    k_ekf_predict<ARM, RIGID><<<grid, EfCfg<ARM>::WARPS * 32, EfCfg<ARM>::SMEM, s>>>(a);
}
void launch_ekf_predict(const EkfArgs& a, bool rigid, bool arm, cudaStream_t s) {
    if (rigid) go_predict<false, true>(a, s);
    else if (arm) go_predict<true, false>(a, s);
    else go_predict<false, false>(a, s);
}
void launch_ekf_update(const EkfUpdArgs& a, cudaStream_t s) {
    constexpr int smem = (int)sizeof(double) * 91 * EKFU_BLOCK;
    static bool configured = false;
    if (!configured) { cudaFuncSetAttribute(k_ekf_update<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); configured = true; }
    k_ekf_update<0><<<blocks_for(a.B, EKFU_BLOCK), EKFU_BLOCK, smem, s>>>(a);
}
}  // namespace kite
