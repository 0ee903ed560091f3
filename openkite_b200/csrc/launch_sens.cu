#include "kite_launch.h"
namespace kite {
void launch_sens_stage_jac(const SensArgs& a, bool rigid, cudaStream_t s) {
    if (rigid) k_sens_stage_jac<true><<<blocks_for(a.B, 128), 128, 0, s>>>(a);
    else k_sens_stage_jac<false><<<blocks_for(a.B, 128), 128, 0, s>>>(a);
}
template <bool ARM, bool RIGID>
static void go_propagate(const SensArgs& a, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_sens_propagate<ARM, RIGID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SensSmem));
        configured = true;
    }
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const unsigned nb = blocks_for(a.B, SENS_UNITS);
    const unsigned grid = nb < (unsigned)sms ? nb : (unsigned)sms;      // persistent: one CTA per SM
    k_sens_propagate<ARM, RIGID><<<grid, SENS_THREADS, sizeof(SensSmem), s>>>(a);
}
void launch_sens_propagate(const SensArgs& a, bool rigid, bool arm, cudaStream_t s) {
    if (rigid) go_propagate<false, true>(a, s);
    else if (arm) go_propagate<true, false>(a, s);
    else go_propagate<false, false>(a, s);
}
}  // namespace kite
