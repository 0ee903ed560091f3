#include "kite_launch.h"
namespace kite {
void launch_sens_stage_jac(const SensArgs& a, bool rigid, cudaStream_t s) {
    if (rigid) k_sens_stage_jac<true><<<blocks_for(a.B, 128), 128, 0, s>>>(a);
    else k_sens_stage_jac<false><<<blocks_for(a.B, 128), 128, 0, s>>>(a);
}
void launch_sens_propagate(const SensArgs& a, bool rigid, bool arm, cudaStream_t s) {
    const unsigned gb = blocks_for(a.B * 16, 256);
    if (rigid) k_sens_propagate<false, true><<<gb, 256, 0, s>>>(a);
    else if (arm) k_sens_propagate<true, false><<<gb, 256, 0, s>>>(a);
    else k_sens_propagate<false, false><<<gb, 256, 0, s>>>(a);
}
}  // namespace kite
