#include "kite_launch.h"
namespace kite {
void launch_ekf_state_jac(const EkfArgs& a, bool rigid, cudaStream_t s) {
    if (rigid) k_ekf_state_jac<true><<<blocks_for(a.B, 128), 128, 0, s>>>(a);
    else k_ekf_state_jac<false><<<blocks_for(a.B, 128), 128, 0, s>>>(a);
}
void launch_ekf_cov(const EkfArgs& a, bool rigid, bool arm, cudaStream_t s) {
    const unsigned gb = blocks_for(a.B * 16, 256);
    if (rigid) k_ekf_cov<false, true><<<gb, 256, 0, s>>>(a);
    else if (arm) k_ekf_cov<true, false><<<gb, 256, 0, s>>>(a);
    else k_ekf_cov<false, false><<<gb, 256, 0, s>>>(a);
}
void launch_ekf_update(const EkfUpdArgs& a, cudaStream_t s) {
    k_ekf_update<0><<<blocks_for(a.B, 128), 128, 0, s>>>(a);
}
}  // namespace kite
