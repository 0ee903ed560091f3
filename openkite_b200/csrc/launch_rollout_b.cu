#include "kite_launch.h"
#ifndef KITE_ROLLOUT_SMEM
#define KITE_ROLLOUT_SMEM false   // measured on B200: 55.4% (smem state, 4 blocks/SM, 40 B spill) vs 58.0% (registers, 3 blocks/SM)
#endif
namespace kite {
template <int UMODE>
static void go(const RolloutArgs& a, bool rigid, bool percoef, cudaStream_t s) {
    dim3 grid(blocks_for(a.B, ROLLOUT_BLOCK)), block(ROLLOUT_BLOCK);
    if (rigid) k_rk4_rollout<UMODE, true, false, false><<<grid, block, 0, s>>>(a);
    else if (percoef) k_rk4_rollout<UMODE, false, true, KITE_ROLLOUT_SMEM><<<grid, block, 0, s>>>(a);
    else k_rk4_rollout<UMODE, false, false, KITE_ROLLOUT_SMEM><<<grid, block, 0, s>>>(a);
}
void launch_rollout_23(const RolloutArgs& a, int umode, bool rigid, bool percoef, cudaStream_t s) {
    if (umode == 2) go<2>(a, rigid, percoef, s); else go<3>(a, rigid, percoef, s);
}
}  // namespace kite
