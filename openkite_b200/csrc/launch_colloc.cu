#include "kite_launch.h"
namespace kite {
void launch_colloc_eval(const CollocArgs& a, bool percoef, cudaStream_t s) {
    const unsigned gb = blocks_for(a.B, 32);
    if (a.M == 11) {
        dim3 block(32, 11);
        if (percoef) k_colloc_eval<true, 11><<<gb, block, 0, s>>>(a);
        else k_colloc_eval<false, 11><<<gb, block, 0, s>>>(a);
    } else {
        dim3 block(32, 8);
        if (percoef) k_colloc_eval<true, 8><<<gb, block, 0, s>>>(a);
        else k_colloc_eval<false, 8><<<gb, block, 0, s>>>(a);
    }
}
}  // namespace kite
