#include "kite_launch.h"
namespace kite {
void launch_point_eval(const PointArgs& a, bool rigid, bool percoef, bool jac, cudaStream_t s) {
    dim3 grid(blocks_for(a.B, 128)), block(128);
    if (rigid) {
        if (jac) k_point_eval<true, false, true><<<grid, block, 0, s>>>(a);
        else k_point_eval<true, false, false><<<grid, block, 0, s>>>(a);
    } else if (percoef) {
        if (jac) k_point_eval<false, true, true><<<grid, block, 0, s>>>(a);
        else k_point_eval<false, true, false><<<grid, block, 0, s>>>(a);
    } else {
        if (jac) k_point_eval<false, false, true><<<grid, block, 0, s>>>(a);
        else k_point_eval<false, false, false><<<grid, block, 0, s>>>(a);
    }
}
void launch_synth_inputs(const SynthArgs& a, cudaStream_t s) {
    long gy = a.N < 1 ? 1 : (a.N > 4096 ? 4096 : a.N);
    dim3 grid(blocks_for(a.B, 256), (unsigned)gy), block(256);
    k_synth_inputs<0><<<grid, block, 0, s>>>(a);
}
void launch_synth_id_params(const SynthParamArgs& a, cudaStream_t s) {
    k_synth_id_params<0><<<blocks_for(a.B, 256), 256, 0, s>>>(a);
}
void launch_math_selftest(const double* x, double* out, long n, int which, cudaStream_t s) {
    k_math_selftest<0><<<blocks_for(n, 256), 256, 0, s>>>(x, out, n, which);
}
void launch_fp64_peak(double* out, int iters, int blocks, int threads, cudaStream_t s) {
    k_fp64_peak<0><<<blocks, threads, 0, s>>>(out, iters, 1.0);
}
void launch_fp64_peak3(double* out, int iters, int blocks, int threads, cudaStream_t s) {
    k_fp64_peak3<0><<<blocks, threads, 0, s>>>(out, iters, 1.0);
}
}  // namespace kite
