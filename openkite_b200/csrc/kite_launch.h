// kite_launch.h -- launcher functions, one translation unit per kernel family so nvcc can build them in parallel.
#pragma once
#include "kite_kernels.cuh"

namespace kite {
void launch_point_eval(const PointArgs& a, bool rigid, bool percoef, bool jac, cudaStream_t s);
void launch_rollout_01(const RolloutArgs& a, int umode, bool rigid, bool percoef, cudaStream_t s);
void launch_rollout_23(const RolloutArgs& a, int umode, bool rigid, bool percoef, cudaStream_t s);
void launch_synth_inputs(const SynthArgs& a, cudaStream_t s);
void launch_synth_id_params(const SynthParamArgs& a, cudaStream_t s);
void launch_sens_fused(const SensArgs& a, bool rigid, bool arm, bool tma_out, cudaStream_t s);
bool sens_make_tensor_map(CUtensorMap* tm, double* base, long B, long ld, int rows, long N);   // false: layout not TMA-eligible
int current_device_sms();      // SM count of the CURRENT device (0 if none)
long sens_fused_max_warps();   // upper bound of the resident warps of the persistent fused kernel over all visible devices (scratch sizing)
void launch_ekf_predict(const EkfArgs& a, bool rigid, bool arm, double* lines, cudaStream_t s);
size_t ekf_predict_scratch_bytes();   // pre-step state lines of the resident warps of the TMA kernel (independent of B)
void launch_ekf_update(const EkfUpdArgs& a, cudaStream_t s);
void launch_colloc_eval(const CollocArgs& a, bool percoef, int fmt, cudaStream_t s);   // fmt 0 dense, 1 / 2 compact (no arm / arm), 3 values only
void launch_colloc_cost(const CostArgs& a, cudaStream_t s);
void launch_math_selftest(const double* x, double* out, long n, int which, cudaStream_t s);
void launch_fp64_peak(double* out, int iters, int blocks, int threads, cudaStream_t s);
void launch_fp64_peak3(double* out, int iters, int blocks, int threads, cudaStream_t s);
inline unsigned blocks_for(long n, int bs) { return (unsigned)((n + bs - 1) / bs); }
}  // namespace kite
