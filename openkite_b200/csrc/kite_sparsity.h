// kite_sparsity.h -- structural non-zeros of the kite Jacobians, shared by the kernels, the C ABI's
// sparsity queries and the CasADi-external shim (plain C++: no CUDA headers needed).
//   d f / d x: 104 non-zeros with a zero tether arm; the rows w_dot gain the r, q columns (+21) when the arm is non-zero.
//   d f / d u: 7 non-zeros (T -> v_dot0; dE -> v_dot0, v_dot2, w_dot1; dR -> v_dot1, w_dot0, w_dot2).
#pragma once
#if defined(__CUDACC__)
#define KITE_HD __host__ __device__
#else
#define KITE_HD
#endif

namespace kite {

KITE_HD constexpr bool jx_nz(int i, int j, bool arm) {
    if (i < 3) return !((i == 0 && j == 3) || (i == 2 && j == 5));
    if (i < 6) return (j < 6) || arm;
    if (i < 9) return (j < 3) || (j >= 9);
    return (j >= 3 && j < 6) || (j >= 9);
}
KITE_HD constexpr bool ju_nz(int i, int j) {
    return (i == 0 && j == 0) || (i == 0 && j == 1) || (i == 2 && j == 1) || (i == 4 && j == 1) ||
           (i == 1 && j == 2) || (i == 3 && j == 2) || (i == 5 && j == 2);
}
// rigid-body kinematics (kite.cpp:622-661): only the r_dot and q_dot rows are non-zero
KITE_HD constexpr bool jx_nz_rigid(int i, int j) { return i >= 6 && jx_nz(i, j, false); }

}  // namespace kite
