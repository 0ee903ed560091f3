// =====================================================================================
// kite_kernels.cuh -- sm_100a kernels of the batched kite engine.
//
// Data layout: structure-of-arrays, component c of unit i at ptr[c*ld + i]; a warp touches 32
// consecutive units of one component = one 256 B coalesced request.  All kernels are FP64
// FMA-pipe bound (no tensor cores: the work is scalar per-trajectory arithmetic);
// model constants travel as a __grid_constant__ kernel parameter so they are read straight from
// the constant bank as DFMA operands and cost no registers.
// =====================================================================================
#pragma once
#include <cuda_pipeline.h>

#include "kite_model.cuh"

namespace kite {

// ---- compact Jacobian storage -----------------------------------------------------------------
// Structural non-zeros of [Jx | Ju], row-major slot numbering.  The 21 entries
// that only exist with a tether arm (rows w_dot, cols r,q) get slots too but are only touched when
// has_arm.  JAC_SLOTS slots of [ld] doubles each.
__host__ __device__ constexpr bool jx_nz(int i, int j, bool arm) {
    if (i < 3) return !((i == 0 && j == 3) || (i == 2 && j == 5));
    if (i < 6) return (j < 6) || arm;
    if (i < 9) return (j < 3) || (j >= 9);
    return (j >= 3 && j < 6) || (j >= 9);
}
constexpr int JX_SLOTS = 125;
__host__ __device__ constexpr bool ju_nz(int i, int j) {
    return (i == 0 && j == 0) || (i == 0 && j == 1) || (i == 2 && j == 1) || (i == 4 && j == 1) ||
           (i == 1 && j == 2) || (i == 3 && j == 2) || (i == 5 && j == 2);
}
constexpr int JAC_SLOTS = JX_SLOTS + 7;   // 132
struct SlotTab { int jx[13][13]; int ju[13][3]; int col[13][16]; };
constexpr int JAC_SLOTS_NOARM = 111;      // 104 + 7: the slots a zero-arm model touches are the first 111
constexpr SlotTab make_slot_tab() {
    SlotTab t{};
    int s = 0;
    for (int i = 0; i < 13; ++i)
        for (int j = 0; j < 13; ++j) t.jx[i][j] = jx_nz(i, j, false) ? s++ : -1;
    for (int i = 0; i < 13; ++i)
        for (int j = 0; j < 3; ++j) t.ju[i][j] = ju_nz(i, j) ? s++ : -1;
    for (int i = 0; i < 13; ++i)
        for (int j = 0; j < 13; ++j)
            if (jx_nz(i, j, true) && !jx_nz(i, j, false)) t.jx[i][j] = s++;
    // slot of entry (row i, tangent column c) of [Jx | Ju]
    for (int i = 0; i < 13; ++i)
        for (int c = 0; c < 16; ++c) t.col[i][c] = (c < 13) ? t.jx[i][c] : t.ju[i][c - 13];
    return t;
}
__device__ constexpr SlotTab SLOT_TAB = make_slot_tab();
__device__ __forceinline__ constexpr int jx_slot(int i, int j) { return SLOT_TAB.jx[i][j]; }
__device__ __forceinline__ constexpr int ju_slot(int i, int j) { return SLOT_TAB.ju[i][j]; }
static_assert(make_slot_tab().jx[12][12] == 103, "slot numbering");
static_assert(make_slot_tab().ju[5][2] == JAC_SLOTS_NOARM - 1, "slot numbering");
static_assert(make_slot_tab().jx[5][12] == JAC_SLOTS - 1, "slot numbering");

// Sink: compact slots, SoA over units.
struct CompactSink {
    double* base;   // &J[0*ld + unit]
    long ld;
    __device__ __forceinline__ void jx(int i, int j, double v) const { base[(long)jx_slot(i, j) * ld] = v; }
    __device__ __forceinline__ void ju(int i, int j, double v) const { base[(long)ju_slot(i, j) * ld] = v; }
};
// Sink: dense 13x13 / 13x3 row-major, SoA over units (buffers pre-zeroed by the caller).
struct DenseSink {
    double* jxp; double* jup; long ld;
    __device__ __forceinline__ void jx(int i, int j, double v) const { if (jxp) jxp[(long)(i * 13 + j) * ld] = v; }
    __device__ __forceinline__ void ju(int i, int j, double v) const { if (jup) jup[(long)(i * 3 + j) * ld] = v; }
};

__device__ __forceinline__ void load_coef(const KiteConsts& K, const double* __restrict__ p, long ld, long i, AeroCoef& A) {
    double raw[21];
#pragma unroll
    for (int c = 0; c < 21; ++c) raw[c] = __ldg(p + (long)c * ld + i);
    derive_coef(K, raw, A);
}

__device__ __forceinline__ bool all_finite13(const double (&x)[13]) {
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 13; ++c) s += x[c] * 0.0;     // NaN/Inf poison the sum
    return s == 0.0;
}

// ================================================================================================
// rhs_batch / jac_batch : pointwise evaluators (kite.cpp:324, :327-328)
// ================================================================================================
struct PointArgs {
    KiteConsts K;
    long B, ld;
    const double* x; const double* u; const double* p;
    double* f; double* Jx; double* Ju;
};

template <bool RIGID, bool PERCOEF, bool JAC>
__global__ void __launch_bounds__(128) k_point_eval(const __grid_constant__ PointArgs a) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.B) return;
    double x[13], u[3], f[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) x[c] = __ldg(a.x + (long)c * a.ld + i);
#pragma unroll
    for (int c = 0; c < 3; ++c) u[c] = a.u ? __ldg(a.u + (long)c * a.ld + i) : 0.0;
    AeroCoef A = a.K.A;
    if constexpr (PERCOEF) load_coef(a.K, a.p, a.ld, i, A);
    if constexpr (JAC) {
        DenseSink s{a.Jx ? a.Jx + i : nullptr, a.Ju ? a.Ju + i : nullptr, a.ld};
        model_eval<RIGID, true>(a.K, A, x, u, f, s);
    } else {
        NoSink s;
        model_eval<RIGID, false>(a.K, A, x, u, f, s);
    }
    if (a.f) {
#pragma unroll
        for (int c = 0; c < 13; ++c) a.f[(long)c * a.ld + i] = f[c];
    }
}

// ================================================================================================
// rk4_rollout : B trajectories x N steps, thread per trajectory, state in registers.
//   HBM traffic per state-step: 24 B of controls (KITE_U_PER_STEP), nothing else -> FP64-pipe bound.
// ================================================================================================
struct RolloutArgs {
    KiteConsts K;
    long B, ld, N;
    double h;
    const double* x0; const double* u; const double* p;
    double* xf; double* traj; long save_every;
    const double* y; double* cost;
    int32_t* status;
    long index0;
};

constexpr int ROLLOUT_BLOCK = 128;

template <int UMODE, bool RIGID, bool PERCOEF, bool SMEM>
__global__ void __launch_bounds__(ROLLOUT_BLOCK, 3) k_rk4_rollout(const __grid_constant__ RolloutArgs a) {
    // SMEM: the step base state is parked in shared memory during the four stages (rk4_step_xsmem), [13][block] columns
    __shared__ double sh[SMEM ? 13 * ROLLOUT_BLOCK : 1];
    double* const sx = sh + threadIdx.x;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.B) return;
    double x[13], u[3], un[3];
    if constexpr (UMODE == 3) {
        synth_x0((uint64_t)(a.index0 + i), x);
    } else {
#pragma unroll
        for (int c = 0; c < 13; ++c) x[c] = __ldg(a.x0 + (long)c * a.ld + i);
    }
    AeroCoef A = a.K.A;
    if constexpr (PERCOEF) load_coef(a.K, a.p, a.ld, i, A);

    // first control
    if constexpr (UMODE == 0 || UMODE == 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) un[c] = __ldg(a.u + (long)c * a.ld + i);
    } else if constexpr (UMODE == 2) {
#pragma unroll
        for (int c = 0; c < 3; ++c) un[c] = __ldg(a.u + c);
    } else {
        synth_control((uint64_t)(a.index0 + i), 0, un);
    }
    double cost = 0.0;
    const double Qc[13] = {1e3, 1e2, 1e2, 1e2, 1e2, 1e2, 1e1, 1e1, 1e2, 1e2, 1e2, 1e2, 1e2};  // kite_identification_test.cpp:193
    long next_save = a.save_every;
    long saved = 0;
    for (long k = 0; k < a.N; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) u[c] = un[c];
        // software prefetch of the next step's controls: the load is in flight during the 4 RHS evaluations
        if (k + 1 < a.N) {
            if constexpr (UMODE == 1) {
                const double* up = a.u + (long)(k + 1) * 3 * a.ld + i;
#pragma unroll
                for (int c = 0; c < 3; ++c) un[c] = __ldg(up + (long)c * a.ld);
            } else if constexpr (UMODE == 2) {
#pragma unroll
                for (int c = 0; c < 3; ++c) un[c] = __ldg(a.u + (k + 1) * 3 + c);
            } else if constexpr (UMODE == 3) {
                synth_control((uint64_t)(a.index0 + i), (uint64_t)(k + 1), un);
            }
        }
        if constexpr (SMEM) rk4_step_xsmem<RIGID>(a.K, A, sx, ROLLOUT_BLOCK, x, u, a.h);
        else rk4_step<RIGID>(a.K, A, x, u, a.h);
        if (a.y) {                                  // uniform branch: identification cost fused into the rollout
            double e = 0.0;
#pragma unroll
            for (int c = 0; c < 13; ++c) {
                const double dlt = __ldg(a.y + k * 13 + c) - x[c];
                e = fma(Qc[c] * dlt, dlt, e);
            }
            cost += e;
        }
        if (a.traj && k + 1 == next_save) {
#pragma unroll
            for (int c = 0; c < 13; ++c) a.traj[((long)saved * 13 + c) * a.ld + i] = x[c];
            ++saved; next_save += a.save_every;
        }
    }
#pragma unroll
    for (int c = 0; c < 13; ++c) a.xf[(long)c * a.ld + i] = x[c];
    if (a.y) a.cost[i] = cost * (1.0 / (double)a.N);
    if (a.status) a.status[i] = all_finite13(x) ? 0 : 1;
}

// Fill the synthetic workload buffers (x0 [13][ld], u [N][3][ld]).
struct SynthArgs { long B, ld, N, index0; double* x0; double* u; };
template <int DUMMY = 0>
__global__ void __launch_bounds__(256) k_synth_inputs(const __grid_constant__ SynthArgs a) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.B) return;
    if (blockIdx.y == 0 && a.x0) {
        double x0[13];
        synth_x0((uint64_t)(a.index0 + i), x0);
#pragma unroll
        for (int c = 0; c < 13; ++c) a.x0[(long)c * a.ld + i] = x0[c];
    }
    if (a.u) {
        for (long k = blockIdx.y; k < a.N; k += gridDim.y) {
            double u[3];
            synth_control((uint64_t)(a.index0 + i), (uint64_t)k, u);
#pragma unroll
            for (int c = 0; c < 3; ++c) a.u[((long)k * 3 + c) * a.ld + i] = u[c];
        }
    }
}

// ================================================================================================
// RK4 step sensitivities, two kernels.
//   A: thread per unit   -- primal RK4 stages + analytic stage Jacobians -> compact scratch Jw[4][132][ld]
//   B: 16 lanes per unit -- lane c owns tangent column c of [Phi | Gamma]; S_i = J_i' + a_i h Jx_i S_{i-1}
//      chained through the tableau with S in registers; stage Jacobian entries are broadcast loads.
// ================================================================================================
struct SensArgs {
    KiteConsts K;
    long B, ld;
    double h;
    const double* x; const double* u;
    double* xn; double* Phi; double* Gamma;
    double* Jw;     // [4][JAC_SLOTS][ld]
};

template <bool RIGID>
__global__ void __launch_bounds__(128) k_sens_stage_jac(const __grid_constant__ SensArgs a) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.B) return;
    double x[13], u[3], k[13], acc[13], xt[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) { x[c] = __ldg(a.x + (long)c * a.ld + i); xt[c] = x[c]; acc[c] = 0.0; }
#pragma unroll
    for (int c = 0; c < 3; ++c) u[c] = __ldg(a.u + (long)c * a.ld + i);
    const double hh = 0.5 * a.h;
    const long stage_stride = (long)JAC_SLOTS * a.ld;
    // stage loop kept rolled: one copy of the f + Jacobian code in the instruction stream (the 4x unrolled body
    // stalled on instruction fetch: no_instruction 1.34 per issue in profiles/r1a)
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
        CompactSink s{a.Jw + st * stage_stride + i, a.ld};
        model_eval<RIGID, true>(a.K, a.K.A, xt, u, k, s);
        const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
        const double an = (st == 2) ? a.h : hh;
#pragma unroll
        for (int c = 0; c < 13; ++c) { acc[c] = fma(wgt, k[c], acc[c]); xt[c] = fma(an, k[c], x[c]); }
    }
    const double h6 = a.h / 6.0;
#pragma unroll
    for (int c = 0; c < 13; ++c) a.xn[(long)c * a.ld + i] = fma(h6, acc[c], x[c]);
}

// ---- kernel B: tangent propagation ---------------------------------------------------------------------
// CTA = 256 threads = 32 units x 8 lanes; lane l owns tangent columns {2l, 2l+1} of [Phi | Gamma] (13 state seeds,
// 3 control seeds).  Per stage the unit's compact Jacobian is staged global -> shared with cp.async (coalesced
// 256 B rows, double buffered against the previous stage's FMAs), transposed to tile[unit][slot] so that the 8 lanes
// of a unit read it back as broadcast LDS.128 (two entries per load, 1 load per 4 DFMA).  S, S_next and the tableau
// accumulator (3 x 13 x 2 doubles) live in registers; results leave through shared memory as coalesced 256 B rows.
constexpr int SENS_UNITS = 32;                 // units per CTA
constexpr int SENS_TS = 133;                   // tile row stride in doubles: odd => the 32 units of a fill row and the 4 units of a
                                               // broadcast read land in distinct banks (profiles/r1c: stride 134 had 4-way conflicts)
constexpr int SENS_OS = 33;                    // output staging row stride (padded)
constexpr int SENS_THREADS = 256;

constexpr int SENS_RING = 4;                   // stage tiles in flight (cp.async ring): the next batch streams in behind the FMAs

struct SensSmem {
    double tile[SENS_RING][SENS_UNITS * SENS_TS];   // ring of stage Jacobian tiles [unit][slot]
    double out[208 * SENS_OS];                      // output staging [component row][unit]
};

template <bool ARM, bool RIGID>
__device__ __forceinline__ void sens_fill_tile(double* __restrict__ tile, const double* __restrict__ Jst, long ld, long unit0,
                                               long B, int tid) {
    constexpr int NS = ARM ? JAC_SLOTS : JAC_SLOTS_NOARM;
    const int u = tid & 31;
    if (unit0 + u < B) {
        const double* src = Jst + unit0 + u;
        double* dst = tile + u * SENS_TS;
#pragma unroll 4
        for (int sl = tid >> 5; sl < NS; sl += SENS_THREADS / 32) __pipeline_memcpy_async(dst + sl, src + (long)sl * ld, 8);
    }
}

// Persistent CTA: loops over batches of 32 units; tile t = batch * 4 + stage.  Tiles t+1..t+3 are always in flight.
template <bool ARM, bool RIGID>
__global__ void __launch_bounds__(SENS_THREADS, 1) k_sens_propagate(const __grid_constant__ SensArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SensSmem& sm = *reinterpret_cast<SensSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int lu = tid >> 3;                       // unit within the batch (8 consecutive lanes share a unit)
    const int l = tid & 7;
    const int c0 = 2 * l, c1 = 2 * l + 1;          // tangent columns of this lane
    const long stage_stride = (long)JAC_SLOTS * a.ld;
    const long nbatch = (a.B + SENS_UNITS - 1) / SENS_UNITS;
    const long my_batches = (nbatch > blockIdx.x) ? (nbatch - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long ntiles = my_batches * 4;
    auto issue = [&](long t) {                     // tile t of this CTA -> ring slot t % RING (always commits a group)
        if (t < ntiles) {
            const long batch = blockIdx.x + (t >> 2) * gridDim.x;
            sens_fill_tile<ARM, RIGID>(sm.tile[t % SENS_RING], a.Jw + (t & 3) * stage_stride, a.ld, batch * SENS_UNITS, a.B, tid);
        }
        __pipeline_commit();
    };
    // pad slot (index JAC_SLOTS of every tile row) reads as 0.0: structural zeros of a Jacobian column point at it
    for (int t = tid; t < SENS_RING * SENS_UNITS; t += SENS_THREADS) sm.tile[t / SENS_UNITS][(t % SENS_UNITS) * SENS_TS + JAC_SLOTS] = 0.0;
    for (int t = 0; t < SENS_RING - 1; ++t) issue(t);
    // per-lane slot indices of Jacobian columns c0, c1 (13 rows each), one byte per row packed into registers:
    // stage 1 has D = E, so S_1 = [Jx | Ju] E is a pure gather of two columns -- no FMAs, 26 loads instead of 111
    unsigned pk0[4] = {0, 0, 0, 0}, pk1[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        int s0 = SLOT_TAB.col[i][c0], s1 = SLOT_TAB.col[i][c1];
        if (s0 < 0 || (!ARM && s0 >= JAC_SLOTS_NOARM) || (RIGID && (i < 6 || c0 >= 13))) s0 = JAC_SLOTS;
        if (s1 < 0 || (!ARM && s1 >= JAC_SLOTS_NOARM) || (RIGID && (i < 6 || c1 >= 13))) s1 = JAC_SLOTS;
        pk0[i >> 2] |= (unsigned)s0 << (8 * (i & 3));
        pk1[i >> 2] |= (unsigned)s1 << (8 * (i & 3));
    }

    // Stage recursion in "input tangent" form, which is uniform across lanes (no per-lane gather of Jacobian columns):
    //   D_i = E + a_i h S_{i-1}   (E = seed matrix [I | 0]),   S_i = Jx_i D_i + Ju_i Eu   (Eu = [0 | I]),
    //   [Phi | Gamma] = E + h/6 (S_1 + 2 S_2 + 2 S_3 + S_4).
    // Each lane carries columns c0, c1 of D/S and of the accumulator; Jacobian entries are broadcast LDS.
    double D0[13], D1[13], N0[13], N1[13], A0[13], A1[13];
    double U0[3], U1[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { U0[m] = (13 + m == c0) ? 1.0 : 0.0; U1[m] = (13 + m == c1) ? 1.0 : 0.0; }
    const double h6 = a.h / 6.0;

    for (long bi = 0; bi < my_batches; ++bi) {
        const long unit0 = (blockIdx.x + bi * gridDim.x) * SENS_UNITS;
#pragma unroll
        for (int j = 0; j < 13; ++j) { D0[j] = (j == c0) ? 1.0 : 0.0; D1[j] = (j == c1) ? 1.0 : 0.0; }
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            const long t = bi * 4 + st;
            __pipeline_wait_prior(SENS_RING - 2);  // tile t landed (t+1, t+2 may still be in flight)
            __syncthreads();                       // ... and is visible; everyone is past tile t-1, so its ring slot is free
            issue(t + SENS_RING - 1);
            const double* __restrict__ T = sm.tile[t % SENS_RING] + lu * SENS_TS;
            if (st == 0) {
#pragma unroll
                for (int i = 0; i < 13; ++i) {
                    N0[i] = T[(pk0[i >> 2] >> (8 * (i & 3))) & 0xffu];
                    N1[i] = T[(pk1[i >> 2] >> (8 * (i & 3))) & 0xffu];
                }
            } else {
                // column-major traversal of the sparse Jacobian: for each input row j the (up to 13) entries J[i][j]
                // update 26 independent accumulator chains N[i][c], so consecutive DFMAs never depend on each other
#pragma unroll
                for (int i = 0; i < 13; ++i) { N0[i] = 0.0; N1[i] = 0.0; }
#pragma unroll
                for (int j = 0; j < 13; ++j) {
#pragma unroll
                    for (int i = (RIGID ? 6 : 0); i < 13; ++i) {
                        if (jx_nz(i, j, ARM)) {
                            const double jv = T[jx_slot(i, j)];   // broadcast LDS.64 (4 distinct addresses per warp)
                            N0[i] = fma(jv, D0[j], N0[i]);
                            N1[i] = fma(jv, D1[j], N1[i]);
                        }
                    }
                }
                if (!RIGID) {
#pragma unroll
                    for (int m = 0; m < 3; ++m) {
#pragma unroll
                        for (int i = 0; i < 13; ++i) {
                            if (ju_nz(i, m)) {
                                const double jv = T[ju_slot(i, m)];
                                N0[i] = fma(jv, U0[m], N0[i]);
                                N1[i] = fma(jv, U1[m], N1[i]);
                            }
                        }
                    }
                }
            }
            const double wgt = (st == 0 || st == 3) ? 1.0 : 2.0;
            const double an = (st == 2) ? a.h : 0.5 * a.h;        // a_{i+1} h, applied after stage i
#pragma unroll
            for (int i = 0; i < 13; ++i) {
                A0[i] = (st == 0) ? N0[i] : fma(wgt, N0[i], A0[i]);
                A1[i] = (st == 0) ? N1[i] : fma(wgt, N1[i], A1[i]);
                if (st < 3) {
                    D0[i] = fma(an, N0[i], (i == c0) ? 1.0 : 0.0);
                    D1[i] = fma(an, N1[i], (i == c1) ? 1.0 : 0.0);
                }
            }
        }
        // results -> shared [row r = component][unit] (padded rows) -> coalesced 256 B rows
#pragma unroll
        for (int i = 0; i < 13; ++i) {
            const double v0 = fma(h6, A0[i], (i == c0) ? 1.0 : 0.0);
            const double v1 = fma(h6, A1[i], (i == c1) ? 1.0 : 0.0);
            const int r0 = (c0 < 13) ? (i * 13 + c0) : (169 + i * 3 + (c0 - 13));
            const int r1 = (c1 < 13) ? (i * 13 + c1) : (169 + i * 3 + (c1 - 13));
            sm.out[r0 * SENS_OS + lu] = v0;
            sm.out[r1 * SENS_OS + lu] = v1;
        }
        __syncthreads();
        const int u = tid & 31;
        if (unit0 + u < a.B) {
            for (int r = tid >> 5; r < 208; r += SENS_THREADS / 32) {
                const double v = sm.out[r * SENS_OS + u];
                if (r < 169) a.Phi[(long)r * a.ld + unit0 + u] = v;
                else a.Gamma[(long)(r - 169) * a.ld + unit0 + u] = v;
            }
        }
        // the next iteration's first __syncthreads orders these reads of sm.out before it is overwritten again
    }
    __pipeline_wait_prior(0);
}

// ================================================================================================
// EKF predict (kiteEKF.cpp:75-98), two kernels.
//   A: thread per filter   -- xn = RK4(x,u,dt) and Jx(x,u) at the PRE-step state -> compact scratch
//   B: 16 lanes per filter -- Q = P A^T (lane j owns row j), transposed through shared memory,
//                             Pn = A Q + W (lane j owns column j); A = I + dt Jx, sparse broadcast loads.
// ================================================================================================
struct EkfArgs {
    KiteConsts K;
    long B, ld;
    double dt;
    const double* x; const double* u; const double* P;
    double* xn; double* Pn;
    double* Jw;              // [JAC_SLOTS][ld]
    const double* W;         // device [169]
};

template <bool RIGID>
__global__ void __launch_bounds__(128) k_ekf_state_jac(const __grid_constant__ EkfArgs a) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.B) return;
    double x[13], u[3], f[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) x[c] = __ldg(a.x + (long)c * a.ld + i);
#pragma unroll
    for (int c = 0; c < 3; ++c) u[c] = a.u ? __ldg(a.u + (long)c * a.ld + i) : 0.0;
    {
        CompactSink s{a.Jw + i, a.ld};
        model_eval<RIGID, true>(a.K, a.K.A, x, u, f, s);
    }
    rk4_step<RIGID>(a.K, a.K.A, x, u, a.dt);
#pragma unroll
    for (int c = 0; c < 13; ++c) a.xn[(long)c * a.ld + i] = x[c];
}

template <bool ARM, bool RIGID>
__global__ void __launch_bounds__(256) k_ekf_cov(const __grid_constant__ EkfArgs a) {
    __shared__ double tile[16][13][14];            // [unit in CTA][row][col], padded
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long unit = t >> 4;
    const int j = (int)(t & 15);
    const int lu = (int)(threadIdx.x >> 4);
    const bool active = (unit < a.B) && (j < 13);
    const double* __restrict__ J = a.Jw + (unit < a.B ? unit : 0);
    double pr[13], q[13];
    if (active) {
#pragma unroll
        for (int k = 0; k < 13; ++k) pr[k] = __ldg(a.P + (long)(j * 13 + k) * a.ld + unit);   // row j of P
        // q = A pr = pr + dt Jx pr    (row j of Q = P A^T)
#pragma unroll
        for (int i = 0; i < 13; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 13; ++k)
                if (jx_nz(i, k, ARM) && !(RIGID && i < 6)) s = fma(__ldg(J + (long)jx_slot(i, k) * a.ld), pr[k], s);
            q[i] = fma(a.dt, s, pr[i]);
        }
#pragma unroll
        for (int i = 0; i < 13; ++i) tile[lu][j][i] = q[i];
    }
    __syncwarp();
    if (active) {
#pragma unroll
        for (int k = 0; k < 13; ++k) pr[k] = tile[lu][k][j];      // column j of Q
#pragma unroll
        for (int i = 0; i < 13; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 13; ++k)
                if (jx_nz(i, k, ARM) && !(RIGID && i < 6)) s = fma(__ldg(J + (long)jx_slot(i, k) * a.ld), pr[k], s);
            a.Pn[(long)(i * 13 + j) * a.ld + unit] = fma(a.dt, s, pr[i]) + __ldg(a.W + i * 13 + j);
        }
    }
}

// EKF measurement update with H = [0_{7x6} I_7] (kiteEKF.cpp:115-125), thread per filter, out of place for P.
struct EkfUpdArgs {
    long B, ld;
    const double* z; const double* P; double* x; double* Pout;
    const double* V;         // device [49]
};
template <int DUMMY = 0>
__global__ void __launch_bounds__(128) k_ekf_update(const __grid_constant__ EkfUpdArgs a) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.B) return;
    // S = P[6:13,6:13] + V, inverted by Gauss-Jordan (SPD: no pivoting needed)
    double S[7][7], Si[7][7];
#pragma unroll
    for (int r = 0; r < 7; ++r)
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            S[r][c] = __ldg(a.P + (long)((6 + r) * 13 + 6 + c) * a.ld + i) + __ldg(a.V + r * 7 + c);
            Si[r][c] = (r == c) ? 1.0 : 0.0;
        }
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const double inv = 1.0 / S[c][c];
#pragma unroll
        for (int k = 0; k < 7; ++k) { S[c][k] *= inv; Si[c][k] *= inv; }
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            if (r == c) continue;
            const double mlt = S[r][c];
#pragma unroll
            for (int k = 0; k < 7; ++k) { S[r][k] = fma(-mlt, S[c][k], S[r][k]); Si[r][k] = fma(-mlt, Si[c][k], Si[r][k]); }
        }
    }
    double y[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) y[k] = __ldg(a.z + (long)k * a.ld + i) - a.x[(long)(6 + k) * a.ld + i];
    for (int r = 0; r < 13; ++r) {
        double pk[7], kr[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) pk[k] = __ldg(a.P + (long)(r * 13 + 6 + k) * a.ld + i);
        double dx = 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 7; ++k) s = fma(pk[k], Si[k][c], s);
            kr[c] = s;
            dx = fma(s, y[c], dx);
        }
        a.x[(long)r * a.ld + i] += dx;
        for (int c = 0; c < 13; ++c) {
            double s = __ldg(a.P + (long)(r * 13 + c) * a.ld + i);
#pragma unroll
            for (int k = 0; k < 7; ++k) s = fma(-kr[k], __ldg(a.P + (long)((6 + k) * 13 + c) * a.ld + i), s);
            a.Pout[(long)(r * 13 + c) * a.ld + i] = s;
        }
    }
}

// ================================================================================================
// NMPC collocation constraint + Jacobian blocks (chebyshev.hpp:241-271, kiteNMPF.cpp:58-111,169-171).
//   block = 32 scenarios x M nodes (warp k <-> node k); thread (s,k) evaluates the scaled augmented RHS and
//   its Jacobian at node k of scenario s and the row block k of (compD (x) I) X.
// ================================================================================================
struct CollocArgs {
    KiteConsts K;
    long B, ld;
    int M;
    double tau;
    double sx[15], isx[15], su[4], isu[4];
    const double* compD;     // device [M][M]
    const double* z; const double* p;
    double* G; double* JX; double* JU; double* gnorm;
};

struct CollocSink {
    double* jxp; double* jup; long ld;
    const double* sx; const double* isx; const double* isu;
    __device__ __forceinline__ void jx(int i, int j, double v) const { if (jxp) jxp[(long)(i * 15 + j) * ld] = sx[i] * v * isx[j]; }
    __device__ __forceinline__ void ju(int i, int j, double v) const { if (jup) jup[(long)(i * 4 + j) * ld] = sx[i] * v * isu[j]; }
};

template <bool PERCOEF, int NPB>
__global__ void __launch_bounds__(32 * NPB) k_colloc_eval(const __grid_constant__ CollocArgs a) {
    __shared__ double red[NPB][32];                // partial ||G||^2 per node row of the block
    const int lane = threadIdx.x;
    const int ky = threadIdx.y;
    const long s = (long)blockIdx.x * 32 + lane;
    const int M = a.M;
    double g2 = 0.0;
    if (s < a.B) {
        AeroCoef A = a.K.A;
        if constexpr (PERCOEF) load_coef(a.K, a.p, a.ld, s, A);
        for (int k = ky; k < M; k += NPB) {
            double x[13], u[3], f[13];
            const double* zx = a.z + (long)(k * 15) * a.ld + s;
            const double* zu = a.z + (long)(M * 15 + k * 4) * a.ld + s;
#pragma unroll
            for (int c = 0; c < 13; ++c) x[c] = a.isx[c] * __ldg(zx + (long)c * a.ld);
            const double x14 = a.isx[14] * __ldg(zx + 14L * a.ld);
#pragma unroll
            for (int c = 0; c < 3; ++c) u[c] = a.isu[c] * __ldg(zu + (long)c * a.ld);
            const double u3 = a.isu[3] * __ldg(zu + 3L * a.ld);
            CollocSink sink{a.JX ? a.JX + (long)(k * 225) * a.ld + s : nullptr,
                            a.JU ? a.JU + (long)(k * 60) * a.ld + s : nullptr, a.ld, a.sx, a.isx, a.isu};
            kite_eval<true>(a.K, A, x, u, f, sink);
            // augmented rows: theta_dot = V1 (x[14]), V1_dot = u_v (u[3])   (kiteNMPF.cpp:62-73)
            if (a.JX) sink.jxp[(long)(13 * 15 + 14) * a.ld] = a.sx[13] * a.isx[14];
            if (a.JU) sink.jup[(long)(14 * 4 + 3) * a.ld] = a.sx[14] * a.isu[3];
            double fa[15];
#pragma unroll
            for (int c = 0; c < 13; ++c) fa[c] = a.sx[c] * f[c];
            fa[13] = a.sx[13] * x14;
            fa[14] = a.sx[14] * u3;
            // G_k = sum_l compD[k][l] X_l - tau f_s
            double acc[15];
#pragma unroll
            for (int c = 0; c < 15; ++c) acc[c] = 0.0;
            for (int l = 0; l < M; ++l) {
                const double dkl = __ldg(a.compD + k * M + l);
                if (dkl != 0.0) {                  // warp-uniform: k is the same for the whole warp
                    const double* zl = a.z + (long)(l * 15) * a.ld + s;
#pragma unroll
                    for (int c = 0; c < 15; ++c) acc[c] = fma(dkl, __ldg(zl + (long)c * a.ld), acc[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 15; ++c) {
                const double gv = fma(-a.tau, fa[c], acc[c]);
                a.G[(long)(k * 15 + c) * a.ld + s] = gv;
                g2 = fma(gv, gv, g2);
            }
        }
    }
    if (a.gnorm) {
        red[ky][lane] = g2;
        __syncthreads();
        if (ky == 0 && s < a.B) {
            double t = 0.0;
#pragma unroll
            for (int l = 0; l < NPB; ++l) t += red[l][lane];
            a.gnorm[s] = t;
        }
    }
}

// ================================================================================================
// FP64 FMA peak microbenchmark: 8 independent register-resident DFMA chains per thread.
// ================================================================================================
template <int DUMMY = 0>
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 0.9999999, b = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    out[(long)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
constexpr long FP64_PEAK_FMAS_PER_ITER = 64;

// Accuracy self-test of kite_math.cuh on the real MUFU seeds: which = 0 rcp, 1 rsqrt, 2 asin_poly, 3 logistic, 4 asin_sc(x, sqrt(1-x^2)).
template <int DUMMY = 0>
__global__ void __launch_bounds__(256) k_math_selftest(const double* __restrict__ x, double* __restrict__ out, long n, int which) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = x[i];
    double r;
    switch (which) {
        case 0: r = fast_rcp(a); break;
        case 1: r = fast_rsqrt(a); break;
        case 2: r = asin_poly(a); break;
        case 3: r = fast_logistic(a); break;
        default: { const double c2 = fma(-a, a, 1.0); r = asin_sc(a, c2 > 0.0 ? c2 * fast_rsqrt(c2) : 0.0); } break;
    }
    out[i] = r;
}

}  // namespace kite
