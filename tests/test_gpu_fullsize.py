"""GPU parity at the FULL sizes of BASELINE.json's configs, through the C ABI.

The oracle cannot redo a 10^9-step batch in seconds, so each test combines
  * a SAMPLE check: a few hundred units of the full-size result, picked by global index, recomputed by the oracle on the
    same counter-RNG inputs (inputs depend on the global index only, so any subset can be regenerated on the CPU), and
  * size-independent PROPERTIES of the whole result: non-finite set empty / identical, unit quaternions, bitwise
    invariance under re-sharding of the batch, symmetry and positive diagonal of propagated covariances, Phi -> I and
    Gamma -> 0 as h -> 0, G(z) linear in the differentiation part.
Tolerance as everywhere: |gpu - oracle| <= 1e-9 max(|oracle|, 1).
"""
import numpy as np
import pytest
import torch

from conftest import assert_close
from test_gpu_parity import aos, eng, okb, params, soa  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _sample(rng, total, n):
    idx = np.unique(np.concatenate([[0, 1, 31, 32, total - 1], rng.integers(0, total, n)]))
    return idx


def test_config2_full_size_rollout(eng, okb, oracle):
    """config 2: 1,048,576 trajectories x 1000 RK4 steps, per-trajectory random controls (device-generated, bit-identical
    to the CPU generator)."""
    B, N, h = 1 << 20, 1000, 1e-3
    out = eng.rollout(None, None, N, h, okb.U_SYNTH, index0=0, B=B)
    torch.cuda.synchronize()
    assert int(out["status"].sum()) == 0                       # no trajectory left the finite range
    xf = out["xf"]
    assert bool(torch.isfinite(xf).all())
    qn = (xf[9:13] ** 2).sum(0).sqrt()
    assert float((qn - 1).abs().max()) < 1e-6                  # lambda = -5 stabiliser keeps |q| = 1 (kite.cpp:316-317)
    # sample of global indices against the oracle
    idx = _sample(np.random.default_rng(0), B, 192)
    ref = np.stack([oracle.rollout(None, None, N, h, u_mode=3, traj0=int(i), n=1)[0] for i in idx])
    assert_close(aos(xf[:, torch.from_numpy(idx).cuda()]), ref, RTOL, what="config-2 sample")
    # re-sharding: the second half computed as its own call with index0 = B/2 must be bitwise identical
    half = eng.rollout(None, None, N, h, okb.U_SYNTH, index0=B // 2, B=B // 2)["xf"]
    assert torch.equal(half, xf[:, B // 2:])


def test_config2_explicit_controls_match_synth(eng, okb):
    """The HBM-resident control stream path (KITE_U_PER_STEP, what bench.py times) and the on-the-fly generator agree bitwise."""
    B, N, h = 1 << 18, 200, 1e-3
    x0, u = eng.synth_inputs(B, N, index0=12345)
    a = eng.rollout(x0, u, N, h, okb.U_PER_STEP)["xf"]
    b = eng.rollout(None, None, N, h, okb.U_SYNTH, index0=12345, B=B)["xf"]
    assert torch.equal(a, b)


def test_config3_full_size_sensitivities(eng, oracle):
    """config 3: 100,000 trajectories x NMPC horizon (10 shooting intervals) with Phi_k, Gamma_k per step.  h = 0.02 is the
    reference's sampling time (AlgorithmProperties, simulator 50 Hz); explicit RK4 at h = 0.1 from random states leaves
    the stability region of the tether spring and is covered separately below through the non-finite sets."""
    B, N, h = 100000, 10, 0.02
    x0, u = eng.synth_inputs(B, N)
    xs, Phi, Gam = eng.sens_rollout(x0, u, h)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(Phi).all()) and bool(torch.isfinite(Gam).all()) and bool(torch.isfinite(xs).all())
    idx = _sample(np.random.default_rng(1), B, 96)
    ti = torch.from_numpy(idx).cuda()
    x0s = aos(x0[:, ti]); us = u[:, :, ti].permute(2, 0, 1).cpu().numpy()          # [n, N, 3]
    rxs, rPhi, rGam = oracle.rk4_sens_rollout(x0s, us, h, nthreads=4)
    got_xs = xs[:, :, ti].permute(2, 0, 1).cpu().numpy()
    got_Phi = Phi[:, :, ti].permute(2, 0, 1).cpu().numpy().reshape(len(idx), N, 13, 13)
    got_Gam = Gam[:, :, ti].permute(2, 0, 1).cpu().numpy().reshape(len(idx), N, 13, 3)
    assert_close(got_xs, rxs, RTOL, what="config-3 states")
    assert_close(got_Phi, rPhi, RTOL, what="config-3 Phi")
    assert_close(got_Gam, rGam, RTOL, what="config-3 Gamma")
    # property: the chained sensitivities predict the effect of a small initial perturbation on the final state
    d = 1e-7 * torch.randn(13, B, dtype=torch.float64, device="cuda")
    xs2, _, _ = eng.sens_rollout(x0 + d, u, h)
    v = d
    for k in range(N):
        v = torch.einsum("ijb,jb->ib", Phi[k].reshape(13, 13, B), v)
    err = (xs2[-1] - xs[-1] - v).abs().amax(0)
    scale = v.abs().amax(0).clamp_min(1e-12)
    assert float(torch.quantile(err / scale, 0.999)) < 1e-3    # second-order remainder only
    # h = 0.1: trajectories that leave the stability region must be the same ones as in the oracle (per trajectory;
    # the step at which an exploding state overflows to Inf/NaN depends on rounding and is not compared)
    xs3, Phi3, _ = eng.sens_rollout(x0[:, ti].contiguous(), u[:, :, ti].contiguous(), 0.1)
    rxs3, rPhi3, _ = oracle.rk4_sens_rollout(x0s, us, 0.1, nthreads=4)
    g3 = xs3.permute(2, 0, 1).cpu().numpy()
    with np.errstate(invalid="ignore"):
        blown = lambda a: ~(np.abs(a[:, -1, :]).max(axis=1) < 1e6)       # NaN compares False -> counted as blown
        assert np.array_equal(blown(g3), blown(rxs3))
    # (with the shipped parameters every randomly started trajectory is outside the RK4 stability region at h = 0.1;
    #  the first step is still finite and must agree)
    assert_close(g3[:, 0], rxs3[:, 0], RTOL, what="config-3 first step at h = 0.1")
    assert_close(Phi3[0].permute(1, 0).cpu().numpy().reshape(-1, 13, 13), rPhi3[:, 0], RTOL, what="config-3 first Phi at h = 0.1")
    # property: h -> 0 gives Phi -> I, Gamma -> 0
    _, P0, G0 = eng.sens_step(x0[:, :4096].contiguous(), u[0, :, :4096].contiguous(), 1e-12)
    eye = torch.eye(13, dtype=torch.float64, device="cuda").reshape(169, 1)
    assert float((P0 - eye).abs().max()) < 1e-8 and float(G0.abs().max()) < 1e-8


def test_sensitivity_kernel_is_schedule_independent(eng):
    """The fused sensitivity kernel hands stage Jacobians from phase A to phase B through a per-warp L2 scratch and TMA
    bulk copies (generic -> async proxy).  A stale or torn tile would show up as a schedule-dependent result: the same
    1,048,576 units computed (a) twice in one launch each and (b) as 61 launches of odd sizes (different warp -> unit
    mapping, different scratch reuse pattern) must agree bit for bit."""
    B = 1 << 20
    x0, u = eng.synth_inputs(B, 1)
    u0 = u[0].contiguous()
    a = eng.sens_step(x0, u0, 0.02)
    b = eng.sens_step(x0, u0, 0.02)
    for ta, tb in zip(a, b):
        assert torch.equal(ta, tb)
    chunk = 17203                                            # odd, not a multiple of 32
    for lo in range(0, B, chunk):
        hi = min(B, lo + chunk)
        c = eng.sens_step(x0[:, lo:hi].contiguous(), u0[:, lo:hi].contiguous(), 0.02)
        for ta, tc in zip(a, c):
            assert torch.equal(ta[:, lo:hi], tc), (lo, hi)


def test_sens_rollout_1m_trajectories(eng, oracle):
    """The north_star shape (RK4 + sensitivity rollouts of >= 1 M trajectories; bench.py's `rk4_sens_rollout` line):
    1,048,576 trajectories x 10 steps in one launch, 18.9 GB of [Phi | Gamma].  A sample of global indices against the
    oracle, bitwise re-sharding (the second half as its own call), the semigroup property of the chained sensitivities,
    and status flags all clear."""
    B, N, h = 1 << 20, 10, 0.02
    x0, u = eng.synth_inputs(B, N)
    st = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    eng.set_status_buffer(st)
    try:
        xs, Phi, Gam = eng.sens_rollout(x0, u, h)
        torch.cuda.synchronize()
    finally:
        eng.set_status_buffer(None)
    assert int(st.abs().sum()) == 0                                        # finite everywhere, no singular evaluation point
    idx = _sample(np.random.default_rng(11), B, 64)
    ti = torch.from_numpy(idx).cuda()
    x0s = aos(x0[:, ti]); us = u[:, :, ti].permute(2, 0, 1).cpu().numpy()
    rxs, rPhi, rGam = oracle.rk4_sens_rollout(x0s, us, h, nthreads=4)
    assert_close(xs[:, :, ti].permute(2, 0, 1).cpu().numpy(), rxs, RTOL, what="1M sens rollout states")
    assert_close(Phi[:, :, ti].permute(2, 0, 1).cpu().numpy().reshape(len(idx), N, 13, 13), rPhi, RTOL, what="1M sens rollout Phi")
    assert_close(Gam[:, :, ti].permute(2, 0, 1).cpu().numpy().reshape(len(idx), N, 13, 3), rGam, RTOL, what="1M sens rollout Gamma")
    # re-sharding: the second half as its own call (what rank 1 of 2 would compute) is bitwise the same
    H = B // 2
    xs2, Phi2, Gam2 = eng.sens_rollout(x0[:, H:].contiguous(), u[:, :, H:].contiguous(), h)
    assert torch.equal(xs2, xs[:, :, H:]) and torch.equal(Phi2[N - 1], Phi[N - 1, :, H:]) and torch.equal(Gam2[0], Gam[0, :, H:])
    del xs2, Phi2, Gam2
    # semigroup: the product of the step sensitivities is the sensitivity of the whole horizon (finite difference on a slice)
    S = 4096
    d = torch.zeros(13, S, dtype=torch.float64, device="cuda"); d[2] = 1e-7
    xsd, _, _ = eng.sens_rollout((x0[:, :S] + d).contiguous(), u[:, :, :S].contiguous(), h)
    v = d
    for k in range(N):
        v = torch.einsum("ijb,jb->ib", Phi[k, :, :S].reshape(13, 13, S), v)
    err = (xsd[-1] - xs[-1, :, :S] - v).abs().amax(0)
    assert float(torch.quantile(err / v.abs().amax(0).clamp_min(1e-12), 0.999)) < 1e-3


def test_config4_full_size_collocation(eng, okb, oracle, golden):
    """config 4: 65,536 NMPC scenarios (P = 5, S = 2, nmpf_node scaling): G, dG blocks, cost and gradient."""
    from openkite_b200.collocation import comp_diff_matrix, quad_weights
    c = golden["colloc_nmpc_P5_S2_scaled"]
    B, M = 65536, 11
    g = torch.Generator(device="cuda").manual_seed(4)
    z = (torch.tensor(c["z"], dtype=torch.float64, device="cuda").reshape(209, 1)
         * (1 + 0.02 * torch.randn(209, B, dtype=torch.float64, device="cuda", generator=g))).contiguous()
    compD = comp_diff_matrix(5, 2)
    G, JX, JU, gn = eng.colloc_eval(z, M, compD, 0.25, c["sx"], c["su"])
    q = (np.cos(np.pi / 8), 0.0, np.sin(np.pi / 8), 0.0)
    cost, grad = eng.colloc_cost(z, 5, 2, quad_weights(5), 0.25, c["sx"], okb.NmpcCost.defaults(c["sx"], q_rot=q))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(G).all()) and bool(torch.isfinite(JX).all()) and bool(torch.isfinite(grad).all())
    assert_close(gn.cpu().numpy(), (G ** 2).sum(0).cpu().numpy(), 1e-12, what="||G||^2 reduction")
    idx = _sample(np.random.default_rng(2), B, 128)
    ti = torch.from_numpy(idx).cuda()
    zs = aos(z[:, ti])
    rG, rJX, rJU = oracle.colloc_eval(zs, 5, 2, 0.0, 1.0, c["sx"], c["su"], nthreads=4)
    assert_close(aos(G[:, ti]), rG, RTOL, what="config-4 G")
    assert_close(aos(JX[:, ti], M, 15, 15), rJX, RTOL, what="config-4 JX")
    assert_close(aos(JU[:, ti], M, 15, 4), rJU, RTOL, what="config-4 JU")
    rc, rg = oracle.colloc_cost(zs, 5, 2, 0.0, 1.0, c["sx"], oracle.nmpc_cost_params(c["sx"], q_rot=q), nthreads=4)
    assert_close(cost[ti].cpu().numpy(), rc, RTOL, what="config-4 cost")
    assert_close(aos(grad[:, ti]), rg, RTOL, what="config-4 cost gradient")
    # structure: the dense node blocks carry exactly the structural non-zeros (104 + 1 per 15 x 15 block, 7 + 1 per 15 x 4)
    nzx = (JX.reshape(M, 225, B)[:, :, :64] != 0).sum(1)
    nzu = (JU.reshape(M, 60, B)[:, :, :64] != 0).sum(1)
    assert int(nzx.max()) <= 105 and int(nzu.max()) <= 8


def test_config5_full_size_id_sweep_and_ekf(okb, params, oracle, golden):
    """config 5 (one GPU's shard): 1,048,576 parameter samples x 2000 steps with the fused fitting cost, and
    1,048,576 EKF predict steps (dt = 0.0084, P0 = 10 W)."""
    e = okb.Engine(params, okb.KITE_ID)
    B, N, h = 1 << 20, 2000, 1e-3
    pnom = np.array(golden["rhs_id"]["nominal"]["p"])
    # parameter samples uniform inside the bounds of kite_identification_test.cpp:127-148, keyed on the global sample
    # index (kite_synth_id_params): the shard of "rank 3 of 8" of the 8,388,608-sample sweep
    i0 = 3 * B
    p = e.synth_id_params(B, index0=i0, ref=pnom)
    assert np.array_equal(aos(p[:, :257]), oracle.synth_id_params(i0, 257, pnom))      # bit-identical to the CPU generator
    p[:, 0] = torch.from_numpy(pnom).cuda()
    x0h = np.array(golden["rollout_config1"]["x0"])
    k = np.arange(N)
    u = np.stack([0.1 * np.ones(N), 0.1 * np.sign(np.sin(0.037 * k)), 0.1 * np.sign(np.sin(0.023 * k + 1))], 1)   # PRBS-like log
    _, ytraj = oracle.rollout(x0h, u, N, h, u_mode=2, p=pnom, kind=1, want_traj=True)
    y = ytraj[0, 1:, :].copy()
    x0 = torch.from_numpy(x0h.reshape(13, 1)).cuda().expand(13, B).contiguous()
    out = e.rollout(x0, torch.from_numpy(u).cuda(), N, h, okb.U_SHARED, p=p, y=torch.from_numpy(y).cuda())
    torch.cuda.synchronize()
    cost, st = out["cost"], out["status"]
    assert float(cost[0]) < 1e-18                                  # nominal parameters reproduce the measurement log
    fin = st == 0
    assert bool(torch.isfinite(cost[fin]).all()) and float(cost[fin].min()) >= 0.0
    idx = _sample(np.random.default_rng(3), B, 96)
    ti = torch.from_numpy(idx).cuda()
    ps = aos(p[:, ti])
    rcost, rxf = oracle.id_cost_rollout(x0h, u, y, ps, h, nthreads=8)
    assert np.array_equal(np.isfinite(rxf).all(1), (st[ti] == 0).cpu().numpy())   # same non-finite set
    ok = np.isfinite(rxf).all(1)
    # Tolerance 1e-9 as everywhere.  Two seconds of flight with coefficients up to 50 % off amplify round-off, so the
    # oracle's own conditioning is MEASURED rather than assumed: the same samples in 80-bit extended precision.  A sample
    # may exceed 1e-9 only where the double-precision oracle itself is that far from the extended result (x 16).
    lcost, lxf = oracle.id_cost_rollout(x0h, u, y, ps, h, nthreads=8, extended=True)
    own_x = np.abs(rxf - lxf) / np.maximum(np.abs(lxf), 1.0)
    own_c = np.abs(rcost - lcost) / np.maximum(np.abs(lcost), 1e-6)
    gx = aos(out["xf"][:, ti]); gc = cost[ti].cpu().numpy()
    err_x = np.abs(gx - lxf) / np.maximum(np.abs(lxf), 1.0)
    err_c = np.abs(gc - lcost) / np.maximum(np.abs(lcost), 1e-6)
    print("config-5 sample: oracle double-vs-extended max %.2e (states) %.2e (cost); engine-vs-extended max %.2e / %.2e"
          % (np.nanmax(own_x[ok]), np.nanmax(own_c[ok]), np.nanmax(err_x[ok]), np.nanmax(err_c[ok])))
    assert bool((err_x[ok] <= np.maximum(RTOL, 16 * own_x[ok])).all()), "config-5 final states"
    assert bool((err_c[ok] <= np.maximum(RTOL, 16 * own_c[ok])).all()), "config-5 cost"
    well = ok & (own_x.max(1) < 1e-11)                               # well-conditioned samples: strict 1e-9 against the oracle
    assert well.sum() >= len(idx) // 2
    assert_close(gc[well], rcost[well], RTOL, scale=1e-6, what="config-5 cost")
    assert_close(gx[well], rxf[well], RTOL, what="config-5 final states")
    e.close()

    ek = okb.Engine(params, okb.KITE)
    Bk = 1 << 20
    x, uu = ek.synth_inputs(Bk, 1)
    u0 = uu[0].contiguous()
    W, V = oracle.ekf_defaults()
    P0 = torch.from_numpy((10 * W).reshape(169, 1)).cuda().expand(169, Bk).contiguous()
    xn, Pn = ek.ekf_predict(x, u0, 0.0084, P0, W)
    torch.cuda.synchronize()
    Pm = Pn.reshape(13, 13, Bk)
    assert float((Pm - Pm.transpose(0, 1)).abs().max()) < 1e-10   # A P A^T + W stays symmetric
    assert float(torch.diagonal(Pm, dim1=0, dim2=1).min()) > 0.0
    idx = _sample(np.random.default_rng(6), Bk, 128)
    ti = torch.from_numpy(idx).cuda()
    rxn, rPn = oracle.ekf_predict(aos(x[:, ti]), aos(u0[:, ti]), 0.0084, np.tile(10 * W, (len(idx), 1, 1)), W)
    assert_close(aos(xn[:, ti]), rxn, RTOL, what="config-5 EKF xn")
    assert_close(aos(Pn[:, ti], 13, 13), rPn, RTOL, what="config-5 EKF Pn")
    ek.close()
