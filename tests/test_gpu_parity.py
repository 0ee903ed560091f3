"""GPU parity tests (run on the B200 box): the CUDA engine, called through the C ABI, against
  (a) the committed sympy/mpmath golden vectors (tests/golden/golden.json) and
  (b) the C++ oracle on seeded inputs.
Tolerance (BASELINE.json north_star): |gpu - ref| <= 1e-9 * max(|ref|, 1) on final states and Jacobian entries.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def okb():
    import openkite_b200 as okb
    return okb


@pytest.fixture(scope="module")
def params(okb, yaml_path):
    return okb.load_properties(yaml_path)


@pytest.fixture(scope="module")
def eng(okb, params):
    return okb.Engine(params, okb.KITE)


def soa(a):
    """[B, n...] numpy AoS -> [n, B] CUDA SoA"""
    a = np.asarray(a, dtype=np.float64)
    a = a.reshape(a.shape[0], -1)
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()


def aos(t, *shape):
    a = t.detach().cpu().numpy().T
    return a.reshape(a.shape[0], *shape) if shape else a


# ---------------------------------------------------------------- golden vectors ------------------
def test_library_loaded_is_in_tree(okb):
    import os
    assert os.path.exists(okb.LIB_PATH) and okb.LIB_PATH.endswith("openkite_b200/libkite_b200.so")
    assert b"sm_100a" in okb.load_library().kite_version()


def test_rhs_jac_golden(eng, golden):
    names = list(golden["rhs"])
    x = soa([golden["rhs"][n]["x"] for n in names]); u = soa([golden["rhs"][n]["u"] for n in names])
    f = aos(eng.rhs(x, u)); Jx, Ju = eng.jac(x, u)
    Jx = aos(Jx, 13, 13); Ju = aos(Ju, 13, 3)
    for i, n in enumerate(names):
        c = golden["rhs"][n]
        assert_close(f[i], c["f"], RTOL, what=f"f[{n}]")
        assert_close(Jx[i], c["Jx"], RTOL, what=f"Jx[{n}]")
        assert_close(Ju[i], c["Ju"], RTOL, what=f"Ju[{n}]")


def test_aero_function_golden_and_batch(eng, okb, params, oracle, golden):
    """kite_aero_batch = Function "Aero" (kite.cpp:330): goldens, a batch against the oracle, and the id variant."""
    for name, c in golden["rhs"].items():
        F = eng.aero(soa([c["x"]]), soa([c["u"]]))
        assert_close(aos(F)[0], c["aero"], RTOL, what=f"aero[{name}]")
    B = 1031
    x = oracle.synth_x0(7, B); u = oracle.synth_controls(7, B, 1)[:, 0, :]
    assert_close(aos(eng.aero(soa(x), soa(u))), oracle.aero(x, u), RTOL, what="aero batch")
    e = okb.Engine(params, okb.KITE_ID)
    for name, c in golden["rhs_id"].items():
        F = e.aero(soa([c["x"]]), soa([c["u"]]), soa([c["p"]]))
        assert_close(aos(F)[0], c["aero"], RTOL, what=f"aero id[{name}]")
    e.close()


def test_config1_rollout_golden(eng, golden, okb):
    """BASELINE.json configs[0]: umx_radian, 10 s at dt = 1 ms, one trajectory (B = 1 semantics)."""
    c = golden["rollout_config1"]
    x0 = soa([c["x0"]]); u = soa([c["u"]])
    out = eng.rollout(x0, u, 10000, c["h"], okb.U_CONST, save_every=1000)
    assert int(out["status"][0]) == 0
    assert_close(aos(out["xf"])[0], c["states_after"]["10000"], RTOL, what="final state after 10 s")
    traj = out["traj"].cpu().numpy()[:, :, 0]
    assert_close(traj[0], c["states_after"]["1000"], RTOL, what="state after 1000 steps")
    assert_close(traj[4], c["states_after"]["5000"], RTOL, what="state after 5000 steps")


def test_rk4_sens_golden(eng, golden):
    for n, c in golden["rk4_step"].items():
        xn, Phi, Gam = eng.sens_step(soa([c["x"]]), soa([c["u"]]), c["h"])
        assert_close(aos(xn)[0], c["xn"], RTOL, what=f"xn[{n}]")
        assert_close(aos(Phi, 13, 13)[0], c["Phi"], RTOL, what=f"Phi[{n}]")
        assert_close(aos(Gam, 13, 3)[0], c["Gamma"], RTOL, what=f"Gamma[{n}]")


def test_ekf_predict_golden(eng, golden):
    c = golden["ekf_predict"]
    xn, Pn = eng.ekf_predict(soa([c["x"]]), soa([c["u"]]), c["dt"], soa([np.array(c["P"]).ravel()]), c["W"])
    assert_close(aos(xn)[0], c["xn"], RTOL, what="ekf xn")
    assert_close(aos(Pn, 13, 13)[0], c["Pn"], RTOL, what="ekf Pn")


@pytest.mark.parametrize("case", ["colloc_generics_P10_S1", "colloc_nmpc_P5_S2_scaled"])
def test_colloc_golden(eng, golden, oracle, case):
    c = golden[case]
    M = c["S"] * c["P"] + 1
    compD = np.array(golden["cheb"]["P%d_S%d" % (c["P"], c["S"])]["compD"])
    tau = (c["tf"] - c["t0"]) / (2 * c["S"])
    G, JX, JU, gn = eng.colloc_eval(soa([c["z"]]), M, compD, tau, c["sx"], c["su"])
    assert_close(aos(G)[0], c["G"], RTOL, what="G")
    assert_close(aos(JX, M, 15, 15)[0], c["JX"], RTOL, what="JX")
    assert_close(aos(JU, M, 15, 4)[0], c["JU"], RTOL, what="JU")
    assert_close(gn.cpu().numpy()[0], np.sum(np.array(c["G"]) ** 2), 1e-9, what="|G|^2")


def test_tether_arm_golden(okb, params, golden):
    c = golden["tether_arm"]
    p2 = okb.KiteParams.from_buffer_copy(params)
    p2.rx, p2.ry, p2.rz = c["tether_arm"]
    e = okb.Engine(p2, okb.KITE)
    x, u = soa([c["x"]]), soa([c["u"]])
    assert_close(aos(e.rhs(x, u))[0], c["f"], RTOL, what="f arm")
    Jx, Ju = e.jac(x, u)
    assert_close(aos(Jx, 13, 13)[0], c["Jx"], RTOL, what="Jx arm")
    xn, Phi, Gam = e.sens_step(x, u, c["h"])
    assert_close(aos(xn)[0], c["xn"], RTOL, what="xn arm")
    assert_close(aos(Phi, 13, 13)[0], c["Phi"], RTOL, what="Phi arm")
    assert_close(aos(Gam, 13, 3)[0], c["Gamma"], RTOL, what="Gamma arm")
    e.close()


def test_tether_arm_batch_vs_oracle(okb, params, yaml_path):
    """Tether-arm model (21 extra Jacobian entries, its own kernel instantiations): ragged batches against the oracle, both
    through the TMA-output path (even B) and the direct-store path (odd B); EKF predict with the arm as well."""
    from oracle.oracle_py import Oracle, params_from_yaml, PARAM_FIELDS
    arm = [0.012, -0.004, 0.021]
    p2 = okb.KiteParams.from_buffer_copy(params)
    p2.rx, p2.ry, p2.rz = arm
    prm = params_from_yaml(yaml_path)
    for key, v in zip(("rx", "ry", "rz"), arm):
        prm[PARAM_FIELDS.index(("tether", key))] = v
    orc = Oracle(prm)
    e = okb.Engine(p2, okb.KITE)
    h = 0.02
    for B in (330, 77):
        x = orc.synth_x0(9, B); u = orc.synth_controls(9, B, 1)[:, 0, :]
        xn, Phi, Gam = e.sens_step(soa(x), soa(u), h)
        rxn, rPhi, rGam = orc.rk4_sens(x, u, h)
        assert_close(aos(xn), rxn, RTOL, what="arm xn B=%d" % B)
        assert_close(aos(Phi, 13, 13), rPhi, RTOL, what="arm Phi B=%d" % B)
        assert_close(aos(Gam, 13, 3), rGam, RTOL, what="arm Gamma B=%d" % B)
        W, _ = orc.ekf_defaults()
        P = np.tile(10 * W, (B, 1, 1))
        xe, Pe = e.ekf_predict(soa(x), soa(u), 0.0084, soa(P), W)
        rxe, rPe = orc.ekf_predict(x, u, 0.0084, P, W)
        assert_close(aos(xe), rxe, RTOL, what="arm EKF xn B=%d" % B)
        assert_close(aos(Pe, 13, 13), rPe, RTOL, what="arm EKF Pn B=%d" % B)
    e.close()


def test_identification_variant_golden(okb, params, golden):
    e = okb.Engine(params, okb.KITE_ID)
    for n, c in golden["rhs_id"].items():
        x, u, p = soa([c["x"]]), soa([c["u"]]), soa([c["p"]])
        assert_close(aos(e.rhs(x, u, p))[0], c["f"], RTOL, what=f"id f[{n}]")
        Jx, Ju = e.jac(x, u, p)
        assert_close(aos(Jx, 13, 13)[0], c["Jx"], RTOL, what=f"id Jx[{n}]")
        out = e.rollout(x, u, 1, c["h"], okb.U_CONST, p=p)
        assert_close(aos(out["xf"])[0], c["xn"], RTOL, what=f"id xn[{n}]")
    e.close()


def test_rigid_body_golden(okb, params, golden):
    c = golden["rigid_body"]
    e = okb.Engine(params, okb.RIGID_BODY)
    x, u = soa([c["x"]]), soa([c["u"]])
    assert_close(aos(e.rhs(x, u))[0], c["f"], RTOL, what="rb f")
    Jx, Ju = e.jac(x, u)
    assert_close(aos(Jx, 13, 13)[0], c["Jx"], RTOL, what="rb Jx")
    assert float(Ju.abs().max()) == 0.0
    xn, Phi, Gam = e.sens_step(x, u, c["h"])
    assert_close(aos(xn)[0], c["xn"], RTOL, what="rb xn")
    assert_close(aos(Phi, 13, 13)[0], c["Phi"], RTOL, what="rb Phi")
    e.close()


# ---------------------------------------------------------------- oracle, seeded batches -----------
def test_rhs_jac_vs_oracle_batch(eng, oracle):
    B = 4099                                   # ragged: not a multiple of the block size
    x = oracle.synth_x0(0, B)
    u = oracle.synth_controls(0, B, 1)[:, 0, :]
    f = aos(eng.rhs(soa(x), soa(u))); Jx, Ju = eng.jac(soa(x), soa(u))
    assert_close(f, oracle.rhs(x, u), RTOL, what="f batch")
    jx, ju = oracle.jac(x, u)
    assert_close(aos(Jx, 13, 13), jx, RTOL, what="Jx batch")
    assert_close(aos(Ju, 13, 3), ju, RTOL, what="Ju batch")


def test_synth_inputs_bit_identical(eng, oracle):
    B, N, i0 = 1000, 7, 12345
    x0, u = eng.synth_inputs(B, N, index0=i0)
    assert np.array_equal(aos(x0), oracle.synth_x0(i0, B))
    assert np.array_equal(u.cpu().numpy().transpose(2, 0, 1), oracle.synth_controls(i0, B, N))


@pytest.mark.parametrize("mode", ["const", "per_step", "shared", "synth"])
def test_rollout_vs_oracle(eng, oracle, okb, mode):
    B, N, h = 777, 200, 1e-3
    x0 = oracle.synth_x0(5000, B)
    uall = oracle.synth_controls(5000, B, N)                # [B][N][3]
    if mode == "const":
        ref = oracle.rollout(x0, uall[:, 0, :].copy(), N, h, u_mode=0)
        out = eng.rollout(soa(x0), soa(uall[:, 0, :]), N, h, okb.U_CONST)
    elif mode == "per_step":
        ref = oracle.rollout(x0, uall, N, h, u_mode=1)
        u_d = torch.from_numpy(np.ascontiguousarray(uall.transpose(1, 2, 0))).cuda()     # [N][3][B]
        out = eng.rollout(soa(x0), u_d, N, h, okb.U_PER_STEP)
    elif mode == "shared":
        ref = oracle.rollout(x0, uall[0].copy(), N, h, u_mode=2)
        out = eng.rollout(soa(x0), torch.from_numpy(uall[0].copy()).cuda(), N, h, okb.U_SHARED)
    else:
        ref = oracle.rollout(None, None, N, h, u_mode=3, traj0=5000, n=B)
        out = eng.rollout(None, None, N, h, okb.U_SYNTH, index0=5000, B=B)
    assert int(out["status"].sum()) == 0
    assert_close(aos(out["xf"]), ref, RTOL, what=f"rollout {mode}")


def test_rollout_empty_and_single(eng, okb):
    out = eng.rollout(torch.empty(13, 0, dtype=torch.float64, device="cuda"),
                      torch.empty(3, 0, dtype=torch.float64, device="cuda"), 10, 1e-3, okb.U_CONST)
    assert out["xf"].shape == (13, 0)


def test_rollout_zero_and_one_step(eng, okb, oracle):
    """N = 0 returns the initial states untouched in every control mode (the per-step stream may be a one-row placeholder:
    nothing of it is consumed); N = 1 is the single RK4 step -- the shortest horizon of the control-prefetch logic."""
    B, h = 77, 1e-3
    x0 = oracle.synth_x0(900, B)
    uall = oracle.synth_controls(900, B, 2)                 # [B][2][3]
    u_step = torch.from_numpy(np.ascontiguousarray(uall.transpose(1, 2, 0))).cuda()      # [2][3][B]
    for mode, u in ((okb.U_CONST, soa(uall[:, 0, :])), (okb.U_PER_STEP, u_step[:1].contiguous()),
                    (okb.U_SHARED, torch.from_numpy(uall[0].copy()).cuda())):
        out = eng.rollout(soa(x0), u, 0, h, mode)
        assert np.array_equal(aos(out["xf"]), x0), "N = 0, mode %d" % mode
    out = eng.rollout(soa(x0), u_step[:1].contiguous(), 1, h, okb.U_PER_STEP)
    assert_close(aos(out["xf"]), oracle.rollout(x0, uall[:, :1, :].copy(), 1, h, u_mode=1), RTOL, what="one step")
    out = eng.rollout(soa(x0), u_step, 2, h, okb.U_PER_STEP)
    assert_close(aos(out["xf"]), oracle.rollout(x0, uall, 2, h, u_mode=1), RTOL, what="two steps")


def test_rollout_flags_nonfinite(eng, okb, oracle):
    x0 = oracle.synth_x0(0, 4)
    x0[2, 6:9] = 0.0                            # |r| = 0 -> division by zero in the tether term
    out = eng.rollout(soa(x0), soa(np.zeros((4, 3))), 3, 1e-3, okb.U_CONST)
    st = out["status"].cpu().numpy()
    ref = oracle.rollout(x0, np.zeros((4, 3)), 3, 1e-3)
    assert list(st) == [0, 0, 1, 0]
    assert np.array_equal(np.isfinite(ref).all(1), st == 0)        # same non-finite set as the oracle


def test_sharding_bitwise_identical(eng, okb):
    """1-GPU vs sharded evaluation must agree bit for bit per global trajectory index."""
    N, h = 50, 1e-3
    full = eng.rollout(None, None, N, h, okb.U_SYNTH, index0=0, B=4096)["xf"]
    parts = [eng.rollout(None, None, N, h, okb.U_SYNTH, index0=o, B=b)["xf"] for o, b in ((0, 1000), (1000, 2072), (3072, 1024))]
    assert torch.equal(full, torch.cat(parts, dim=1))


def test_rollout_host_pipeline(eng, okb, oracle):
    B, N, h = 3000, 40, 1e-3
    x0 = oracle.synth_x0(0, B); uall = oracle.synth_controls(0, B, N)
    x0_h = torch.from_numpy(np.ascontiguousarray(x0.T)).pin_memory()
    u_h = torch.from_numpy(np.ascontiguousarray(uall.transpose(1, 2, 0))).pin_memory()
    xf_h = torch.empty(13, B, dtype=torch.float64).pin_memory()
    st_h = torch.empty(B, dtype=torch.int32).pin_memory()
    eng.rollout_host(x0_h, u_h, N, h, okb.U_PER_STEP, xf_h, status_h=st_h)
    ref = oracle.rollout(x0, uall, N, h, u_mode=1)
    assert_close(xf_h.numpy().T, ref, RTOL, what="host pipeline rollout")
    assert int(st_h.sum()) == 0


def test_sens_rollout_vs_oracle(eng, oracle):
    """config 3 shape: NMPC horizon N = 10.  h = 0.02: explicit RK4 is unstable for this model at h >= 0.05
    (the oracle itself overflows), so the survey's h = 0.1 is not a usable multiple-shooting step."""
    B, N, h = 130, 10, 0.02
    x0 = oracle.synth_x0(100, B); uall = oracle.synth_controls(100, B, N)
    xs, Phi, Gam = eng.sens_rollout(soa(x0), torch.from_numpy(np.ascontiguousarray(uall.transpose(1, 2, 0))).cuda(), h)
    rxs, rPhi, rGam = oracle.rk4_sens_rollout(x0, uall, h)
    assert_close(xs.cpu().numpy().transpose(2, 0, 1), rxs, RTOL, what="sens rollout states")
    assert_close(Phi.cpu().numpy().transpose(2, 0, 1).reshape(B, N, 13, 13), rPhi, RTOL, what="Phi")
    assert_close(Gam.cpu().numpy().transpose(2, 0, 1).reshape(B, N, 13, 3), rGam, RTOL, what="Gamma")


def test_padded_leading_dimension(eng, okb, oracle):
    """ld > B: units live in the first B columns of wider SoA buffers (sub-batches of a larger allocation)."""
    import ctypes as C
    B, ld, h = 77, 128, 0.01
    x = oracle.synth_x0(3, B); u = oracle.synth_controls(3, B, 1)[:, 0, :]
    xd = torch.zeros(13, ld, dtype=torch.float64, device="cuda"); xd[:, :B] = soa(x)
    ud = torch.zeros(3, ld, dtype=torch.float64, device="cuda"); ud[:, :B] = soa(u)
    xn = torch.full((13, ld), -7.0, dtype=torch.float64, device="cuda")
    Phi = torch.full((169, ld), -7.0, dtype=torch.float64, device="cuda"); Gam = torch.full((39, ld), -7.0, dtype=torch.float64, device="cuda")
    w = eng.workspace(eng.L.kite_rk4_sens_work_bytes(B))
    p = lambda t: C.c_void_p(t.data_ptr())
    eng._use_torch_stream()
    eng._ck(eng.L.kite_rk4_sens_step(eng.ctx, B, ld, h, p(xd), p(ud), p(xn), p(Phi), p(Gam), p(w)))
    rxn, rPhi, rGam = oracle.rk4_sens(x, u, h)
    assert_close(aos(xn[:, :B]), rxn, RTOL, what="xn ld>B"); assert_close(aos(Phi[:, :B], 13, 13), rPhi, RTOL, what="Phi ld>B")
    assert_close(aos(Gam[:, :B], 13, 3), rGam, RTOL, what="Gamma ld>B")
    assert float((xn[:, B:] + 7.0).abs().max()) == 0.0 and float((Phi[:, B:] + 7.0).abs().max()) == 0.0   # padding untouched
    W, _ = oracle.ekf_defaults()
    Pd = torch.zeros(169, ld, dtype=torch.float64, device="cuda"); Pd[:, :B] = torch.from_numpy((10 * W).reshape(169, 1)).cuda()
    Pn = torch.full((169, ld), -7.0, dtype=torch.float64, device="cuda")
    Wh = np.ascontiguousarray(W)
    eng._ck(eng.L.kite_ekf_predict_batch(eng.ctx, B, ld, 0.0084, p(xd), p(ud), p(Pd), Wh.ctypes.data_as(C.c_void_p), p(xn), p(Pn), None))
    rxe, rPe = oracle.ekf_predict(x, u, 0.0084, np.tile(10 * W, (B, 1, 1)), W)
    assert_close(aos(Pn[:, :B], 13, 13), rPe, RTOL, what="EKF Pn ld>B"); assert_close(aos(xn[:, :B]), rxe, RTOL, what="EKF xn ld>B")
    assert float((Pn[:, B:] + 7.0).abs().max()) == 0.0
    xf = torch.full((13, ld), -7.0, dtype=torch.float64, device="cuda")
    eng._ck(eng.L.kite_rk4_rollout(eng.ctx, B, ld, 25, 1e-3, p(xd), p(ud), okb.U_CONST, None, p(xf), None, 0, None, None, None, 0))
    assert_close(aos(xf[:, :B]), oracle.rollout(x, u, 25, 1e-3), RTOL, what="rollout ld>B")
    # pointwise Jacobians with ld > B: the structural zeros are cleared for the B valid columns only, the padding of every
    # row (a neighbouring sub-batch of a larger [rows][ld] allocation) keeps its sentinel
    Jx = torch.full((169, ld), -7.0, dtype=torch.float64, device="cuda"); Ju = torch.full((39, ld), -7.0, dtype=torch.float64, device="cuda")
    fo = torch.full((13, ld), -7.0, dtype=torch.float64, device="cuda")
    eng._ck(eng.L.kite_jac_batch(eng.ctx, B, ld, p(xd), p(ud), None, p(Jx), p(Ju)))
    eng._ck(eng.L.kite_rhs_batch(eng.ctx, B, ld, p(xd), p(ud), None, p(fo)))
    rJx, rJu = oracle.jac(x, u)
    assert_close(aos(Jx[:, :B], 13, 13), rJx, RTOL, what="Jx ld>B"); assert_close(aos(Ju[:, :B], 13, 3), rJu, RTOL, what="Ju ld>B")
    assert_close(aos(fo[:, :B]), oracle.rhs(x, u), RTOL, what="f ld>B")
    assert float((Jx[:, B:] + 7.0).abs().max()) == 0.0 and float((Ju[:, B:] + 7.0).abs().max()) == 0.0
    assert float((fo[:, B:] + 7.0).abs().max()) == 0.0
    # a sub-batch at a column offset inside the same allocation: neighbours on both sides stay untouched
    off, B2 = 40, 30
    Jx.fill_(-7.0)
    po = lambda t: C.c_void_p(t.data_ptr() + 8 * off)
    eng._ck(eng.L.kite_jac_batch(eng.ctx, B2, ld, po(xd), po(ud), None, po(Jx), None))
    torch.cuda.synchronize()
    assert_close(aos(Jx[:, off:off + B2], 13, 13), rJx[off:off + B2], RTOL, what="Jx sub-batch")
    assert float((Jx[:, :off] + 7.0).abs().max()) == 0.0 and float((Jx[:, off + B2:] + 7.0).abs().max()) == 0.0


def test_sens_scratch_stays_inside_work_bytes(eng, oracle):
    """kite_rk4_sens_work_bytes(B) must cover every resident warp of the persistent kernel (groups are claimed
    dynamically): guard words behind the workspace stay untouched for small and ragged batches, repeatedly."""
    import ctypes as C
    for B in (1, 33, 77, 200, 6 * 32 + 5):
        nbytes = eng.L.kite_rk4_sens_work_bytes(B)
        buf = torch.full((nbytes // 8 + 4096,), -3.0, dtype=torch.float64, device="cuda")
        x = oracle.synth_x0(11, B); u = oracle.synth_controls(11, B, 1)[:, 0, :]
        xd, ud = soa(x), soa(u)
        rxn, rPhi, rGam = oracle.rk4_sens(x, u, 0.01)
        for _ in range(8):
            xn, Phi, Gam = eng.empty(13, B), eng.empty(169, B), eng.empty(39, B)
            p = lambda t: C.c_void_p(t.data_ptr())
            eng._use_torch_stream()
            eng._ck(eng.L.kite_rk4_sens_step(eng.ctx, B, B, 0.01, p(xd), p(ud), p(xn), p(Phi), p(Gam), p(buf)))
            assert float((buf[nbytes // 8:] + 3.0).abs().max()) == 0.0, "scratch overrun"
            assert_close(aos(Phi, 13, 13), rPhi, RTOL, what="Phi B=%d" % B)
        # the single-launch rollout (many more work items than groups) must stay inside the same workspace
        N = 7
        x0, uu = eng.synth_inputs(B, N)
        xs, Ps, Gs = eng.empty(N, 13, B), eng.empty(N, 169, B), eng.empty(N, 39, B)
        eng._ck(eng.L.kite_rk4_sens_rollout(eng.ctx, B, B, N, 0.01, p(x0), p(uu), p(xs), p(Ps), p(Gs), p(buf)))
        torch.cuda.synchronize()
        assert float((buf[nbytes // 8:] + 3.0).abs().max()) == 0.0, "scratch overrun (rollout)"
        ref = eng.sens_rollout(x0, uu, 0.01)
        assert torch.equal(xs, ref[0]) and torch.equal(Ps, ref[1]) and torch.equal(Gs, ref[2])


def test_sens_tma_output_matches_direct_stores(eng, oracle):
    """[Phi | Gamma] leave through TMA tensor stores when base and pitch are 16-byte aligned and B is even, through direct
    stores otherwise: both paths must give bitwise the same result, match the oracle, clip the ragged tail and leave the
    padding columns alone."""
    import ctypes as C
    h = 0.02
    for B in (1001, 4, 1002, 37, 2050):
        x = oracle.synth_x0(5, B); u = oracle.synth_controls(5, B, 1)[:, 0, :]
        rxn, rPhi, rGam = oracle.rk4_sens(x, u, h)
        res = []
        for ld in (B + (B & 1) + 6, B + 1 - (B & 1) + 6):          # even pitch (TMA), odd pitch (direct stores)
            xd = torch.zeros(13, ld, dtype=torch.float64, device="cuda"); xd[:, :B] = soa(x)
            ud = torch.zeros(3, ld, dtype=torch.float64, device="cuda"); ud[:, :B] = soa(u)
            xn = torch.full((13, ld), -7.0, dtype=torch.float64, device="cuda")
            Phi = torch.full((169, ld), -7.0, dtype=torch.float64, device="cuda")
            Gam = torch.full((39, ld), -7.0, dtype=torch.float64, device="cuda")
            w = eng.workspace(eng.L.kite_rk4_sens_work_bytes(B))
            p = lambda t: C.c_void_p(t.data_ptr())
            eng._use_torch_stream()
            eng._ck(eng.L.kite_rk4_sens_step(eng.ctx, B, ld, h, p(xd), p(ud), p(xn), p(Phi), p(Gam), p(w)))
            torch.cuda.synchronize()
            assert float((Phi[:, B:] + 7.0).abs().max()) == 0.0 and float((Gam[:, B:] + 7.0).abs().max()) == 0.0, "padding written"
            assert_close(aos(Phi[:, :B], 13, 13), rPhi, RTOL, what="Phi B=%d ld=%d" % (B, ld))
            assert_close(aos(Gam[:, :B], 13, 3), rGam, RTOL, what="Gamma B=%d ld=%d" % (B, ld))
            assert_close(aos(xn[:, :B]), rxn, RTOL, what="xn B=%d ld=%d" % (B, ld))
            res.append((Phi[:, :B].clone(), Gam[:, :B].clone()))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]), "TMA and direct-store outputs differ"


def test_sens_rollout_single_launch_equals_chained_steps(eng, oracle, okb, params):
    """kite_rk4_sens_rollout walks all (step, group) work items in ONE launch, step k of a group waiting for the state
    its step k - 1 published: the result must be bitwise what N chained single-step calls give, for batches smaller and
    larger than one wave of the persistent grid and for TMA-eligible and odd sizes."""
    h = 0.02
    for B, N in ((40000, 6), (4099, 9), (64, 25)):
        x0, u = eng.synth_inputs(B, N)
        xs, Phi, Gam = eng.sens_rollout(x0, u, h)
        xk = x0
        for k in range(N):
            xn, P1, G1 = eng.sens_step(xk, u[k].contiguous(), h)
            assert torch.equal(xn, xs[k]) and torch.equal(P1, Phi[k]) and torch.equal(G1, Gam[k]), "step %d of B=%d" % (k, B)
            xk = xn
        assert bool(torch.isfinite(Phi).all())
    # the rigid-body model takes the direct-store instantiation of the same kernel
    rb = okb.Engine(params, okb.RIGID_BODY)
    B, N = 300, 4
    x0, u = eng.synth_inputs(B, N)
    xs, Phi, Gam = rb.sens_rollout(x0, u, h)
    xk = x0
    for k in range(N):
        xn, P1, G1 = rb.sens_step(xk, u[k].contiguous(), h)
        assert torch.equal(xn, xs[k]) and torch.equal(P1, Phi[k]) and torch.equal(G1, Gam[k]), "rigid step %d" % k
        xk = xn
    rb.close()


def test_persistent_kernels_are_deterministic(eng, oracle):
    """The persistent kernels hand out work dynamically, reuse shared-memory tiles as TMA staging boxes and (EKF) update
    the covariance box in place: a missing fence or barrier would show up as run-to-run differences.  Repeated calls on
    batches around the size of one wave of the grid must be bitwise identical, and the EKF's TMA path must agree bitwise
    with nothing but itself and to 1e-9 with the direct-load kernel's layout fallback (odd pitch)."""
    W, _ = oracle.ekf_defaults()
    for B in (9472, 37890, 262144):
        x0, u = eng.synth_inputs(B, 3)
        P = torch.from_numpy((10 * W).reshape(169, 1)).cuda().expand(169, B).contiguous()
        P = P * (1.0 + 0.01 * torch.rand(169, B, dtype=torch.float64, device="cuda"))
        ref = None
        for rep in range(4):
            s1 = eng.sens_step(x0, u[0].contiguous(), 0.02)
            s2 = eng.sens_rollout(x0, u, 0.02)
            e1 = eng.ekf_predict(x0, u[0].contiguous(), 0.0084, P, W)
            cur = [t.clone() for t in (*s1, *s2, *e1)]
            if ref is None:
                ref = cur
            else:
                for a_, b_ in zip(ref, cur):
                    assert torch.equal(a_, b_), "run-to-run difference at B=%d" % B
    # TMA path (even pitch) against the direct kernel (odd pitch) on the same filters
    import ctypes as C
    B = 5000
    x0, u = eng.synth_inputs(B, 1)
    P = torch.from_numpy((10 * W).reshape(169, 1)).cuda().expand(169, B).contiguous()
    P = P * (1.0 + 0.01 * torch.rand(169, B, dtype=torch.float64, device="cuda"))
    outs = []
    for ld in (B + 2, B + 3):
        def pad(t):
            o = torch.zeros(t.shape[0], ld, dtype=torch.float64, device="cuda"); o[:, :B] = t; return o
        xd, ud, Pd = pad(x0), pad(u[0]), pad(P)
        xn = torch.full((13, ld), -7.0, dtype=torch.float64, device="cuda"); Pn = torch.full((169, ld), -7.0, dtype=torch.float64, device="cuda")
        p = lambda t: C.c_void_p(t.data_ptr())
        Wh = np.ascontiguousarray(W)
        eng._use_torch_stream()
        eng._ck(eng.L.kite_ekf_predict_batch(eng.ctx, B, ld, 0.0084, p(xd), p(ud), p(Pd), Wh.ctypes.data_as(C.c_void_p), p(xn), p(Pn), None))
        torch.cuda.synchronize()
        assert float((Pn[:, B:] + 7.0).abs().max()) == 0.0, "padding written (ld=%d)" % ld
        outs.append((xn[:, :B].clone(), Pn[:, :B].clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert_close(outs[0][1].cpu().numpy(), outs[1][1].cpu().numpy(), RTOL, what="EKF TMA vs direct")


def test_tma_paths_random_shapes(eng, oracle):
    """Random even (B, ld) shapes around the box / round / group / wave sizes: the TMA paths of the sensitivity step and of
    the EKF predict against the direct kernels (odd pitch), padding columns untouched, ragged tails clipped."""
    import ctypes as C
    rng = np.random.default_rng(2026)
    W, _ = oracle.ekf_defaults()
    Wh = np.ascontiguousarray(W)
    sizes = [2, 6, 8, 30, 34, 62, 66, 254, 258, 1022, 1026, 4094] + [int(2 * rng.integers(1, 20000)) for _ in range(6)]
    p = lambda t: C.c_void_p(t.data_ptr())
    for B in sizes:
        x0, u = eng.synth_inputs(B, 1)
        P = torch.from_numpy((10 * W).reshape(169, 1)).cuda().expand(169, B).contiguous()
        P = P * (1.0 + 0.01 * torch.rand(169, B, dtype=torch.float64, device="cuda"))
        res = []
        for ld in (B + 2 * int(rng.integers(0, 5)), B + 1 + 2 * int(rng.integers(0, 5))):       # even pitch: TMA, odd: direct
            def pad(t):
                o = torch.zeros(t.shape[0], ld, dtype=torch.float64, device="cuda"); o[:, :B] = t; return o
            xd, ud, Pd = pad(x0), pad(u[0]), pad(P)
            mk = lambda r: torch.full((r, ld), -7.0, dtype=torch.float64, device="cuda")
            xn, Phi, Gam, xe, Pn = mk(13), mk(169), mk(39), mk(13), mk(169)
            w = eng.workspace(eng.L.kite_rk4_sens_work_bytes(B))
            eng._use_torch_stream()
            eng._ck(eng.L.kite_rk4_sens_step(eng.ctx, B, ld, 0.02, p(xd), p(ud), p(xn), p(Phi), p(Gam), p(w)))
            eng._ck(eng.L.kite_ekf_predict_batch(eng.ctx, B, ld, 0.0084, p(xd), p(ud), p(Pd), Wh.ctypes.data_as(C.c_void_p), p(xe), p(Pn), None))
            torch.cuda.synchronize()
            for t in (xn, Phi, Gam, xe, Pn):
                if ld > B:
                    assert float((t[:, B:] + 7.0).abs().max()) == 0.0, "padding written (B=%d ld=%d)" % (B, ld)
            res.append([t[:, :B].clone() for t in (xn, Phi, Gam, xe, Pn)])
        for a_, b_, name in zip(res[0], res[1], ("xn", "Phi", "Gamma", "ekf xn", "ekf Pn")):
            if name == "ekf Pn":
                assert_close(a_.cpu().numpy(), b_.cpu().numpy(), RTOL, what="%s B=%d" % (name, B))
            else:
                assert torch.equal(a_, b_), "%s differs between the TMA and the direct path (B=%d)" % (name, B)


def test_sens_linearity_property(eng, oracle):
    """Size-independent property: Phi dx + Gamma du predicts the perturbed step to second order."""
    B, h = 2048, 0.02
    x = oracle.synth_x0(0, B); u = oracle.synth_controls(0, B, 1)[:, 0, :]
    rng = np.random.default_rng(1)
    dx = 1e-6 * rng.standard_normal((B, 13)); du = 1e-6 * rng.standard_normal((B, 3))
    xn, Phi, Gam = eng.sens_step(soa(x), soa(u), h)
    xn2, _, _ = eng.sens_step(soa(x + dx), soa(u + du), h)
    pred = aos(xn) + np.einsum("bij,bj->bi", aos(Phi, 13, 13), dx) + np.einsum("bij,bj->bi", aos(Gam, 13, 3), du)
    assert np.abs(pred - aos(xn2)).max() < 1e-9


def test_colloc_vs_oracle_batch_with_param_perturbation(eng, oracle, golden, yaml_path):
    """config 4 shape: NMPC collocation (P=5,S=2, nmpf_node scaling), per-scenario aero perturbation."""
    from oracle.oracle_py import params_from_yaml
    c = golden["colloc_nmpc_P5_S2_scaled"]
    B, M = 200, 11
    rng = np.random.default_rng(2)
    z = np.array(c["z"])[None, :] * (1 + 0.01 * rng.standard_normal((B, 209)))
    prm = params_from_yaml(yaml_path)
    prm_b = np.tile(prm, (B, 1))
    idx21 = [9, 10, 12, 13, 14, 15, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32]   # 21 id coefficients inside the 39-vector
    scale = 1 + 0.1 * (2 * rng.random((B, 21)) - 1)
    prm_b[:, idx21] *= scale
    p21 = prm_b[:, idx21]
    compD = oracle.cheb_compdiff(5, 2)
    G, JX, JU, gn = eng.colloc_eval(soa(z), M, compD, 0.25, c["sx"], c["su"], p=soa(p21))
    rG, rJX, rJU = oracle.colloc_eval(z, 5, 2, 0.0, 1.0, c["sx"], c["su"], prm_batch=prm_b)
    assert_close(aos(G), rG, RTOL, what="G batch")
    assert_close(aos(JX, M, 15, 15), rJX, RTOL, what="JX batch")
    assert_close(aos(JU, M, 15, 4), rJU, RTOL, what="JU batch")


def test_nmpc_cost_vs_oracle_batch(eng, okb, oracle, golden):
    """8f row 1: collocated NMPC cost + gradient (chebyshev.hpp:280-333, kiteNMPF.cpp:116-143), tilted and flat path."""
    from openkite_b200.collocation import quad_weights
    c = golden["colloc_nmpc_P5_S2_scaled"]
    B = 300
    rng = np.random.default_rng(9)
    z = np.array(c["z"])[None, :] * (1 + 0.05 * rng.standard_normal((B, 209)))
    for q, alt in (((np.cos(np.pi / 8), 0.0, np.sin(np.pi / 8), 0.0), 0.0), ((1.0, 0.0, 0.0, 0.0), 1.5)):
        cc = oracle.nmpc_cost_params(c["sx"], q_rot=q, altitude=alt)
        rcost, rgrad = oracle.colloc_cost(z, 5, 2, 0.0, 1.0, c["sx"], cc, nthreads=4)
        cp = okb.NmpcCost.defaults(c["sx"], q_rot=q, altitude=alt)
        cost, grad = eng.colloc_cost(soa(z), 5, 2, quad_weights(5), 0.25, c["sx"], cp)
        assert_close(cost.cpu().numpy(), rcost, RTOL, what="nmpc cost")
        assert_close(aos(grad), rgrad, RTOL, what="nmpc cost gradient")
    cost1, none = eng.colloc_cost(soa(z[:1]), 5, 2, quad_weights(5), 0.25, c["sx"], cp, want_grad=False)   # B = 1, no gradient
    assert none is None and abs(cost1.item() - rcost[0]) <= RTOL * abs(rcost[0])


def test_colloc_sparse_blocks_match_dense(okb, params, oracle, golden):
    """The sparse node-block output (structural non-zeros in CCS order, the storage of the reference's AugJacobian,
    kiteNMPF.cpp:169-171) carries bitwise the values of the dense blocks, and nothing else in them is non-zero -- without
    and with a tether arm, nominal and per-scenario coefficients."""
    from openkite_b200.collocation import comp_diff_matrix
    import copy
    c = golden["colloc_nmpc_P5_S2_scaled"]
    M, B = 11, 77
    rng = np.random.default_rng(12)
    z = soa(np.array(c["z"])[None, :] * (1 + 0.05 * rng.standard_normal((B, 209))))
    pnom = np.array(golden["rhs_id"]["nominal"]["p"])
    pb = soa(pnom[None, :] * (1 + 0.1 * (2 * rng.random((B, 21)) - 1)))
    compD = comp_diff_matrix(5, 2)
    for arm in (None, (0.02, -0.01, 0.03)):
        prm = copy.copy(params)
        if arm:
            prm.rx, prm.ry, prm.rz = arm
        e = okb.Engine(prm, okb.KITE)
        nnz = e.colloc_nnz_per_node()
        assert nnz == (134 if arm else 113)
        rows, cols = e.colloc_sparsity()
        assert len(rows) == nnz and sorted(zip(cols, rows)) == list(zip(cols, rows))        # CCS: columns, then rows, ascending
        for p in (None, pb):
            G, JX, JU, gn = e.colloc_eval(z, M, compD, 0.25, c["sx"], c["su"], p=p)
            G2, JV, gn2 = e.colloc_eval_sparse(z, M, compD, 0.25, c["sx"], c["su"], p=p)
            assert torch.equal(G, G2) and torch.equal(gn, gn2)
            G3, none_x, none_u, gn3 = e.colloc_eval(z, M, compD, 0.25, c["sx"], c["su"], p=p, want_jac=False)   # values-only kernel
            assert none_x is None and none_u is None
            assert_close(aos(G3), aos(G), 1e-13, what="G of the values-only kernel"); assert_close(gn3.cpu().numpy(), gn.cpu().numpy(), 1e-12, what="gnorm")
            dense = torch.cat([JX.reshape(M, 15, 15, B), JU.reshape(M, 15, 4, B)], dim=2)       # [M, 15, 19, B]
            picked = dense[:, rows, cols, :]                                                    # [M, nnz, B]
            assert torch.equal(picked.reshape(M * nnz, B), JV)
            mask = torch.zeros(15, 19, dtype=torch.bool, device="cuda"); mask[rows, cols] = True
            assert float(dense[:, ~mask, :].abs().max()) == 0.0                                 # everything else is a structural zero
        e.close()


def test_status_flags_next_to_results(eng, okb, oracle, golden):
    """kite_set_status_buffer: per-unit flags of the sensitivity / EKF / collocation calls: clean
    inputs give 0; planted V -> 0, |r| -> 0 and NaN units are flagged, and only those."""
    from openkite_b200.collocation import comp_diff_matrix
    B = 200
    x = oracle.synth_x0(11, B); u = oracle.synth_controls(11, B, 2)
    x[3, 0:3] = 0.0                     # unit 3: no airspeed
    x[5, 6:9] = 0.0                     # unit 5: at the tether anchor
    x[7, 4] = np.nan                    # unit 7: poisoned
    xd, ud = soa(x), soa(u[:, 0, :])
    st = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    eng.set_status_buffer(st)
    try:
        eng.sens_step(xd, ud, 0.01)
        torch.cuda.synchronize()
        f = st.cpu().numpy()
        assert f[3] & 2 and f[5] & 4 and f[7] & 1
        clean = np.ones(B, bool); clean[[3, 5, 7]] = False
        assert not f[clean].any()
        # rollout: flags OR over the steps (the NaN unit stays flagged, step 2 of the V = 0 unit is clean but the flag stays)
        st.fill_(-1)
        eng.sens_rollout(xd, torch.from_numpy(np.ascontiguousarray(u.transpose(1, 2, 0))).cuda(), 0.01)
        torch.cuda.synchronize()
        f = st.cpu().numpy()
        assert f[3] & 2 and f[5] & 4 and f[7] & 1 and not f[clean].any()
        W, V = oracle.ekf_defaults()
        P = torch.from_numpy((10 * W).reshape(169, 1)).cuda().expand(169, B).contiguous()
        st.fill_(-1)
        xn, Pn = eng.ekf_predict(xd, ud, 0.0084, P, W)
        torch.cuda.synchronize()
        f = st.cpu().numpy()
        assert f[3] & 2 and f[5] & 4 and f[7] & 1 and not f[clean].any()
        st.fill_(-1)
        eng.ekf_update(xd[6:13].contiguous(), V, xn, Pn)
        torch.cuda.synchronize()
        f = st.cpu().numpy()
        assert f[7] == 1 and not f[clean].any()
        c = golden["colloc_nmpc_P5_S2_scaled"]
        z = np.tile(np.array(c["z"])[None, :], (B, 1))
        z[9, 3 * 15 + 0:3 * 15 + 3] = 0.0          # scenario 9, node 3: v = 0
        z[12, 7 * 15 + 2] = np.inf
        st.fill_(-1)
        eng.colloc_eval(soa(z), 11, comp_diff_matrix(5, 2), 0.25, c["sx"], c["su"])
        torch.cuda.synchronize()
        f = st.cpu().numpy()
        ok = np.ones(B, bool); ok[[9, 12]] = False
        assert f[9] & 2 and f[12] & 1 and not f[ok].any()
        st.fill_(-1)
        eng.colloc_eval_sparse(soa(z), 11, comp_diff_matrix(5, 2), 0.25, c["sx"], c["su"])
        torch.cuda.synchronize()
        assert np.array_equal(st.cpu().numpy(), f)
    finally:
        eng.set_status_buffer(None)
    st.fill_(-1)
    eng.sens_step(xd, ud, 0.01)                     # buffer switched off: untouched
    torch.cuda.synchronize()
    assert int((st != -1).sum()) == 0


def test_ekf_predict_vs_oracle_batch(eng, oracle):
    B, dt = 515, 0.0084
    x = oracle.synth_x0(0, B); u = oracle.synth_controls(0, B, 1)[:, 0, :]
    W, V = oracle.ekf_defaults()
    rng = np.random.default_rng(3)
    A = rng.standard_normal((B, 13, 13)) * 0.1
    P = 10 * W[None] + A @ A.transpose(0, 2, 1)
    xn, Pn = eng.ekf_predict(soa(x), soa(u), dt, soa(P.reshape(B, 169)), W)
    rxn, rPn = oracle.ekf_predict(x, u, dt, P, W)
    assert_close(aos(xn), rxn, RTOL, what="ekf xn batch")
    assert_close(aos(Pn, 13, 13), rPn, RTOL, what="ekf Pn batch")
    # update step
    z = x[:, 6:13] + 0.01 * rng.standard_normal((B, 7))
    xu, Pu = eng.ekf_update(soa(z), V, xn.clone(), Pn.clone())
    rxu, rPu = oracle.ekf_update(z, V, rxn, rPn)
    # Tolerance 1e-9 where the update is that well conditioned, which is MEASURED per filter: the same update in 80-bit
    # extended precision tells the oracle's own round-off (7x7 inverse of S = H P H^T + V with V down to 1e-8, then the
    # cancelling P - K H P) apart from an engine error.  Entries may exceed 1e-9 only by 16 x the oracle's own error.
    lxu, lPu = oracle.ekf_update_ld(z, V, rxn, rPn)
    own_x = np.abs(rxu - lxu) / np.maximum(np.abs(lxu), 1.0)
    own_P = np.abs(rPu - lPu) / np.maximum(np.abs(lPu), 1e-3)
    err_x = np.abs(aos(xu) - lxu) / np.maximum(np.abs(lxu), 1.0)
    err_P = np.abs(aos(Pu, 13, 13) - lPu) / np.maximum(np.abs(lPu), 1e-3)
    print("ekf update: oracle double-vs-extended max %.2e (x) %.2e (P); engine-vs-extended max %.2e / %.2e"
          % (own_x.max(), own_P.max(), err_x.max(), err_P.max()))
    assert bool((err_x <= np.maximum(RTOL, 16 * own_x.max(1, keepdims=True))).all()), "ekf update x"
    assert bool((err_P <= np.maximum(RTOL, 16 * own_P.max((1, 2), keepdims=True))).all()), "ekf update P"


def test_id_cost_rollout_vs_oracle(okb, params, oracle, golden):
    """config 5 shape: parameter samples x shared control log, id-variant RHS, fused fitting cost."""
    e = okb.Engine(params, okb.KITE_ID)
    B, N, h = 300, 100, 1e-3
    c = golden["rhs_id"]["nominal"]
    pnom = np.array(c["p"])
    rng = np.random.default_rng(4)
    p = pnom[None] * (1 + 0.1 * (2 * rng.random((B, 21)) - 1))
    p[0] = pnom
    x0 = np.array(golden["rollout_config1"]["x0"])
    k = np.arange(N)
    u = np.stack([0.1 * np.ones(N), 0.1 * np.sign(np.sin(0.37 * k)), 0.1 * np.sign(np.sin(0.23 * k + 1))], 1)   # PRBS-like
    _, ytraj = oracle.rollout(x0, u, N, h, u_mode=2, p=pnom, kind=1, want_traj=True)
    y = ytraj[0, 1:, :].copy()
    rcost, rxf = oracle.id_cost_rollout(x0, u, y, p, h)
    x0b = np.tile(x0, (B, 1))
    out = e.rollout(soa(x0b), torch.from_numpy(u.copy()).cuda(), N, h, okb.U_SHARED, p=soa(p), y=torch.from_numpy(y).cuda())
    assert_close(aos(out["xf"]), rxf, RTOL, what="id xf")
    cost = out["cost"].cpu().numpy()
    assert cost[0] < 1e-20                        # nominal parameters reproduce the measurement
    assert_close(cost, rcost, 1e-9, scale=1e-9, what="id cost")
    e.close()


def test_lean_math_accuracy(eng):
    """kite_math.cuh on the real MUFU seeds: <= 4 ulp-ish relative error over the ranges the model produces."""
    rng = np.random.default_rng(7)
    x = torch.from_numpy(np.concatenate([10.0 ** rng.uniform(-8, 8, 200000), [1.0, 2.0, 1e-4, 37.5]])).cuda()
    for which, ref in ((0, lambda a: 1.0 / a), (1, lambda a: 1.0 / np.sqrt(a))):
        got = eng.math_selftest(x, which).cpu().numpy()
        r = ref(x.cpu().numpy())
        assert np.abs(got / r - 1).max() < 1e-15, which
    xa = torch.from_numpy(rng.uniform(-0.7072, 0.7072, 200000)).cuda()
    got = eng.math_selftest(xa, 2).cpu().numpy()
    assert np.abs(got - np.arcsin(xa.cpu().numpy())).max() < 5e-16
    # the (sin, cos) pair form covers the whole range branch free; near |x| = 1 the error is that of cos = sqrt(1-x^2)
    xf = torch.from_numpy(np.concatenate([rng.uniform(-0.999, 0.999, 200000), [0.0, 0.70710678, -0.70710679, 0.9999]])).cuda()
    got = eng.math_selftest(xf, 4).cpu().numpy()
    assert np.abs(got - np.arcsin(xf.cpu().numpy())).max() < 1e-14
    mid = np.abs(xf.cpu().numpy()) < 0.95
    assert np.abs(got[mid] - np.arcsin(xf.cpu().numpy()[mid])).max() < 1e-15
    got = eng.math_selftest(xf, 5).cpu().numpy()                   # table form (identification-sweep kernels)
    assert np.abs(got - np.arcsin(xf.cpu().numpy())).max() < 1e-14
    assert np.abs(got[mid] - np.arcsin(xf.cpu().numpy()[mid])).max() < 1e-15
    xl = torch.from_numpy(np.concatenate([rng.uniform(-60, 60, 200000), [-800.0, 800.0, 0.0]])).cuda()
    xr = xl.cpu().numpy()
    ref = np.where(xr >= 0, 1 / (1 + np.exp(-np.abs(xr))), np.exp(-np.abs(xr)) / (1 + np.exp(-np.abs(xr))))
    for which in (3, 6):                                           # polynomial / 2^(j/32)-table exponential
        got = eng.math_selftest(xl, which).cpu().numpy()
        assert np.abs(got - ref).max() < 1e-15, which
        assert np.abs((got[ref > 1e-280] / ref[ref > 1e-280]) - 1).max() < 2e-15, which


def test_rollout_post_stall_fallback(eng, okb, oracle):
    """Large angles (|aoa| or |sideslip| beyond 45 deg use the complement form), backwards flight (libm path), v = 0."""
    x0 = oracle.synth_x0(0, 6)
    x0[0, 0:3] = [1.0, 0.2, 4.0]      # aoa ~ 76 deg
    x0[1, 0:3] = [-3.0, 0.1, 0.5]     # flying backwards: aoa in the second quadrant
    x0[2, 0:3] = [2.0, 3.0, 0.1]      # sideslip ~ 56 deg
    x0[3, 0:3] = [0.0, 0.0, 0.0]      # at rest: legal for the standard model (1e-4 regularisers)
    u = oracle.synth_controls(0, 6, 1)[:, 0, :]
    assert_close(aos(eng.rhs(soa(x0), soa(u))), oracle.rhs(x0, u), RTOL, what="post-stall rhs")
    out = eng.rollout(soa(x0), soa(u), 20, 1e-3, okb.U_CONST)
    assert_close(aos(out["xf"]), oracle.rollout(x0, u, 20, 1e-3), RTOL, what="post-stall rollout")


def test_error_paths(eng, okb):
    x = torch.zeros(13, 4, dtype=torch.float64, device="cuda")
    with pytest.raises(okb.KiteError):
        eng.rhs(x, None)                                      # u required
    L = okb.load_library()
    assert L.kite_rhs_batch(None, 1, 1, None, None, None, None) != 0
