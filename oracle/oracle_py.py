"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see oracle/kite_oracle.hpp header).  Nothing under openkite_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libkite_oracle.so")

KITE, KITE_ID, RIGID_BODY = 0, 1, 2

# Order of oracle::Params (39 doubles).  Values come from data/umx_radian.yaml of the reference
# (kept as a fixture copy of the *numbers* in tests/golden/umx_radian.yaml).
PARAM_FIELDS = [
    ("geometry", "b"), ("geometry", "c"), ("geometry", "AR"), ("geometry", "S"),
    ("inertia", "mass"), ("inertia", "Ixx"), ("inertia", "Iyy"), ("inertia", "Izz"), ("inertia", "Ixz"),
    ("aerodynamic", "CL0"), ("aerodynamic", "CLa_total"), ("aerodynamic", "e_oswald"), ("aerodynamic", "CD0_total"),
    ("aerodynamic", "CYb"), ("aerodynamic", "Cm0"), ("aerodynamic", "Cma"), ("aerodynamic", "Cn0"),
    ("aerodynamic", "Cnb"), ("aerodynamic", "Cl0"), ("aerodynamic", "Clb"),
    ("aerodynamic", "CLq"), ("aerodynamic", "Cmq"), ("aerodynamic", "CYr"), ("aerodynamic", "Cnr"),
    ("aerodynamic", "Clr"), ("aerodynamic", "CYp"), ("aerodynamic", "Clp"), ("aerodynamic", "Cnp"),
    ("aerodynamic", "CLde"), ("aerodynamic", "CYdr"), ("aerodynamic", "Cmde"), ("aerodynamic", "Cndr"),
    ("aerodynamic", "Cldr"),
    ("tether", "Ks"), ("tether", "Kd"), ("tether", "length"), ("tether", "rx"), ("tether", "ry"), ("tether", "rz"),
]


def build(force=False):
    if force or not os.path.exists(_LIB) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
        for f in ("kite_oracle.hpp", "kite_oracle_capi.cpp")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libkite_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


def params_from_yaml(path):
    """YAML -> 39-vector.  Missing tether.rx/ry/rz default to 0."""
    import yaml

    with open(path) as fh:
        cfg = yaml.safe_load(fh)
    out = np.zeros(39)
    for i, (sec, key) in enumerate(PARAM_FIELDS):
        if sec == "tether" and key in ("rx", "ry", "rz"):
            out[i] = float(cfg.get(sec, {}).get(key, 0.0))
        else:
            out[i] = float(cfg[sec][key])
    return out


_dp = C.POINTER(C.c_double)


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class Oracle:
    def __init__(self, params39):
        self.lib = C.CDLL(build())
        self.prm = np.ascontiguousarray(params39, dtype=np.float64)
        assert self.prm.shape == (39,)
        L = self.lib
        L.orc_bench_rollout.restype = C.c_double
        L.orc_hardware_threads.restype = C.c_int

    # ---- single / batched pointwise -------------------------------------------------
    def rhs(self, x, u, p=None, kind=KITE):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
        n = x.shape[0]
        f = np.empty((n, 13))
        if p is not None:
            p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        self.lib.orc_rhs(_p(self.prm), C.c_int(kind), C.c_long(n), _p(x), _p(u), _p(p), _p(f))
        return f

    def aero(self, x, u, p=None, kind=KITE):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
        n = x.shape[0]
        F = np.empty((n, 3))
        if p is not None:
            p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        self.lib.orc_aero(_p(self.prm), C.c_int(kind), C.c_long(n), _p(x), _p(u), _p(p), _p(F))
        return F

    def jac(self, x, u, p=None, kind=KITE):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
        n = x.shape[0]
        Jx = np.empty((n, 13, 13))
        Ju = np.empty((n, 13, 3))
        if p is not None:
            p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        self.lib.orc_jac(_p(self.prm), C.c_int(kind), C.c_long(n), _p(x), _p(u), _p(p), _p(Jx), _p(Ju))
        return Jx, Ju

    def rollout(self, x0, u, nsteps, h, u_mode=0, p=None, kind=KITE, want_traj=False, nthreads=1, traj0=0, n=None):
        if u_mode == 3:
            assert n is not None
            x0a = ua = None
        else:
            x0a = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
            ua = np.ascontiguousarray(u, dtype=np.float64)
            n = x0a.shape[0]
        if p is not None:
            p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        xf = np.empty((n, 13))
        traj = np.empty((n, nsteps + 1, 13)) if want_traj else None
        self.lib.orc_rk4_rollout(_p(self.prm), C.c_int(kind), C.c_long(n), C.c_long(nsteps), C.c_double(h), _p(x0a),
                                 _p(ua), C.c_int(u_mode), _p(p), C.c_long(traj0), _p(xf), _p(traj), C.c_int(nthreads))
        return (xf, traj) if want_traj else xf

    def rk4_sens(self, x, u, h, p=None, kind=KITE):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
        n = x.shape[0]
        xn = np.empty((n, 13)); Phi = np.empty((n, 13, 13)); Gam = np.empty((n, 13, 3))
        if p is not None:
            p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        self.lib.orc_rk4_sens(_p(self.prm), C.c_int(kind), C.c_long(n), C.c_double(h), _p(x), _p(u), _p(p), _p(xn),
                              _p(Phi), _p(Gam))
        return xn, Phi, Gam

    def rk4_sens_rollout(self, x0, u, h, kind=KITE, nthreads=1):
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)          # [n][nsteps][3]
        n, nsteps = u.shape[0], u.shape[1]
        xs = np.empty((n, nsteps, 13)); Phi = np.empty((n, nsteps, 13, 13)); Gam = np.empty((n, nsteps, 13, 3))
        self.lib.orc_rk4_sens_rollout(_p(self.prm), C.c_int(kind), C.c_long(n), C.c_long(nsteps), C.c_double(h),
                                      _p(x0), _p(u), _p(xs), _p(Phi), _p(Gam), C.c_int(nthreads))
        return xs, Phi, Gam

    def colloc_eval(self, z, P, S, t0, tf, sx, su, prm_batch=None, kind=KITE, nthreads=1, want_jac=True):
        M = S * P + 1
        z = np.ascontiguousarray(np.atleast_2d(z), dtype=np.float64)
        n = z.shape[0]
        assert z.shape[1] == M * 19
        sx = np.ascontiguousarray(sx, dtype=np.float64); su = np.ascontiguousarray(su, dtype=np.float64)
        G = np.empty((n, M * 15))
        JX = np.empty((n, M, 15, 15)) if want_jac else None
        JU = np.empty((n, M, 15, 4)) if want_jac else None
        if prm_batch is not None:
            prm_batch = np.ascontiguousarray(prm_batch, dtype=np.float64)
        self.lib.orc_colloc_eval(_p(self.prm), _p(prm_batch), C.c_int(kind), C.c_int(P), C.c_int(S), C.c_double(t0),
                                 C.c_double(tf), _p(sx), _p(su), C.c_long(n), _p(z), _p(G), _p(JX), _p(JU),
                                 C.c_int(nthreads))
        return G, JX, JU

    @staticmethod
    def nmpc_cost_params(sx, vel_ref=0.05, radius=2.65, altitude=0.0, q_rot=(1.0, 0.0, 0.0, 0.0), Q=(1e3, 1e3, 1e4),
                         R=(1e-4, 1e-1, 1e-1, 1e-3), W=1e-3):
        """15 doubles {Q[3], R[4], W, vref_scaled, radius, altitude, q_rot[4]} (kiteNMPF.cpp:32-34, kiteNMPF.h:34)."""
        return np.array(list(Q) + list(R) + [W, sx[14] * vel_ref, radius, altitude] + list(q_rot), dtype=np.float64)

    def colloc_cost(self, z, P, S, t0, tf, sx, cc, nthreads=1, want_grad=True):
        M = S * P + 1
        z = np.ascontiguousarray(np.atleast_2d(z), dtype=np.float64)
        n = z.shape[0]
        assert z.shape[1] == M * 19 and len(cc) == 15
        sx = np.ascontiguousarray(sx, dtype=np.float64); cc = np.ascontiguousarray(cc, dtype=np.float64)
        cost = np.empty(n); grad = np.empty((n, M * 19)) if want_grad else None
        self.lib.orc_colloc_cost(_p(cc), C.c_int(P), C.c_int(S), C.c_double(t0), C.c_double(tf), _p(sx), C.c_long(n), _p(z),
                                 _p(cost), _p(grad), C.c_int(nthreads))
        return cost, grad

    def ekf_predict(self, x, u, dt, P, W, kind=KITE):
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        u = np.ascontiguousarray(np.atleast_2d(u), dtype=np.float64)
        n = x.shape[0]
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(n, 13, 13)
        W = np.ascontiguousarray(W, dtype=np.float64).reshape(13, 13)
        xn = np.empty((n, 13)); Pn = np.empty((n, 13, 13))
        self.lib.orc_ekf_predict(_p(self.prm), C.c_int(kind), C.c_long(n), _p(x), _p(u), C.c_double(dt), _p(P), _p(W),
                                 _p(xn), _p(Pn))
        return xn, Pn

    def ekf_update(self, z, V, x, P):
        z = np.ascontiguousarray(np.atleast_2d(z), dtype=np.float64)
        n = z.shape[0]
        x = np.array(np.atleast_2d(x), dtype=np.float64, order="C")
        P = np.array(P, dtype=np.float64, order="C").reshape(n, 13, 13)
        V = np.ascontiguousarray(V, dtype=np.float64).reshape(7, 7)
        self.lib.orc_ekf_update(C.c_long(n), _p(z), _p(V), _p(x), _p(P))
        return x, P

    def ekf_defaults(self):
        W = np.empty((13, 13)); V = np.empty((7, 7))
        self.lib.orc_ekf_defaults(_p(W), _p(V))
        return W, V

    def ekf_update_ld(self, z, V, x, P):
        """ekf_update in 80-bit extended precision (conditioning yardstick)."""
        z = np.ascontiguousarray(np.atleast_2d(z), dtype=np.float64)
        n = z.shape[0]
        x = np.array(np.atleast_2d(x), dtype=np.float64, order="C")
        P = np.array(P, dtype=np.float64, order="C").reshape(n, 13, 13)
        V = np.ascontiguousarray(V, dtype=np.float64).reshape(7, 7)
        self.lib.orc_ekf_update_ld(C.c_long(n), _p(z), _p(V), _p(x), _p(P))
        return x, P

    def id_cost_rollout(self, x0, u, y, p, h, nthreads=1, extended=False):
        x0 = np.ascontiguousarray(x0, dtype=np.float64); u = np.ascontiguousarray(u, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64); p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
        n, nsteps = p.shape[0], u.shape[0]
        cost = np.empty(n); xf = np.empty((n, 13))
        fn = self.lib.orc_id_cost_rollout_ld if extended else self.lib.orc_id_cost_rollout
        fn(_p(self.prm), C.c_long(n), C.c_long(nsteps), C.c_double(h), _p(x0), _p(u), _p(y), _p(p), _p(cost), _p(xf),
           C.c_int(nthreads))
        return cost, xf

    # ---- collocation operators ---------------------------------------------------------
    def cheb_points(self, P):
        o = np.empty(P + 1); self.lib.orc_cheb_points(C.c_int(P), _p(o)); return o

    def cheb_diff(self, P):
        o = np.empty((P + 1, P + 1)); self.lib.orc_cheb_diff(C.c_int(P), _p(o)); return o

    def cheb_weights(self, P):
        o = np.empty(P + 1); self.lib.orc_cheb_weights(C.c_int(P), _p(o)); return o

    def cheb_compdiff(self, P, S):
        m = S * P + 1
        o = np.empty((m, m)); self.lib.orc_cheb_compdiff(C.c_int(P), C.c_int(S), _p(o)); return o

    # ---- synthetic workload --------------------------------------------------------------
    def synth_x0(self, traj0, n):
        o = np.empty((n, 13)); self.lib.orc_synth_x0(C.c_long(traj0), C.c_long(n), _p(o)); return o

    def synth_controls(self, traj0, n, nsteps):
        o = np.empty((n, nsteps, 3)); self.lib.orc_synth_controls(C.c_long(traj0), C.c_long(n), C.c_long(nsteps), _p(o))
        return o

    def synth_id_params(self, traj0, n, ref21):
        ref = np.ascontiguousarray(ref21, dtype=np.float64); assert ref.shape == (21,)
        o = np.empty((n, 21)); self.lib.orc_synth_id_params(C.c_long(traj0), C.c_long(n), _p(ref), _p(o)); return o

    def flop_counts(self):
        o = (C.c_long * 8)()
        self.lib.orc_flop_counts(_p(self.prm), o)
        v = list(o)
        return {"rhs": dict(add=v[0], mul=v[1], div=v[2], special=v[3], flops=sum(v[0:4])),
                "rk4_step": dict(add=v[4], mul=v[5], div=v[6], special=v[7], flops=sum(v[4:8]))}

    def bench_rollout(self, traj0, n, nsteps, h, nthreads):
        xf = np.empty((n, 13))
        t = self.lib.orc_bench_rollout(_p(self.prm), C.c_long(traj0), C.c_long(n), C.c_long(nsteps), C.c_double(h),
                                       C.c_int(nthreads), _p(xf))
        return float(t), xf

    def hardware_threads(self):
        return int(self.lib.orc_hardware_threads())
