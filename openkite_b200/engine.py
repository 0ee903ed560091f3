"""ctypes binding of libkite_b200.so (include/kite_b200.h) for tests and bench.py.

PyTorch is used only as plumbing: device memory (torch tensors, FP64, SoA [components, B]), the current
CUDA stream and torch.distributed.  All numerics run in the hand-written sm_100a kernels behind the C ABI.
There is no CPU fallback: constructing an Engine without the built library or without a CUDA device raises.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkite_b200.so")

KITE, KITE_ID, RIGID_BODY = 0, 1, 2
U_CONST, U_PER_STEP, U_SHARED, U_SYNTH = 0, 1, 2, 3
JAC_SLOTS = 132

PARAM_FIELDS = ["b", "c", "AR", "S", "mass", "Ixx", "Iyy", "Izz", "Ixz",
                "CL0", "CLa_total", "e_oswald", "CD0_total", "CYb", "Cm0", "Cma", "Cn0", "Cnb", "Cl0", "Clb",
                "CLq", "Cmq", "CYr", "Cnr", "Clr", "CYp", "Clp", "Cnp",
                "CLde", "CYdr", "Cmde", "Cndr", "Cldr",
                "Ks", "Kd", "tether_length", "rx", "ry", "rz"]
_YAML_KEY = {"mass": ("inertia", "mass"), "tether_length": ("tether", "length")}
_SECTION = {**{k: "geometry" for k in ("b", "c", "AR", "S")},
            **{k: "inertia" for k in ("mass", "Ixx", "Iyy", "Izz", "Ixz")},
            **{k: "tether" for k in ("Ks", "Kd", "tether_length", "rx", "ry", "rz")}}


class KiteParams(C.Structure):
    """Mirror of `struct kite_params` (include/kite_b200.h)."""
    _fields_ = [(n, C.c_double) for n in PARAM_FIELDS]

    def as_list(self):
        return [getattr(self, n) for n in PARAM_FIELDS]


def load_properties(path):
    """YAML -> KiteParams (python-side convenience; the C++ host has its own loader, include/openkite/kite.hpp).
    Absent tether.rx/ry/rz default to 0 (reference kite.cpp:71-73 reads them, the shipped YAML lacks them)."""
    import yaml

    with open(path) as fh:
        cfg = yaml.safe_load(fh)
    p = KiteParams()
    for n in PARAM_FIELDS:
        sec = _SECTION.get(n, "aerodynamic")
        key = _YAML_KEY.get(n, (sec, n))[1]
        if sec == "tether" and key in ("rx", "ry", "rz"):
            val = cfg.get("tether", {}).get(key, 0.0)
        else:
            val = cfg[sec][key]
        setattr(p, n, float(val))
    return p


class NmpcCost(C.Structure):
    """Mirror of `struct kite_nmpc_cost` (include/kite_b200.h); defaults of kiteNMPF.cpp:32-34."""
    _fields_ = [("Q", C.c_double * 3), ("R", C.c_double * 4), ("W", C.c_double), ("vref_scaled", C.c_double),
                ("path_radius", C.c_double), ("path_altitude", C.c_double), ("path_q", C.c_double * 4)]

    @classmethod
    def defaults(cls, sx, vel_ref=0.05, radius=2.65, altitude=0.0, q_rot=(1.0, 0.0, 0.0, 0.0)):
        c = cls()
        c.Q[:] = [1e3, 1e3, 1e4]; c.R[:] = [1e-4, 1e-1, 1e-1, 1e-3]; c.W = 1e-3
        c.vref_scaled = float(sx[14]) * vel_ref
        c.path_radius, c.path_altitude = radius, altitude
        c.path_q[:] = list(q_rot)
        return c


class KiteError(RuntimeError):
    pass


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KiteError("libkite_b200.so is not built (run `python -m openkite_b200.build`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip, lg, db = C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_double
    L.kite_create.argtypes = [C.POINTER(vp), C.POINTER(KiteParams), ip, ip]
    L.kite_destroy.argtypes = [vp]
    L.kite_set_stream.argtypes = [vp, vp]
    L.kite_synchronize.argtypes = [vp]
    L.kite_reset_stream.argtypes = [vp]
    L.kite_set_status_buffer.argtypes = [vp, vp]
    L.kite_last_error.argtypes = [vp]; L.kite_last_error.restype = C.c_char_p
    L.kite_version.restype = C.c_char_p
    L.kite_launch_count.argtypes = [vp]; L.kite_launch_count.restype = C.c_longlong
    L.kite_rhs_batch.argtypes = [vp, lg, lg, dp, dp, dp, dp]
    L.kite_jac_batch.argtypes = [vp, lg, lg, dp, dp, dp, dp, dp]
    L.kite_aero_batch.argtypes = [vp, lg, lg, dp, dp, dp, dp]
    L.kite_rk4_rollout.argtypes = [vp, lg, lg, lg, db, dp, dp, ip, dp, dp, dp, lg, dp, dp, dp, lg]
    L.kite_rk4_rollout_host.argtypes = [vp, lg, lg, db, dp, dp, ip, dp, dp, dp, dp, dp]
    L.kite_synth_inputs.argtypes = [vp, lg, lg, lg, lg, dp, dp]
    L.kite_synth_id_params.argtypes = [vp, lg, lg, lg, dp, dp]
    L.kite_ctx_malloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    L.kite_ctx_free.argtypes = [vp, vp]
    L.kite_rk4_sens_work_bytes.argtypes = [lg]; L.kite_rk4_sens_work_bytes.restype = C.c_size_t
    L.kite_rk4_sens_step.argtypes = [vp, lg, lg, db, dp, dp, dp, dp, dp, dp]
    L.kite_rk4_sens_rollout.argtypes = [vp, lg, lg, lg, db, dp, dp, dp, dp, dp, dp]
    L.kite_colloc_eval.argtypes = [vp, lg, lg, ip, dp, db, dp, dp, dp, dp, dp, dp, dp, dp]
    L.kite_colloc_eval_sparse.argtypes = [vp, lg, lg, ip, dp, db, dp, dp, dp, dp, dp, dp, dp]
    L.kite_colloc_nnz_per_node.argtypes = [vp]
    L.kite_colloc_sparsity.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.kite_colloc_cost.argtypes = [vp, lg, lg, ip, ip, dp, db, dp, C.POINTER(NmpcCost), dp, dp, dp]
    L.kite_ekf_work_bytes.argtypes = [lg]; L.kite_ekf_work_bytes.restype = C.c_size_t
    L.kite_ekf_predict_batch.argtypes = [vp, lg, lg, db, dp, dp, dp, dp, dp, dp, dp]
    L.kite_ekf_update_batch.argtypes = [vp, lg, lg, dp, dp, dp, dp]
    L.kite_comm_unique_id.argtypes = [C.c_char_p]
    L.kite_comm_init.argtypes = [vp, ip, ip, C.c_char_p]
    L.kite_allgather.argtypes = [vp, dp, dp, lg]
    L.kite_comm_destroy.argtypes = [vp]
    L.kite_fp64_peak.argtypes = [vp, ip, C.POINTER(db)]
    L.kite_fp64_peak_reg3.argtypes = [vp, ip, C.POINTER(db)]
    L.kite_math_selftest.argtypes = [vp, lg, dp, dp, ip]
    _lib = L
    return L


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _hostarr(a):
    import numpy as np

    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.c_void_p)


class Engine:
    """One context on one GPU.  Tensors are FP64 CUDA tensors in SoA layout [components, B] (ld = B)."""

    def __init__(self, params, model_kind=KITE, device=0):
        self.L = load_library()
        if not torch.cuda.is_available():
            raise KiteError("no CUDA device: the kite engine has no CPU fallback")
        self.params = params
        self.kind = model_kind
        self.device = torch.device("cuda", device)
        self.ctx = C.c_void_p()
        rc = self.L.kite_create(C.byref(self.ctx), C.byref(params), model_kind, device)
        if rc != 0:
            raise KiteError("kite_create failed with status %d" % rc)
        self._work = None

    def close(self):
        if self.ctx:
            self.L.kite_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise KiteError("status %d: %s" % (rc, self.L.kite_last_error(self.ctx).decode()))

    def _use_torch_stream(self):
        self.L.kite_set_stream(self.ctx, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))

    def _chk(self, t, rows, B=None):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), "need contiguous FP64 CUDA tensor"
        assert t.shape[0] == rows, (tuple(t.shape), rows)
        if B is not None:
            assert t.shape[-1] == B

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def workspace(self, nbytes):
        if self._work is None or self._work.numel() * 8 < nbytes:
            self._work = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=self.device)
        return self._work

    def set_status_buffer(self, t):
        """int32 CUDA tensor [B] (or None): per-unit kite_status_flag words of the sensitivity / EKF / collocation calls."""
        assert t is None or (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous())
        self._status = t
        self._ck(self.L.kite_set_status_buffer(self.ctx, _ptr(t)))

    @property
    def launch_count(self):
        return int(self.L.kite_launch_count(self.ctx))

    def synchronize(self):
        self._ck(self.L.kite_synchronize(self.ctx))

    # ---- pointwise -----------------------------------------------------------------------------
    def rhs(self, x, u, p=None):
        self._use_torch_stream()
        B = x.shape[1]
        self._chk(x, 13)
        f = self.empty(13, B)
        self._ck(self.L.kite_rhs_batch(self.ctx, B, B, _ptr(x), _ptr(u), _ptr(p), _ptr(f)))
        return f

    def aero(self, x, u, p=None):
        """Body-frame aerodynamic force [3, B] (Function "Aero", kite.cpp:330)."""
        self._use_torch_stream()
        B = x.shape[1]
        F = self.empty(3, B)
        self._ck(self.L.kite_aero_batch(self.ctx, B, B, _ptr(x), _ptr(u), _ptr(p), _ptr(F)))
        return F

    def jac(self, x, u, p=None):
        self._use_torch_stream()
        B = x.shape[1]
        Jx, Ju = self.empty(169, B), self.empty(39, B)
        self._ck(self.L.kite_jac_batch(self.ctx, B, B, _ptr(x), _ptr(u), _ptr(p), _ptr(Jx), _ptr(Ju)))
        return Jx, Ju

    # ---- rollouts ------------------------------------------------------------------------------
    def rollout(self, x0, u, N, h, u_mode=U_CONST, p=None, save_every=0, y=None, want_status=True, index0=0, B=None,
                out=None, cost_out=None, status_out=None):
        """x0 [13,B]; u per u_mode ([3,B] | [N,3,B] | [N,3]); returns dict(xf, traj, cost, status)."""
        self._use_torch_stream()
        if u_mode == U_SYNTH:
            assert B is not None
        else:
            B = x0.shape[1]
        xf = out if out is not None else self.empty(13, B)
        traj = self.empty(N // save_every, 13, B) if save_every else None
        cost = (cost_out if cost_out is not None else self.empty(B)) if y is not None else None
        status = status_out if status_out is not None else (torch.empty(B, dtype=torch.int32, device=self.device) if want_status else None)
        self._ck(self.L.kite_rk4_rollout(self.ctx, B, B, N, h, _ptr(x0), _ptr(u), u_mode, _ptr(p), _ptr(xf), _ptr(traj),
                                         save_every, _ptr(y), _ptr(cost), _ptr(status), index0))
        return dict(xf=xf, traj=traj, cost=cost, status=status)

    def rollout_host(self, x0_h, u_h, N, h, u_mode, xf_h, p_h=None, y_h=None, cost_h=None, status_h=None):
        """HOST (ideally pinned) torch tensors, same SoA layouts; copies are pipelined inside the call."""
        B = x0_h.shape[1]
        self.L.kite_reset_stream(self.ctx)
        hp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        self._ck(self.L.kite_rk4_rollout_host(self.ctx, B, N, h, hp(x0_h), hp(u_h), u_mode, hp(p_h), hp(xf_h), hp(y_h),
                                              hp(cost_h), hp(status_h)))

    def synth_inputs(self, B, N, index0=0, want_u=True):
        self._use_torch_stream()
        x0 = self.empty(13, B)
        u = self.empty(max(N, 1), 3, B) if want_u else None
        self._ck(self.L.kite_synth_inputs(self.ctx, B, B, N if want_u else 0, index0, _ptr(x0), _ptr(u)))
        return x0, u

    def synth_id_params(self, B, index0=0, ref=None):
        """Parameter samples [21, B] of the identification sweep for global indices [index0, index0 + B)."""
        self._use_torch_stream()
        p = self.empty(21, B)
        ra, rp = _hostarr(ref) if ref is not None else (None, None)
        self._ck(self.L.kite_synth_id_params(self.ctx, B, B, index0, rp, _ptr(p)))
        return p

    # ---- multi-GPU: the library's own NCCL path (kite_comm_* / kite_allgather) ------------------
    def comm_init(self, world, rank):
        """Bootstrap the library communicator: rank 0 draws the NCCL id, torch.distributed (already initialised by the
        caller: plumbing) carries it to the other ranks."""
        import torch.distributed as dist
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            self._ck(self.L.kite_comm_unique_id(idbuf))
        t = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device=self.device)
        dist.broadcast(t, 0)
        idbuf = C.create_string_buffer(bytes(t.cpu().tolist()), 128)
        self._ck(self.L.kite_comm_init(self.ctx, world, rank, idbuf))
        self.world = world

    def allgather(self, send, recv):
        """recv[world * n] <- concatenation over ranks of send[n] (FP64), on the current torch stream."""
        self._use_torch_stream()
        assert send.dtype == torch.float64 and recv.dtype == torch.float64 and send.is_contiguous() and recv.is_contiguous()
        self._ck(self.L.kite_allgather(self.ctx, _ptr(send), _ptr(recv), send.numel()))
        return recv

    def comm_destroy(self):
        self._ck(self.L.kite_comm_destroy(self.ctx))

    # ---- sensitivities -------------------------------------------------------------------------
    def sens_step(self, x, u, h, out=None):
        self._use_torch_stream()
        B = x.shape[1]
        xn, Phi, Gam = out if out is not None else (self.empty(13, B), self.empty(169, B), self.empty(39, B))
        w = self.workspace(self.L.kite_rk4_sens_work_bytes(B))
        self._ck(self.L.kite_rk4_sens_step(self.ctx, B, B, h, _ptr(x), _ptr(u), _ptr(xn), _ptr(Phi), _ptr(Gam), _ptr(w)))
        return xn, Phi, Gam

    def sens_rollout(self, x0, u, h, out=None):
        self._use_torch_stream()
        N, B = u.shape[0], u.shape[2]
        if out is None:
            out = (self.empty(N, 13, B), self.empty(N, 169, B), self.empty(N, 39, B))
        xs, Phi, Gam = out
        w = self.workspace(self.L.kite_rk4_sens_work_bytes(B))
        self._ck(self.L.kite_rk4_sens_rollout(self.ctx, B, B, N, h, _ptr(x0), _ptr(u), _ptr(xs), _ptr(Phi), _ptr(Gam),
                                              _ptr(w)))
        return xs, Phi, Gam

    # ---- collocation ---------------------------------------------------------------------------
    def colloc_eval(self, z, M, compD, tau, sx, su, p=None, want_jac=True, want_norm=True, out=None):
        self._use_torch_stream()
        B = z.shape[1]
        self._chk(z, M * 19)
        cd, cdp = _hostarr(compD); sxa, sxp = _hostarr(sx); sua, sup = _hostarr(su)
        if out is not None:
            G, JX, JU, gn = out
        else:
            G = self.empty(M * 15, B)
            JX = self.empty(M * 225, B) if want_jac else None
            JU = self.empty(M * 60, B) if want_jac else None
            gn = self.empty(B) if want_norm else None
        self._ck(self.L.kite_colloc_eval(self.ctx, B, B, M, cdp, tau, sxp, sup, _ptr(z), _ptr(p), _ptr(G), _ptr(JX),
                                         _ptr(JU), _ptr(gn)))
        return G, JX, JU, gn

    def colloc_nnz_per_node(self):
        return int(self.L.kite_colloc_nnz_per_node(self.ctx))

    def colloc_sparsity(self):
        """(rows, cols) of the structural non-zeros of a 15 x 19 node block [d f_s/d x_s | d f_s/d u_s], CCS order."""
        n = self.colloc_nnz_per_node()
        r, c = (C.c_int * n)(), (C.c_int * n)()
        assert self.L.kite_colloc_sparsity(self.ctx, r, c) == n
        return list(r), list(c)

    def colloc_eval_sparse(self, z, M, compD, tau, sx, su, p=None, want_norm=True, out=None):
        """As colloc_eval, node blocks as structural non-zeros: returns G [M*15, B], JV [M*nnz, B], gnorm [B]."""
        self._use_torch_stream()
        B = z.shape[1]
        self._chk(z, M * 19)
        cd, cdp = _hostarr(compD); sxa, sxp = _hostarr(sx); sua, sup = _hostarr(su)
        if out is not None:
            G, JV, gn = out
        else:
            G, JV = self.empty(M * 15, B), self.empty(M * self.colloc_nnz_per_node(), B)
            gn = self.empty(B) if want_norm else None
        self._ck(self.L.kite_colloc_eval_sparse(self.ctx, B, B, M, cdp, tau, sxp, sup, _ptr(z), _ptr(p), _ptr(G), _ptr(JV), _ptr(gn)))
        return G, JV, gn

    def colloc_cost(self, z, P, S, qw, tau, sx, cost_params, want_grad=True, out=None):
        """NMPC performance index + gradient (kite_colloc_cost): z [M*19, B] -> cost [B], grad [M*19, B]."""
        self._use_torch_stream()
        M, B = S * P + 1, z.shape[1]
        self._chk(z, M * 19)
        qa, qp = _hostarr(qw); sxa, sxp = _hostarr(sx)
        cost, grad = out if out is not None else (self.empty(B), self.empty(M * 19, B) if want_grad else None)
        self._ck(self.L.kite_colloc_cost(self.ctx, B, B, P, S, qp, tau, sxp, C.byref(cost_params), _ptr(z), _ptr(cost), _ptr(grad)))
        return cost, grad

    # ---- EKF -----------------------------------------------------------------------------------
    def ekf_predict(self, x, u, dt, P, W, out=None):
        self._use_torch_stream()
        B = x.shape[1]
        Wa, Wp = _hostarr(W)
        xn, Pn = out if out is not None else (self.empty(13, B), self.empty(169, B))
        w = self.workspace(self.L.kite_ekf_work_bytes(B))
        self._ck(self.L.kite_ekf_predict_batch(self.ctx, B, B, dt, _ptr(x), _ptr(u), _ptr(P), Wp, _ptr(xn), _ptr(Pn),
                                               _ptr(w)))
        return xn, Pn

    def ekf_update(self, z, V, x, P):
        self._use_torch_stream()
        B = x.shape[1]
        Va, Vp = _hostarr(V)
        self._ck(self.L.kite_ekf_update_batch(self.ctx, B, B, _ptr(z), Vp, _ptr(x), _ptr(P)))
        return x, P

    # ---- diagnostics ----------------------------------------------------------------------------
    def math_selftest(self, x, which):
        self._use_torch_stream()
        out = torch.empty_like(x)
        self._ck(self.L.kite_math_selftest(self.ctx, x.numel(), _ptr(x), _ptr(out), which))
        return out

    def fp64_peak(self, iters=20000, three_register_operands=False):
        self._use_torch_stream()
        out = C.c_double()
        fn = self.L.kite_fp64_peak_reg3 if three_register_operands else self.L.kite_fp64_peak
        self._ck(fn(self.ctx, iters, C.byref(out)))
        return out.value
