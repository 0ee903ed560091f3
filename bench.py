#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native kite engine (contract: see the task prompt / DESIGN.md section 6).

Metric (BASELINE.json): batched kite RK4 state-steps/s.  One step of this bench = one pass of the hot path over the
config-2 batch: B = 1,048,576 trajectories x N = 1000 RK4 steps per GPU (weak scaling: every rank runs its own 1M
trajectories, global indices offset by rank*B, so inputs are identical under any sharding).

  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # reference arm: the CPU oracle port on all host threads
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic FP64 flops per unit of work: own op-counting oracle for the plain step
# (oracle.flop_counts(): 1888; SURVEY.md 8d survey count 1909), SURVEY.md 8d figures for the other kernels.
FLOPS_RK4_STEP = 1888.0
FLOPS_RK4_SENS_STEP = 27800.0
FLOPS_EKF_PREDICT = 13400.0
BYTES_COLLOC_SCENARIO = 8.0 * (209 + 165 + 11 * 285)      # z in, G + dense node blocks out (HBM-write bound kernel)
BYTES_EKF_FILTER = 8.0 * (13 + 3 + 169 + 13 + 169)
FP64_NOMINAL_TFLOPS = 37.2     # 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz
METRIC = "batched kite RK4 state-steps/sec"
UNIT = "state-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--traj", type=int, default=1 << 20, help="trajectories per GPU (config 2: 1,048,576)")
    ap.add_argument("--horizon", type=int, default=1000, help="RK4 steps per trajectory (config 2: 1000)")
    ap.add_argument("--h", type=float, default=1e-3)
    ap.add_argument("--e2e-traj", type=int, default=0, help="trajectories for the host-buffer e2e leg (0 = same as --traj)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sens", action="store_true", help="skip the secondary kernels (sensitivities, EKF predict, collocation)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index; self.proc = None; self.lines = []; self.lo = 0; self.hi = None

    def mark_begin(self):      # samples before this point (warm-up) are ignored
        self.lo = len(self.lines)

    def mark_end(self):
        self.hi = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.lo:self.hi]:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(orc, nthreads, traj_per_thread, horizon, h):
    n = traj_per_thread * nthreads
    t, _ = orc.bench_rollout(0, n, horizon, h, nthreads)
    return n * horizon / t, t, n


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  The reference's arithmetic engine (CasADi
    v3.0.0-rc2) is not vendored and cannot be built here, so this times the oracle port (oracle/, compiled C++ that
    restates kite.cpp / integrator.cpp:86-98) on all host threads, on bounded samples of the config-2 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle_py import Oracle, params_from_yaml
    orc = Oracle(params_from_yaml(os.path.join(ROOT, "data", "umx_radian.yaml")))
    nthreads = max(1, orc.hardware_threads())
    per_thread = 512                      # x horizon 1000 -> ~0.8 s of work per thread per step
    vals = []
    for i in range(args.warmup + args.steps):
        v, t, n = cpu_baseline(orc, nthreads, per_thread, args.horizon, args.h)
        if i >= args.warmup:
            vals.append((v, t))
    v = sum(n * args.horizon for _ in vals) / sum(t for _, t in vals)
    sample = "%d trajectories x %d RK4 steps per step (%d per thread), synthetic config-2 inputs" % (n, args.horizon, per_thread)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(t for _, t in vals) / len(vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config2: batched open-loop RK4 rollouts, %d steps, h=%g (bounded CPU sample)" % (args.horizon, args.h)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import openkite_b200 as okb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    from openkite_b200.sharding import shard_range

    B, N, h = args.traj, args.horizon, args.h
    eng = okb.Engine(okb.load_properties(os.path.join(ROOT, "data", "umx_radian.yaml")), okb.KITE, device=local)
    index0, count = shard_range(world * B, world, rank)       # weak scaling: every rank owns B trajectories
    assert count == B
    x0, u = eng.synth_inputs(B, N, index0=index0)           # inputs resident in HBM: x0 [13,B], u [N,3,B] (~25 GB)
    xf = eng.empty(13, B)
    gathered = eng.empty(world * 13 * B) if world > 1 else None
    torch.cuda.synchronize()

    fp64_peak = eng.fp64_peak(20000)                        # measured DFMA peak, burst (kernel timed alone)

    def step():
        eng.rollout(x0, u, N, h, okb.U_PER_STEP, out=xf, want_status=False)
        if world > 1:                                       # the path's only exchange: gather of final states
            dist.all_gather_into_tensor(gathered, xf.view(-1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local); sampler.start()       # started before the warm-up: nvidia-smi needs ~0.3 s to come up
    for _ in range(args.warmup):
        step()
    barrier()
    sampler.mark_begin()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ek = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0.record()
    for k in range(args.steps):
        ek[k][0].record()
        eng.rollout(x0, u, N, h, okb.U_PER_STEP, out=xf, want_status=False)
        ek[k][1].record()
        if world > 1:
            dist.all_gather_into_tensor(gathered, xf.view(-1))
    e1.record()
    barrier()
    sampler.mark_end()
    launches = eng.launch_count - l0
    ms_total = e0.elapsed_time(e1)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in ek)     # the rollout kernel alone, on its launch stream
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B * N / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (k_rk4_rollout): FP64 FMA pipe ---------------------------
    achieved_tflops = FLOPS_RK4_STEP * B * N / (kernel_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    try:     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from a committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "rollout_traffic.json")) as fh:
            for row in json.load(fh)["captures"]:
                if row["trajectories"] == B and row["rk4_steps"] == N:
                    traffic, traffic_src = row["dram_bytes_per_launch"], row["source"]
    except Exception:
        pass
    hbm_peak = 6555.8
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            hbm_peak = float(json.load(fh)["hbm_gbs"])
    except Exception:
        pass
    hbm_gbs = (24.0 * B * N + 208.0 * B) / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "fp64_fma", "kernel": "k_rk4_rollout<U_PER_STEP>", "achieved": achieved_tflops, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved_tflops / fp64_peak, "peak_source": "measured here (kite_fp64_peak DFMA microbenchmark); "
                "MEASURED_PEAKS.json has no FP64 entry", "frac_of_nominal_37.2": achieved_tflops / FP64_NOMINAL_TFLOPS,
                "flops_per_state_step": FLOPS_RK4_STEP, "kernel_ms": kernel_ms,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 24.0 * B * N + 208.0 * B,
                "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                        "note": "not the bound: 79 flop per byte of control stream"}}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "config2: batched open-loop RK4 rollouts with per-trajectory random control sequences",
                      "trajectories_per_gpu": B, "rk4_steps": N, "h": h, "kite": "umx_radian.yaml",
                      "l2": "inputs (%.1f GB of controls per pass) are larger than L2; no flush needed" % (24.0 * B * N / 1e9),
                      "sharding": "contiguous blocks of trajectories per rank, global index = rank*B + i"},
           "roofline": roofline, "clocks": clocks, "gpu_launches": launches}

    # ---- secondary kernels (rank 0): RK4 + sensitivities (config 3), EKF predict (config 5), collocation (config 4) ----
    if not args.no_sens and rank == 0:
        def timed(fn, reps):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        Bs = min(B, 1 << 20)
        xs, us = x0[:, :Bs].contiguous(), u[0, :, :Bs].contiguous()
        outs = (eng.empty(13, Bs), eng.empty(169, Bs), eng.empty(39, Bs))
        ms = timed(lambda: eng.sens_step(xs, us, 0.02, out=outs), 10)
        tf = FLOPS_RK4_SENS_STEP * Bs / (ms * 1e-3) / 1e12
        out["rk4_sens"] = {"kernel": "k_sens_fused", "units": Bs, "ms": ms, "state_steps_per_s": Bs / (ms * 1e-3), "achieved_tflops": tf,
                           "frac_of_measured_peak": tf / fp64_peak, "flops_per_unit": FLOPS_RK4_SENS_STEP,
                           "hbm_gbs_out": (169 + 39 + 13) * 8.0 * Bs / (ms * 1e-3) / 1e9}
        # config 3 shape: 100k trajectories x NMPC horizon (20 steps), chained primal + per-step [Phi | Gamma], one launch
        Br, Nr = min(B, 100000), min(N, 20)
        if Br >= 1024 and Nr >= 2:
            xr, ur = x0[:, :Br].contiguous(), u[:Nr, :, :Br].contiguous()
            ro = (eng.empty(Nr, 13, Br), eng.empty(Nr, 169, Br), eng.empty(Nr, 39, Br))
            ms = timed(lambda: eng.sens_rollout(xr, ur, 0.02, out=ro), 5)
            tf = FLOPS_RK4_SENS_STEP * Br * Nr / (ms * 1e-3) / 1e12
            out["rk4_sens_rollout"] = {"kernel": "k_sens_fused (one launch per horizon)", "trajectories": Br, "steps": Nr, "ms": ms,
                                       "state_steps_per_s": Br * Nr / (ms * 1e-3), "achieved_tflops": tf,
                                       "frac_of_measured_peak": tf / fp64_peak}
            del xr, ur, ro
        import numpy as np
        Wd = np.diag(np.array([.5, .5, .5, .5, .5, .5, .5, .1, .1, .01, .05, .05, .05]) ** 2)     # kiteEKF.cpp:6-13
        Pe = torch.from_numpy((10 * Wd).reshape(169, 1)).to(dev).expand(169, Bs).contiguous()
        eo = (outs[0], outs[1])
        ms = timed(lambda: eng.ekf_predict(xs, us, 0.0084, Pe, Wd, out=eo), 5)
        tf = FLOPS_EKF_PREDICT * Bs / (ms * 1e-3) / 1e12
        out["ekf_predict"] = {"kernel": "k_ekf_predict_tma", "units": Bs, "ms": ms, "filters_per_s": Bs / (ms * 1e-3), "achieved_tflops": tf,
                              "frac_of_measured_peak": tf / fp64_peak, "hbm_gbs": BYTES_EKF_FILTER * Bs / (ms * 1e-3) / 1e9,
                              "hbm_frac": BYTES_EKF_FILTER * Bs / (ms * 1e-3) / 1e9 / hbm_peak}
        zf = xs[6:13].contiguous()
        xu, Pu = eo[0].clone(), eo[1].clone()
        ms = timed(lambda: eng.ekf_update(zf, np.eye(7) * 1e-4, xu, Pu), 5)
        bu = 8.0 * (7 + 13 + 169) * 2
        out["ekf_update"] = {"kernel": "k_ekf_update (in place)", "units": Bs, "ms": ms, "filters_per_s": Bs / (ms * 1e-3),
                             "hbm_gbs": bu * Bs / (ms * 1e-3) / 1e9, "hbm_frac": bu * Bs / (ms * 1e-3) / 1e9 / hbm_peak}
        del xu, Pu, zf
        del outs, eo, Pe
        try:
            with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as fh:
                cg = json.load(fh)["cases"]["colloc_nmpc_P5_S2_scaled"]
            Bc, M = 65536, 11
            zc = (torch.tensor(cg["z"], dtype=torch.float64, device=dev).reshape(209, 1)
                  * (1 + 0.01 * torch.randn(209, Bc, dtype=torch.float64, device=dev))).contiguous()
            from openkite_b200.collocation import comp_diff_matrix
            compD = comp_diff_matrix(5, 2)
            co = (eng.empty(M * 15, Bc), eng.empty(M * 225, Bc), eng.empty(M * 60, Bc), eng.empty(Bc))
            ms = timed(lambda: eng.colloc_eval(zc, M, compD, 0.25, cg["sx"], cg["su"], out=co), 5)
            gbs = BYTES_COLLOC_SCENARIO * Bc / (ms * 1e-3) / 1e9
            out["colloc_eval"] = {"kernel": "k_colloc_eval", "scenarios": Bc, "ms": ms, "scenarios_per_s": Bc / (ms * 1e-3),
                                  "bound": "hbm", "achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak}
            del co, zc
        except Exception as e:      # secondary figure only
            out["colloc_eval"] = {"error": str(e)}
        del xs, us

    # ---- e2e: same metric through the C ABI with HOST buffers (H2D of inputs, D2H of results in the timed region)
    if not args.no_e2e:
        Be = args.e2e_traj or B
        need = 8.0 * Be * (13 + 3 * N + 13)
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 64e9
        while need * world > 0.5 * avail and Be > 4096:
            Be //= 2; need = 8.0 * Be * (13 + 3 * N + 13)
        x0_h = torch.empty(13, Be, dtype=torch.float64).pin_memory()
        u_h = torch.empty(N, 3, Be, dtype=torch.float64).pin_memory()
        xf_h = torch.empty(13, Be, dtype=torch.float64).pin_memory()
        x0_h.copy_(x0[:, :Be]); u_h.copy_(u[:, :, :Be])      # fill the pinned host buffers once (untimed)
        del u, x0
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        for _ in range(1):
            eng.rollout_host(x0_h, u_h, N, h, okb.U_PER_STEP, xf_h)
        barrier()
        reps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.rollout_host(x0_h, u_h, N, h, okb.U_PER_STEP, xf_h)      # synchronises internally (results on host)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        out["e2e"] = {"value": world * Be * N / dt, "unit": UNIT, "h2d_bytes_per_step": int(8 * Be * (13 + 3 * N)),
                      "d2h_bytes_per_step": int(8 * Be * 13), "trajectories_per_gpu": Be, "s_per_step": dt,
                      "api": "kite_rk4_rollout_host (pinned host SoA buffers, chunked H2D/compute/D2H pipeline)"}
    else:
        out["e2e"] = None

    # ---- CPU baseline beside it (rank 0, N=1 only): oracle port on all host threads, bounded sample ----
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        from oracle.oracle_py import Oracle, params_from_yaml
        orc = Oracle(params_from_yaml(os.path.join(ROOT, "data", "umx_radian.yaml")))
        nthreads = max(1, orc.hardware_threads())
        v, tsec, n = cpu_baseline(orc, nthreads, 2048, N, h)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port",
                               "sample": "%d trajectories x %d steps of the same synthetic workload (%.1f s wall); compiled "
                                         "straight-line C++ oracle, faster than the CasADi SX VM the reference runs" % (n, N, tsec)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
