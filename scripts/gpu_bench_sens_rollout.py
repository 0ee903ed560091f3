#!/usr/bin/env python
"""GPU experiment (not part of the product): config 3 shape, B trajectories x N steps of RK4 with per-step [Phi | Gamma]
(kite_rk4_sens_rollout).  Usage (on the GPU box): python scripts/gpu_bench_sens_rollout.py [B [N]]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import openkite_b200 as okb
FLOPS = 27800.0
eng = okb.Engine(okb.load_properties(os.path.join(ROOT, "data", "umx_radian.yaml")), okb.KITE, device=0)
peak = eng.fp64_peak(20000)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
x0, u = eng.synth_inputs(B, N)
out = (eng.empty(N, 13, B), eng.empty(N, 169, B), eng.empty(N, 39, B))
run = lambda: eng.sens_rollout(x0, u, 0.02, out=out)
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
a.record()
for _ in range(reps): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
tf = FLOPS * B * N / (ms * 1e-3) / 1e12
print("sens rollout B=%d N=%d: %.3f ms (%.1f us per step)  %.3e unit-steps/s  %.2f TF  frac %.4f (peak %.2f)" % (
    B, N, ms, 1e3 * ms / N, B * N / ms * 1e3, tf, tf / peak, peak))
