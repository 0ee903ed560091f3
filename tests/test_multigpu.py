"""Multi-GPU path on real GPUs (needs >= 2 devices; skipped on the 1-GPU box): torchrun with one process per GPU runs
tests/multigpu_worker.py -- library NCCL gather vs torch.distributed, sharded vs single-GPU bitwise identity."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_library_nccl_gather_and_sharding():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_worker.py")]
    env = dict(os.environ, NCCL_DEBUG="INFO", NCCL_DEBUG_SUBSYS="INIT")      # rank / channel lines of the library's communicator
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    # keep the evidence: gpurun merges gpurun_out/ back, the summary is committed under profiles/
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "multigpu_test_%dgpu.log" % n), "w") as fh:
            fh.write("$ " + " ".join(cmd) + "\nrc=%d\n" % r.returncode)
            keep = [ln for ln in (r.stdout + r.stderr).splitlines()
                    if "MULTIGPU" in ln or "comm 0x" in ln or "Init COMPLETE" in ln or "nranks" in ln or "Error" in ln or "assert" in ln.lower()]
            fh.write("\n".join(keep[:400]) + "\n")
    except OSError:
        pass
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU_OK" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_two_contexts_on_two_devices_in_one_process():
    """The ABI promises one context per device and many contexts per process: every kernel family (including the ones
    that opt in to > 48 KB of dynamic shared memory and size their grids from the SM count) must run on a second device
    of the same process and give bitwise the results of device 0."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import openkite_b200 as okb
    prm = okb.load_properties(os.path.join(ROOT, "data", "umx_radian.yaml"))
    res = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            eng = okb.Engine(prm, okb.KITE, device=dev)
            B = 2048 + 6
            x0, u = eng.synth_inputs(B, 2)
            u0 = u[0].contiguous()
            xn, Phi, Gam = eng.sens_step(x0, u0, 0.02)
            W = np.diag(np.array([.5, .5, .5, .5, .5, .5, .5, .1, .1, .01, .05, .05, .05]) ** 2)
            P = torch.from_numpy((10 * W).reshape(169, 1)).to(eng.device).expand(169, B).contiguous()
            xe, Pn = eng.ekf_predict(x0, u0, 0.0084, P, W)
            xu, Pu = xe.clone(), Pn.clone()
            eng.ekf_update(x0[6:13].contiguous(), np.eye(7) * 1e-4, xu, Pu)
            xf = eng.rollout(x0, u0, 20, 1e-3, okb.U_CONST)["xf"]
            torch.cuda.synchronize(dev)
            res.append([t.cpu() for t in (xn, Phi, Gam, xe, Pn, xu, Pu, xf)])
            eng.close()
    for a, b in zip(*res):
        assert torch.equal(a, b)
