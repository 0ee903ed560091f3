"""CPU, world_size 2, gloo: the N > 1 host path -- block partition of the global trajectory index, inputs keyed on the
GLOBAL index, gather in global order.  The per-rank compute here is the oracle (test infrastructure); on the GPU box the
same host logic drives the CUDA engine (bench.py, tests/test_gpu_parity.py::test_sharding_bitwise_identical)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from openkite_b200.sharding import gather_units, shard_range  # noqa: E402


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 8, 9, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                i0, n = shard_range(total, world, r)
                seen += list(range(i0, i0 + n)) if total <= 1000 else [(i0, n)]
            if total <= 1000:
                assert seen == list(range(total)), (total, world)
            else:
                assert sum(n for _, n in seen) == total and all(a[0] + a[1] == b[0] for a, b in zip(seen, seen[1:]))
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, nsteps, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle_py import Oracle, params_from_yaml
    orc = Oracle(params_from_yaml(os.path.join(ROOT, "data", "umx_radian.yaml")))
    i0, n = shard_range(total, world, rank)
    xf = orc.rollout(None, None, nsteps, 1e-3, u_mode=3, traj0=i0, n=n)            # inputs keyed on the global index
    cost = np.arange(i0, i0 + n, dtype=np.float64)                                  # stand-in per-unit scalar
    g = gather_units(torch.from_numpy(np.ascontiguousarray(xf.T)), total)           # SoA [13, n] -> [13, total]
    gc = gather_units(torch.from_numpy(cost), total)
    if rank == 0:
        np.save(os.path.join(outdir, "gathered.npy"), g.numpy())
        np.save(os.path.join(outdir, "cost.npy"), gc.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 37])     # even split and ragged last block
def test_two_rank_gather_matches_single_process(tmp_path, oracle, total):
    nsteps = 20
    mp.spawn(_worker, args=(2, _free_port(), total, nsteps, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npy")
    ref = oracle.rollout(None, None, nsteps, 1e-3, u_mode=3, traj0=0, n=total)
    assert got.shape == (13, total)
    assert np.array_equal(got, ref.T), "sharded + gathered result must be bitwise identical to the single-process one"
    assert np.array_equal(np.load(tmp_path / "cost.npy"), np.arange(total, dtype=np.float64))
