"""Worker for tests/test_multigpu.py (launched by torchrun, one process per GPU): the library's own NCCL path
(kite_comm_unique_id / kite_comm_init / kite_allgather, include/kite_b200.h) against torch.distributed's all_gather,
and bitwise identity of a sharded rollout with the single-GPU result."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import openkite_b200 as okb  # noqa: E402
from openkite_b200.sharding import gather_units, shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = okb.Engine(okb.load_properties(os.path.join(ROOT, "data", "umx_radian.yaml")), okb.KITE, device=local)
    L = eng.L
    # ---- communicator bootstrap: rank 0 creates the id, everyone receives it through the existing process group
    idbuf = C.create_string_buffer(128)
    if rank == 0:
        assert L.kite_comm_unique_id(idbuf) == 0
    t = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    idbuf = C.create_string_buffer(bytes(t.cpu().tolist()), 128)
    assert L.kite_comm_init(eng.ctx, world, rank, idbuf) == 0, L.kite_last_error(eng.ctx)
    # ---- sharded rollout: total trajectories split in contiguous blocks, inputs keyed on the global index
    total, N, h = 8192 * world, 50, 1e-3
    i0, n = shard_range(total, world, rank)
    out = eng.rollout(None, None, N, h, okb.U_SYNTH, index0=i0, B=n)
    xf = out["xf"]                                             # [13, n]
    eng._use_torch_stream()
    recv = eng.empty(world * 13 * n)
    assert L.kite_allgather(eng.ctx, C.c_void_p(xf.data_ptr()), C.c_void_p(recv.data_ptr()), 13 * n) == 0, L.kite_last_error(eng.ctx)
    torch.cuda.synchronize()
    mine = recv.view(world, 13, n).permute(1, 0, 2).reshape(13, total)
    ref = gather_units(xf, total)                              # torch.distributed path
    assert torch.equal(mine, ref), "kite_allgather differs from torch all_gather"
    if rank == 0:
        single = eng.rollout(None, None, N, h, okb.U_SYNTH, index0=0, B=total)["xf"]
        assert torch.equal(single, mine), "sharded result is not bitwise identical to the single-GPU result"
        print("MULTIGPU_OK world=%d" % world)
    assert L.kite_comm_destroy(eng.ctx) == 0
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
