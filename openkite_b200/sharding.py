"""Host-side sharding of trajectory batches over the GPUs of one box.

Units (trajectories / scenarios / parameter samples) never interact, so the only multi-GPU logic is
  * the contiguous block partition of the global index range, and
  * the gather of per-unit results (final states [13, B], costs [B]) in global-index order.
The gather uses torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests); a non-Python host uses
kite_comm_init / kite_allgather from include/kite_b200.h instead.  No collective runs inside the time loop.
"""
import torch
import torch.distributed as dist


def shard_range(total, world, rank):
    """Rank `rank` owns global indices [index0, index0 + count): blocks of ceil(total / world), last one ragged."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request: total=%r world=%r rank=%r" % (total, world, rank))
    block = -(-total // world)
    index0 = min(total, rank * block)
    return index0, max(0, min(total, index0 + block) - index0)


def gather_units(local, total, group=None):
    """all-gather SoA results: `local` is [rows, count_of_this_rank] (or [count]); returns [rows, total] (or [total])
    on every rank, columns in global-index order.  Ragged last blocks are padded to the block size for the collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    block = -(-total // world)
    squeeze = local.dim() == 1
    loc = local.unsqueeze(0) if squeeze else local
    rows = loc.shape[0]
    send = loc
    if loc.shape[1] != block:
        send = torch.zeros(rows, block, dtype=loc.dtype, device=loc.device)
        send[:, :loc.shape[1]] = loc
    recv = torch.empty(world * rows, block, dtype=loc.dtype, device=loc.device)      # rank-major concatenation
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    out = recv.view(world, rows, block).permute(1, 0, 2).reshape(rows, world * block)[:, :total].contiguous()
    return out[0] if squeeze else out
