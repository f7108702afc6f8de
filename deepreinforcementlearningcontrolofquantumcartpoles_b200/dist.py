"""Multi-GPU plumbing: one process per GPU, trajectories sharded over ranks, one all-gather per control step.

The reference has no distributed code: its only parallelism is one OS process per trajectory
(`num_of_actors` forked actors, quartic main_parallel.py:254-369).  Here the batch axis shards over ranks with NO data-path
collective (trajectories are independent); the only exchange is the all-gather of the per-control-step result
block [B_local, K + 4 + 1] (moments, aux, flags) to every rank (the learner reads rank 0's copy).  In-kernel Philox noise is
keyed by the GLOBAL trajectory id, so results do not depend on the number of ranks.
"""
import os


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` trajectories owned by `rank` (first `total % world` ranks get one extra)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def init_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def pack_block(moments, aux, flags):
    """[B_local, K] f64, [B_local, 4] f64, [B_local] u8 -> one contiguous [B_local, K+5] f64 block (one collective, not three)."""
    import torch
    return torch.cat([moments, aux, flags.to(torch.float64).unsqueeze(1)], dim=1).contiguous()


def unpack_block(block, K):
    import torch
    return block[:, :K], block[:, K:K + 4], block[:, K + 4].to(torch.uint8)


def all_gather_block(block, world, out=None):
    """All-gather equal-sized per-rank blocks along dim 0.  Returns the [world*B_local, C] tensor (every rank)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return block
    if out is None:
        out = torch.empty((world * block.shape[0], block.shape[1]), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, block)
    return out
