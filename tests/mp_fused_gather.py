"""Rank program of tests/test_multi_gpu.py: run under torchrun with one process per GPU.  Every rank advances its shard, the SSE kernel
stores the result rows into all ranks' gather areas over CUDA-IPC peer memory, and the block is compared with a plain NCCL all-gather."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import initial_states  # noqa: E402
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import BatchedSim, configs, dist as qdist  # noqa: E402


def main():
    rank, local_rank, world = qdist.init_process_group("nccl")
    torch.cuda.set_device(local_rank)
    params = configs.quartic(n_sub=8)
    B = 96
    sim = BatchedSim(params, batch=B, device=local_rank, seed=5, traj_offset=rank * B)
    sim.set_state(initial_states(params, B, seed=20 + rank))
    fg = qdist.FusedGather(sim, rank, world)
    g = torch.Generator(device="cuda:%d" % local_rank)
    g.manual_seed(100 + rank)
    ok = True
    refs = {}
    for step in range(1, 9):
        act = torch.randint(0, params["n_levels"], (B,), device="cuda:%d" % local_rank, dtype=torch.int32, generator=g)
        out = sim.step(act)
        refs[step] = qdist.all_gather_block(qdist.pack_block(out["moments"], out["aux"], out["flags"]), world)
        if step <= 4:                                      # plain schedule
            fg.wait()
            torch.cuda.synchronize()
            ok = ok and fg.seq() == step and bool(torch.equal(fg.block(), refs[step]))
            ok = ok and bool(torch.equal(fg.block()[rank * B:(rank + 1) * B, :sim.K], out["moments"]))
        else:                                              # overlapped schedule: consume step k-1 behind step k, no host sync in between
            fg.wait(step - 1)
            ok_dev = torch.equal(fg.block(step - 1), refs[step - 1])
            ok = ok and bool(ok_dev)
    ok = ok and fg.failed_ranks() == []
    fg.close()
    flag = torch.tensor([1 if ok else 0], device="cuda:%d" % local_rank)
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    if rank == 0:
        print("FUSED_GATHER_OK" if int(flag.item()) == 1 else "FUSED_GATHER_MISMATCH", "world", world, flush=True)
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
