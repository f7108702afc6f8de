"""Multi-process, multi-GPU test of the fused result exchange (needs >= 2 GPUs on the box; skipped otherwise).  The single-process
variant with several sims acting as ranks on one device is tests/test_gpu_api.py::test_fused_result_exchange_single_process_ranks."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
def test_fused_result_exchange_across_processes():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(HERE, "mp_fused_gather.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "FUSED_GATHER_OK world %d" % world in res.stdout, (res.stdout[-500:], res.stderr[-1500:])
