"""BASELINE.json config 5: grid-size sweep of the quartic cartpole (x_max 13, dt ~ h^2, 160 substeps per control step).
For every N the resident kernel supports: throughput (CUDA events), algorithmic TFLOP/s and HBM GB/s, and the fp64 tolerance check
against the CPU oracle on a small batch.  Writes one JSON line per N."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import initial_states, oracle_for, oracle_control_step
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim, measure_peaks, QcartError

fp64_peak, _ = measure_peaks(0)
for npts, B in ((257, 4096), (513, 4096), (641, 4096), (769, 4096), (1025, 2048), (1281, 2048), (1409, 2048), (1537, 1024), (1793, 1024), (2049, 1024), (2501, 592), (4097, 296), (8193, 148)):
    params = configs.quartic_sweep(npts)
    rec = {"N": npts, "B": B, "dt": params["dt"], "n_sub": params["n_sub"]}
    try:
        # tolerance check vs the oracle (4 trajectories, 8 substeps)
        pc = dict(params, n_sub=8)
        rng = np.random.default_rng(0)
        psi0 = initial_states(pc, 4, 1); act = np.array([0, 7, 13, 20], np.int32); noise = rng.standard_normal((4, 8, 2))
        s = BatchedSim(pc, batch=4); s.set_state(psi0)
        s.step(torch.as_tensor(act, device="cuda"), noise=torch.as_tensor(noise, device="cuda")); torch.cuda.synchronize()
        ref, _, _ = oracle_control_step(oracle_for(pc), pc, psi0, act, noise)
        got = s.get_state()
        rec["max_rel_err_vs_oracle"] = float(np.max(np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)))
        del s
        sim = BatchedSim(params, batch=B, seed=1)
        p0 = initial_states(params, min(B, 64), 2)
        sim.set_state(np.tile(p0, ((B + 63) // 64, 1))[:B])
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        a = torch.randint(0, 21, (B,), device="cuda", dtype=torch.int32, generator=g)
        out = sim.alloc_outputs()
        for _ in range(2):
            sim.step(a, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 4
        e0.record()
        for _ in range(steps):
            sim.step(a, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        flops = params["n_sub"] * 568.0 * npts + 250.0 * npts
        rec.update({"ms_per_step": ms, "traj_control_steps_per_s": B / (ms * 1e-3), "tflops_algorithmic": B / (ms * 1e-3) * flops / 1e12,
                    "frac_of_fp64_peak": B / (ms * 1e-3) * flops / fp64_peak, "hbm_GBps_algorithmic": B * (32.0 * npts + 200) / (ms * 1e-3) / 1e9,
                    "kernel": sim.kernel_info(), "norm_dev": float((out["aux"][:, 3] - 1).abs().max())})
    except QcartError as e:
        rec["unsupported"] = str(e)
    print(json.dumps(rec), flush=True)
