"""Accuracy/speed of the truncated implicit solve as a function of the decay tolerance (QCART_SOLVE_TOL, read at library load)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import TASKS, oracle_for, initial_states, oracle_control_step
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim
for task in TASKS:
    params = configs.PRESETS[task]()
    n_sub = params["n_sub"]
    B = 16
    rng = np.random.default_rng(5)
    psi0 = initial_states(params, B, 1)
    actions = rng.integers(0, params["n_levels"], B).astype(np.int32)
    noise = rng.standard_normal((B, n_sub, 2))
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
    g = sim.get_state()
    ref, _, _ = oracle_control_step(oracle_for(params), params, psi0, actions, noise)
    err = np.linalg.norm(g - ref, axis=1) / np.linalg.norm(ref, axis=1)
    # speed at the bench batch
    Bb = 1024 if task == "quartic" else 8192
    sim2 = BatchedSim(params, batch=Bb)
    sim2.set_state(np.tile(psi0, (Bb // B, 1)))
    act = torch.randint(0, 21, (Bb,), device="cuda", dtype=torch.int32)
    out = sim2.alloc_outputs()
    for _ in range(3): sim2.step(act, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    n = 10
    for _ in range(n): sim2.step(act, out=out)
    e1.record(); torch.cuda.synchronize()
    print("tol=%s %-18s psi err max %.2e  | B=%d %.3f ms/step | %s" % (os.environ.get("QCART_SOLVE_TOL", "1e-18"), task, err.max(), Bb, e0.elapsed_time(e1) / n, sim2.kernel_info()[-60:]), flush=True)
