"""Development aid: cost of the HERMITIAN-descriptor term of the inverted harmonic oscillator (herm_mode 0 vs 2 = term absent), 8192 trajectories."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import initial_states
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim
B = 8192
for hm in (0, 2, 1, 0):
    params = configs.inverted_harmonic(herm_mode=hm)
    sim = BatchedSim(params, batch=B, seed=1)
    psi0 = initial_states(params, 256, 1)
    sim.set_state(np.tile(psi0, (B // 256, 1)))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    act = torch.randint(0, params["n_levels"], (B,), device="cuda", dtype=torch.int32, generator=g)
    out = sim.alloc_outputs()
    for _ in range(2): sim.step(act, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): sim.step(act, out=out)
    e1.record(); torch.cuda.synchronize()
    print("herm_mode %d  %.3f ms/step | %s" % (hm, e0.elapsed_time(e1) / 5, sim.kernel_info()), flush=True)
