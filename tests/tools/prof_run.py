"""Tiny driver for ncu: a few control steps of one task at the default launch plan."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import initial_states
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim
task = sys.argv[1] if len(sys.argv) > 1 else "quartic"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
params = configs.PRESETS[task]()
sim = BatchedSim(params, batch=B, seed=1)
psi0 = initial_states(params, min(B, 256), 1)
sim.set_state(np.tile(psi0, ((B + 255) // 256, 1))[:B])
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = torch.randint(0, params["n_levels"], (B,), device="cuda", dtype=torch.int32, generator=g)
out = sim.alloc_outputs()
for _ in range(steps):
    sim.step(act, out=out)
torch.cuda.synchronize()
print(sim.kernel_info(), float(out["aux"][:, 3].sub(1).abs().max()))
