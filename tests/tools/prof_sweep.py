import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim
from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states
npts, B = int(sys.argv[1]), int(sys.argv[2])
params = configs.quartic_sweep(npts)
sim = BatchedSim(params, batch=B, seed=1)
p0 = initial_states(params, min(B, 64), 2)
sim.set_state(np.tile(p0, ((B + 63) // 64, 1))[:B])
a = torch.zeros(B, dtype=torch.int32, device="cuda")
out = sim.alloc_outputs()
for _ in range(2): sim.step(a, out=out)
torch.cuda.synchronize(); print(sim.kernel_info())
