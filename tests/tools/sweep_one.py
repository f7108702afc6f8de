"""Development aid: time one grid size of the sweep.   python tests/tools/sweep_one.py N B [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim, measure_peaks
from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states
npts, B = int(sys.argv[1]), int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
fp64_peak, _ = measure_peaks(0)
params = configs.quartic_sweep(npts)
sim = BatchedSim(params, batch=B, seed=1)
p0 = initial_states(params, min(B, 64), 2)
sim.set_state(np.tile(p0, ((B + 63) // 64, 1))[:B])
g = torch.Generator(device="cuda"); g.manual_seed(0)
a = torch.randint(0, 21, (B,), device="cuda", dtype=torch.int32, generator=g)
out = sim.alloc_outputs()
for _ in range(2): sim.step(a, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): sim.step(a, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
flops = params["n_sub"] * 568.0 * npts + 250.0 * npts
print("N=%d B=%d %.3f ms/step  %.1f %% of FP64 peak  normdev %.1e | %s" % (npts, B, ms, 100 * B / (ms * 1e-3) * flops / fp64_peak, float((out["aux"][:, 3] - 1).abs().max()), sim.kernel_info()))
