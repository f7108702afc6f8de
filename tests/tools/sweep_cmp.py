import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim
from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states
def run(npts, B, env, steps=3):
    for k in list(os.environ):
        if k.startswith("QCART_") and k != "QCART_LIB": os.environ.pop(k)
    os.environ.update(env)
    params = configs.quartic_sweep(npts)
    sim = BatchedSim(params, batch=B, seed=1)
    p0 = initial_states(params, min(B, 64), 2)
    sim.set_state(np.tile(p0, ((B + 63) // 64, 1))[:B])
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    a = torch.randint(0, 21, (B,), device="cuda", dtype=torch.int32, generator=g)
    out = sim.alloc_outputs()
    sim.step(a, out=out); torch.cuda.synchronize()
    psi = sim.get_state().copy()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): sim.step(a, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fl = 160 * 568.0 * npts + 250.0 * npts
    print("N=%d B=%d %-16s %8.3f ms  %8.0f traj/s  %.1f%% | %s" % (npts, B, env, ms, B / ms * 1e3, B / ms * 1e3 * fl / 35.86e12 * 100, sim.kernel_info()), flush=True)
    return psi
cases = [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1:] if ":" in a and "=" not in a] or [(2049, 1024), (1281, 2048)]
envs = [dict(kv.split("=") for kv in a.split()) for a in sys.argv[1:] if "=" in a] or [{}]
for npts, B in cases:
    a = run(npts, B, {"QCART_PIPE": "0", "QCART_CLUSTER": "0"})
    for env in envs:
        b = run(npts, B, env)
        print("   max rel diff", float(np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(a, axis=1))))
