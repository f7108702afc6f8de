"""Development aid: compare two launch configurations of the same control step bit-for-bit-ish (max relative state difference, moments)
and time both.   python tests/tools/cmp_pipe.py iq:8192 "QCART_PIPE=0" "QCART_PIPE=1"   """
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import initial_states
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim
TASK = {"iq": "inverted_quartic", "q": "quartic", "ih": "inverted_harmonic", "h": "harmonic"}

def run(task, B, env, steps, nctl=2):
    for k in list(os.environ):
        if k.startswith("QCART_") and k != "QCART_LIB": os.environ.pop(k)
    os.environ.update(dict(kv.split("=") for kv in env.split()) if env.strip() else {})
    params = configs.PRESETS[task]()
    sim = BatchedSim(params, batch=B, seed=1)
    psi0 = initial_states(params, min(B, 256), 1)
    sim.set_state(np.tile(psi0, ((B + 255) // 256, 1))[:B])
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    out = sim.alloc_outputs()
    for c in range(nctl):
        act = torch.randint(0, params["n_levels"], (B,), device="cuda", dtype=torch.int32, generator=g)
        sim.step(act, out=out)
    torch.cuda.synchronize()
    psi = sim.get_state().copy(); mom = out["moments"].cpu().numpy().copy(); aux = out["aux"].cpu().numpy().copy(); fl = out["flags"].cpu().numpy().copy()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): sim.step(act, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("%-22s %8.3f ms/step %9.0f traj-steps/s | %s" % (env, ms, B / ms * 1e3, sim.kernel_info()), flush=True)
    return psi, mom, aux, fl

if __name__ == "__main__":
    spec = sys.argv[1].split(":"); task = TASK.get(spec[0], spec[0]); B = int(spec[1]); steps = int(spec[2]) if len(spec) > 2 else 5
    ref = None
    for env in sys.argv[2:]:
        r = run(task, B, env, steps)
        if ref is None: ref = r; continue
        d = np.linalg.norm(r[0] - ref[0], axis=1) / np.linalg.norm(ref[0], axis=1)
        print("   vs first: max rel state diff %.3e  (nan: %d)  moments max abs diff %.3e  aux %.3e  flags differ %d" % (
            np.nanmax(d), int(np.isnan(d).sum()), np.nanmax(np.abs(r[1] - ref[1])), np.nanmax(np.abs(r[2] - ref[2])), int((r[3] != ref[3]).sum())), flush=True)
