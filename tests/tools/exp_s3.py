"""Development aid (session 3 of round 2): (a) start offset between the two warps of a scheduler for one-warp trajectories (QCART_STAGGER),
(b) qc_step_host with page-locked (kernel-written) vs pageable (copied) result buffers."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import torch
from gpu_time import run
from common import initial_states
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim, _lib as L

def e2e(task, B, pinned, steps=200):
    params = configs.PRESETS[task]()
    sim = BatchedSim(params, batch=B, seed=1)
    psi0 = initial_states(params, min(B, 256), 1)
    sim.set_state(np.tile(psi0, ((B + psi0.shape[0] - 1) // psi0.shape[0], 1))[:B])
    act = torch.randint(0, params["n_levels"], (steps + 5, B), dtype=torch.int32).pin_memory()
    mk = (lambda *s, dt=torch.float64: torch.empty(s, dtype=dt).pin_memory()) if pinned else (lambda *s, dt=torch.float64: torch.empty(s, dtype=dt))
    mom, aux, flg = mk(B, sim.K), mk(B, L.QC_AUX_COUNT), mk(B, dt=torch.uint8)
    for i in range(5):
        sim.step_host(act[i], moments=mom, aux=aux, flags=flg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        sim.step_host(act[5 + i], moments=mom, aux=aux, flags=flg)
    dt = time.perf_counter() - t0
    print("e2e %-18s B=%5d pinned=%d  %8.1f us/step  %10.0f traj-steps/s  checksum %.12f" % (task, B, pinned, dt / steps * 1e6, B * steps / dt, float(mom.sum())), flush=True)

if __name__ == "__main__":
    for st in (0, 1500, 3000, 5000, 7000, 0):
        run("quartic", 1024, {"QCART_STAGGER": st}, steps=50)
    os.environ.pop("QCART_STAGGER", None)
    for _ in range(2):
        e2e("quartic", 1024, True); e2e("quartic", 1024, False)
    e2e("inverted_quartic", 8192, True, steps=10); e2e("inverted_quartic", 8192, False, steps=10)
