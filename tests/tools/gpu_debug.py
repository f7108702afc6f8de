"""Development aid: run every task through the CUDA path for several launch geometries and print errors vs the oracle."""
import os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import TASKS, oracle_for, initial_states, oracle_control_step, fock_observation
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim, _lib as L

def one(task, B, n_sub, env):
    for k in ("QCART_L", "QCART_T", "QCART_P", "QCART_BIN", "QCART_JACOBI"):
        os.environ.pop(k, None)
    os.environ.update(env)
    params = configs.PRESETS[task](); params["n_sub"] = n_sub
    rng = np.random.default_rng(5)
    psi0 = initial_states(params, B, 1)
    actions = rng.integers(0, params["n_levels"], B).astype(np.int32)
    noise = rng.standard_normal((B, n_sub, 2))
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    out = sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
    torch.cuda.synchronize()
    g = sim.get_state()
    orc = oracle_for(params)
    ref, fails, _ = oracle_control_step(orc, params, psi0, actions, noise)
    err = np.linalg.norm(g - ref, axis=1) / np.linalg.norm(ref, axis=1)
    mom = out["moments"].cpu().numpy(); aux = out["aux"].cpu().numpy()
    if "quartic" in task:
        mref = np.stack([orc.get_moments(ref[b]) for b in range(B)])
        merr = np.max(np.abs(mom - mref) / np.maximum(np.abs(mref), 1e-3))
    else:
        mref = np.stack([fock_observation(ref[b], sim.n)[0] for b in range(B)])
        merr = np.max(np.abs(mom - mref))
    print("%-18s %-28s n_sub=%3d  psi err max %.2e (argmax %d)  moment err %.2e  norm dev %.1e  flags %s/%s | %s" % (
        task, env, n_sub, err.max(), int(err.argmax()), merr, np.max(abs(aux[:, 3] - 1)), out["flags"].cpu().numpy().tolist()[:6], fails.tolist()[:6], sim.kernel_info()), flush=True)

if __name__ == "__main__":
    cases = []
    for task in TASKS:
        Ls = ["3", "5", "6", "2", "9"] if "quartic" in task else ["1", "2", "3"]
        for Lv in Ls:
            cases.append((task, 9, 1, {"QCART_L": Lv}))
            cases.append((task, 9, 8, {"QCART_L": Lv}))
        cases.append((task, 9, 8, {"QCART_P": "1"}))
        cases.append((task, 9, 8, {"QCART_T": "1"}))
        cases.append((task, 40, 8, {"QCART_BIN": "1", "QCART_T": "3"}))
        cases.append((task, 40, 8, {"QCART_BIN": "1", "QCART_T": "2", "QCART_JACOBI": "0"}))
    for c in cases:
        try:
            one(*c)
        except Exception as e:
            print("FAILED", c, repr(e)); traceback.print_exc()
