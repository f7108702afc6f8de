"""Development aid: time the control-step kernel for a task over launch geometries (CUDA events)."""
import os, sys, itertools
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import initial_states, oracle_for, oracle_control_step
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim

FLOPS = {"harmonic": 1.80e6, "inverted_harmonic": 5.28e6, "quartic": 7.81e6, "inverted_quartic": 47.5e6}

def run(task, B, env, steps=20, check=False):
    for k in ("QCART_L", "QCART_T", "QCART_P", "QCART_TABS", "QCART_GC", "QCART_JACOBI", "QCART_BIN", "QCART_MAXT"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    params = configs.PRESETS[task]()
    sim = BatchedSim(params, batch=B, seed=1)
    psi0 = initial_states(params, min(B, 256), 1)
    reps = (B + psi0.shape[0] - 1) // psi0.shape[0]
    sim.set_state(np.tile(psi0, (reps, 1))[:B])
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    act = torch.randint(0, params["n_levels"], (B,), device="cuda", dtype=torch.int32, generator=g)
    out = sim.alloc_outputs()
    try:
        for _ in range(3):
            sim.step(act, out=out)
        torch.cuda.synchronize()
    except Exception as e:
        print("%-18s B=%6d %-30s FAILED %r" % (task, B, env, e)); return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sim.step(act, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    rate = B / (ms * 1e-3)
    aux = out["aux"].cpu().numpy()
    print("%-18s B=%6d %-34s %8.3f ms/step  %10.0f traj-steps/s  %6.2f TFLOP/s(alg)  normdev %.1e | %s" % (
        task, B, env, ms, rate, rate * FLOPS[task] / 1e12, np.max(abs(aux[:, 3] - 1)), sim.kernel_info()), flush=True)

if __name__ == "__main__":
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import measure_peaks
    print("measured peaks: fp64 %.2f TFLOP/s, smem %.2f TB/s" % tuple(v / 1e12 for v in measure_peaks(0)))
    which = sys.argv[1] if len(sys.argv) > 1 else "quartic"
    if which == "quartic":
        for L, T, P in [(6, 7, 32), (6, 4, 32), (6, 8, 32)]:
            run("quartic", 1024, {"QCART_L": L, "QCART_T": T, "QCART_P": P})
        for L, T, P in [(6, 7, 32), (6, 8, 32), (6, 4, 32)]:
            run("quartic", 8192, {"QCART_L": L, "QCART_T": T, "QCART_P": P})
    elif which == "iq":
        for L, T, P in [(6, 3, 32), (6, 2, 32), (6, 1, 32), (9, 2, 32), (9, 3, 32), (3, 2, 32), (5, 2, 32)]:
            run("inverted_quartic", 8192, {"QCART_L": L, "QCART_T": T, "QCART_P": P}, steps=5)
        run("inverted_quartic", 8192, {"QCART_L": 6, "QCART_T": 3, "QCART_P": 32, "QCART_TABS": 0}, steps=5)
        run("inverted_quartic", 8192, {"QCART_L": 6, "QCART_T": 4, "QCART_P": 32, "QCART_TABS": 0}, steps=5)
        run("inverted_quartic", 1024, {}, steps=5)
    elif which == "fock":
        for task, Ls in (("harmonic", [(3, 8), (3, 4), (2, 8), (1, 4)]), ("inverted_harmonic", [(6, 8), (6, 4), (3, 8), (3, 4), (2, 4)])):
            for L, T in Ls:
                run(task, 8192, {"QCART_L": L, "QCART_T": T}, steps=10)
            run(task, 1024, {}, steps=10)
