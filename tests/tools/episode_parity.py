"""Full-length episodes, GPU (through the C-ABI) vs the CPU oracle with identical initial state, forces and noise: error growth curve.

    python tests/tools/episode_parity.py [n_ctrl_cooling=1800] [n_ctrl_inverted=600] > gpurun_out/episode_parity.jsonl

Cooling tasks (harmonic, quartic): t_max = 100 time units = 1800 control steps, forces from a damping feedback on the oracle's own
moments (quantised to the 21 levels).  Inverted tasks: a stabilising linear feedback keeps the cartpole alive; the run ends at the first
failure of either side or after n_ctrl_inverted control steps.  One JSON line per task with the relative state error every 50 steps."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import oracle_for, fock_observation, level_force
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim, _lib as L
from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states


GAINS = {"harmonic": (0.0, 1.0), "inverted_harmonic": (2.0, 2.0), "quartic": (0.0, 0.3), "inverted_quartic": (3.0, 2.0)}


def quantise(task, params, x, p):
    """Quantised linear feedback F = -(c1 <x> + c2 <p>) on the 21 force levels."""
    c = GAINS[task]
    F = -(c[0] * x + c[1] * p)
    half = (params["n_levels"] - 1) // 2
    a = int(round(F / (params["f_max"] / half))) + half
    return min(max(a, 0), params["n_levels"] - 1)


def feedback(task, params, orc, psi, n):
    """The feedback evaluated on the ORACLE's state."""
    if task in ("harmonic", "inverted_harmonic"):
        obs, _ = fock_observation(psi, n)
        return quantise(task, params, obs[0], obs[1])
    m = orc.get_moments(psi)
    return quantise(task, params, m[0], m[1])


def run(task, n_ctrl, B=2, seed=12, every=50, log=None, own_feedback=False):
    """own_feedback=False: both sides receive the force computed from the oracle's state (open loop for the GPU side).
    own_feedback=True: each side closes its own loop on its own moments, as two independent installations of the environment would."""
    params = configs.PRESETS[task]()
    rng = np.random.default_rng(seed)
    if task == "inverted_quartic":
        psi0 = initial_states(params, B, 5)
        x = params["grid_size"] * (np.arange(psi0.shape[1]) - psi0.shape[1] // 2)
        psi0 = np.tile(np.exp(-x ** 2 / 4) / (2 * np.pi) ** 0.25, (B, 1)).astype(np.complex128)     # Gaussian_packet(inf, 0, 1), inverted quartic main_parallel.py:182-183
    elif task == "inverted_harmonic":
        psi0 = np.zeros((B, params["n_max"] + 1), np.complex128); psi0[:, 0] = 1.0
    else:
        psi0 = initial_states(params, B, 5)
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    orc = oracle_for(params)
    ref = psi0.copy()
    curve, steps_done, failed, action_mismatch = [], 0, False, 0
    mom = sim.get_moments()["moments"].cpu().numpy()
    for c in range(n_ctrl):
        actions = np.array([feedback(task, params, orc, ref[b], sim.n) for b in range(B)], np.int32)
        gpu_actions = np.array([quantise(task, params, mom[b, 0], mom[b, 1]) for b in range(B)], np.int32) if own_feedback else actions
        action_mismatch += int(np.sum(gpu_actions != actions))
        noise = rng.standard_normal((B, params["n_sub"], 2))
        out = sim.step(torch.as_tensor(gpu_actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
        mom = out["moments"].cpu().numpy()
        fl = []
        for b in range(B):
            f, _, _ = orc.run(ref[b], params["dt"], level_force(params, int(actions[b])), params["gamma"], noise[b])
            fl.append(f)
        steps_done = c + 1
        flags = out["flags"].cpu().numpy()
        aux = out["aux"].cpu().numpy()
        if steps_done % every == 0 or steps_done == n_ctrl:
            got = sim.get_state()
            err = float(np.max(np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)))
            curve.append((steps_done, err))
            if log: print(task, steps_done, err, file=log, flush=True)
        if task == "inverted_harmonic" and np.any(np.abs(aux[:, L.QC_AUX_XMEAN]) > params["f_max"]): failed = True
        if task == "inverted_quartic" and np.any(flags & L.QC_FLAG_ESCAPED): failed = True
        if np.any(flags & L.QC_FLAG_FAIL) or any(fl): failed = True
        if failed:
            got = sim.get_state()
            curve.append((steps_done, float(np.max(np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)))))
            break
    return {"task": task, "mode": "each side closes its own loop" if own_feedback else "shared force sequence (from the oracle's state)", "action_mismatches": action_mismatch,
            "control_steps": steps_done, "substeps": steps_done * params["n_sub"], "ended_by_failure": failed, "trajectories": B,
            "final_rel_err": curve[-1][1], "max_rel_err": max(e for _, e in curve), "curve": curve, "kernel": sim.kernel_info()}


def run_cpu_pair(task, n_ctrl, seed=12, every=25):
    """The same experiment between two CPU builds of the oracle (strict IEEE vs -Ofast, the reference's own flag): how fast two correct
    implementations of the reference algorithm separate under a shared force sequence.  This is the achievable bound for any port."""
    params = configs.PRESETS[task]()
    rng = np.random.default_rng(seed)
    a_orc, b_orc = oracle_for(params), oracle_for(params, fast=True)
    if task == "inverted_quartic":
        n = 2 * int(params["x_max"] / params["grid_size"] + 0.5) + 1
        x = params["grid_size"] * (np.arange(n) - n // 2)
        psi = (np.exp(-x ** 2 / 4) / (2 * np.pi) ** 0.25).astype(np.complex128)
    else:
        n = params["n_max"] + 1
        psi = np.zeros(n, np.complex128); psi[0] = 1.0
    a, b = psi.copy(), psi.copy()
    curve = []
    for c in range(n_ctrl):
        act = feedback(task, params, a_orc, a, n)
        noise = rng.standard_normal((params["n_sub"], 2))
        fa, _, _ = a_orc.run(a, params["dt"], level_force(params, act), params["gamma"], noise)
        fb, _, _ = b_orc.run(b, params["dt"], level_force(params, act), params["gamma"], noise)
        if (c + 1) % every == 0:
            curve.append((c + 1, float(np.linalg.norm(a - b) / np.linalg.norm(a))))
        if fa or fb:
            break
    return {"task": task, "mode": "CPU oracle strict vs CPU oracle -Ofast, shared force sequence", "control_steps": c + 1, "curve": curve}


if __name__ == "__main__":
    n_cool = int(sys.argv[1]) if len(sys.argv) > 1 else 1800
    n_inv = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    for task, n in (("harmonic", n_cool), ("quartic", n_cool), ("inverted_harmonic", n_inv), ("inverted_quartic", n_inv)):
        print(json.dumps(run(task, n, log=sys.stderr)), flush=True)
    for task in ("inverted_harmonic", "inverted_quartic"):
        print(json.dumps(run(task, n_inv, log=sys.stderr, own_feedback=True, every=25)), flush=True)
        print(json.dumps(run_cpu_pair(task, min(n_inv, 200))), flush=True)
