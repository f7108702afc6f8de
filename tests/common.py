"""Shared helpers of the test-suite: task parameter sets, seeded initial states, oracle drivers."""
import numpy as np
from math import pi, factorial

from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states  # noqa: F401  (re-exported for the tests)
from oracle.sse_oracle import Oracle

TASKS = ["harmonic", "inverted_harmonic", "quartic", "inverted_quartic"]


def oracle_for(params, fast=False, herm_mode=None):
    v = params["variant"]
    if v in ("quartic", "inverted_quartic"):
        return Oracle(v, x_max=params["x_max"], grid_size=params["grid_size"], lambda_=params["lambda_"], mass=params["mass"],
                      moment_order=params.get("moment_order", 5), fast=fast)
    hm = params.get("herm_mode", 0) if herm_mode is None else herm_mode
    return Oracle(v, n_max=params["n_max"], omega=params["omega"], herm_mode=hm, fast=fast)


def level_force(params, a):
    half = (params.get("n_levels", 21) - 1) // 2
    return (a - half) * params["f_max"] / half


def oracle_control_step(orc, params, psi, actions, noise, want_q=False):
    """Reference semantics of one control step for every trajectory: n_sub calls of step() at the chosen force, Fail latched."""
    B = psi.shape[0]
    out = psi.copy()
    fails = np.zeros(B, np.int32)
    qs = []
    for b in range(B):
        st = out[b].copy()
        f, q, xm = orc.run(st, params["dt"], level_force(params, int(actions[b])), params["gamma"], noise[b], want_q=want_q)
        out[b] = st
        fails[b] = f
        qs.append((q, xm))
    return out, fails, qs


def fock_observation(psi, n):
    """get_data_xp + phonon_number of harmonic main_parallel.py:88-89,128-130 restated with dense operators."""
    a = np.diag(np.sqrt(np.arange(1, n)), 1)
    xh = np.sqrt(0.5) * (a.T + a)
    ph = 1j * np.sqrt(0.5) * (a.T - a)
    ex = lambda op: float(np.real(np.vdot(psi, op @ psi)))
    x, p = ex(xh), ex(ph)
    obs = np.array([x, p, ex(xh @ xh) - x * x, ex(np.real(ph @ ph)) - p * p, ex(xh @ ph + ph @ xh) / 2 - x * p])
    nph = float(np.sum(np.abs(psi) ** 2 * np.arange(n)))
    return obs, nph
