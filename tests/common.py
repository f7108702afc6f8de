"""Shared helpers of the test-suite: task parameter sets, seeded initial states, oracle drivers."""
import numpy as np
from math import pi, factorial

from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
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


def initial_states(params, B, seed=0):
    """Seeded synthetic initial states: grid = Gaussian packets (std 1, mean U(-1,1), wavenumber U(-0.3,0.3));
    Fock = coherent states |alpha|<=1, truncated and normalised."""
    rng = np.random.default_rng(seed)
    v = params["variant"]
    if v in ("quartic", "inverted_quartic"):
        h = params["grid_size"]
        half = int(params["x_max"] / h + 0.5)
        n = 2 * half + 1
        x = h * (np.arange(n) - half)
        mean = rng.uniform(-1, 1, B)
        k = rng.uniform(-0.3, 0.3, B)
        psi = np.exp(2j * pi * (x[None, :] - mean[:, None]) * k[:, None]) * np.exp(-(x[None, :] - mean[:, None]) ** 2 / 4) / (2 * pi) ** 0.25
        psi /= np.sqrt(np.sum(np.abs(psi) ** 2, axis=1, keepdims=True) * h)
    else:
        n = params["n_max"] + 1
        r = np.sqrt(rng.uniform(0, 1, B))
        ph = rng.uniform(0, 2 * pi, B)
        alpha = r * np.exp(1j * ph)
        kk = np.arange(n)
        logf = np.array([0.5 * np.sum(np.log(np.arange(1, m + 1))) for m in kk])
        psi = np.zeros((B, n), np.complex128)
        for b in range(B):
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                mag = np.where(kk == 0, 1.0, np.abs(alpha[b]) ** kk) / np.exp(logf)
            psi[b] = mag * np.exp(1j * np.angle(alpha[b]) * kk)
        psi[~np.isfinite(psi)] = 0
        psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    return np.ascontiguousarray(psi.astype(np.complex128))


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
