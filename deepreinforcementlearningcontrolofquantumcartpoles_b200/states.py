"""Seeded synthetic initial states of the benchmark and the tests: what the reference's actors start from, with a spread.

grid tasks: Gaussian packets like Gaussian_packet() of quartic main_parallel.py:75-76 (std 1, mean U(-1,1), wavenumber U(-0.3,0.3));
Fock tasks: coherent states |alpha| <= 1, truncated at n_max and normalised (the reference starts from the vacuum, harmonic main_parallel.py:226-227).
"""
import numpy as np
from math import pi


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
