"""Presets of the reference's four tasks (defaults of each task's arguments.py) and the grid-size sweep.

Citations are into /root/reference/implementation codes/<task>/arguments.py and main_parallel.py.
pi-scalings: lambda = args.lambda*pi, gamma = args.gamma*pi, mass = args.mass/pi (quartic main_parallel.py:29-31);
omega = pi, gamma = args.gamma*pi (harmonic main_parallel.py:42-44).
"""
from math import pi


def harmonic(**kw):
    """harmonic oscillator cooling: n_max 70, gamma 1, 1440 substeps / 18 controls per unit time, F_max 5."""
    d = dict(variant="harmonic", n_max=70, omega=pi, gamma=1.0 * pi, dt=1.0 / 1440, n_sub=80, f_max=5.0, n_levels=21,
             input_scaling=1.0, reward_scale=10.0, phonon_cutoff=20.0, t_max=100.0)
    d.update(kw)
    return d


def inverted_harmonic(**kw):
    """inverted harmonic cartpole: n_max 180, gamma 2, F_max 8."""
    d = dict(variant="inverted_harmonic", n_max=180, omega=pi, gamma=2.0 * pi, dt=1.0 / 1440, n_sub=80, f_max=8.0, n_levels=21,
             input_scaling=1.0, herm_mode=0)
    d.update(kw)
    return d


def quartic(**kw):
    """quartic oscillator cooling: x_max 8.5, h 0.1 (N=171), lambda 0.04 pi, m 1/pi, gamma 0.01 pi, moments up to order 5."""
    d = dict(variant="quartic", x_max=8.5, grid_size=0.1, lambda_=0.04 * pi, mass=1.0 / pi, gamma=0.01 * pi, dt=1.0 / 1440,
             n_sub=80, f_max=5.0, n_levels=21, moment_order=5, input_scaling=1.0, reward_scale=1.0, energy_cutoff=12.0,
             init_energy_cutoff=7.5, t_max=100.0)
    d.update(kw)
    return d


def inverted_quartic(**kw):
    """inverted quartic cartpole: x_max 13, h 0.05 (N=521), lambda -0.01 pi, gamma pi, 2880 substeps / 18 controls."""
    lam = -0.01 * pi
    f_max = 5.0
    d = dict(variant="inverted_quartic", x_max=13.0, grid_size=0.05, lambda_=lam, mass=1.0 / pi, gamma=1.0 * pi, dt=1.0 / 2880,
             n_sub=160, f_max=f_max, n_levels=21, moment_order=5, input_scaling=1.0,
             x_threshold=(f_max / abs(lam) / 4 * pi) ** (1 / 3))      # xth, inverted quartic main_parallel.py:150
    d.update(kw)
    return d


def quartic_sweep(n_points, **kw):
    """grid-size sweep of the quartic cartpole (BASELINE.json config 5): x_max 13 fixed, N = n_points (odd),
    dt scaled with h^2 from (h=0.05, dt=1/2880) as implementation codes/readme.md:9 advises."""
    assert n_points % 2 == 1
    h = 2 * 13.0 / (n_points - 1)
    d = inverted_quartic(grid_size=h, dt=(1.0 / 2880) * (h / 0.05) ** 2)
    d.update(kw)
    return d


PRESETS = {"harmonic": harmonic, "inverted_harmonic": inverted_harmonic, "quartic": quartic, "inverted_quartic": inverted_quartic}
