"""Analytic controllers evaluated on the GPU from the moment vector (the first "next" row).

The reference computes these per actor on the CPU from the wavefunction (quartic oscillator/controllers.py:7-29, LQG inline in
harmonic oscillator/main_parallel.py:192-203 and inverted harmonic oscillator/main_parallel.py:194-206) and then quantises the
force to the 21 DQN levels (quartic oscillator/main_parallel.py:166-176).  Everything they need is already in the moment block the
control-step kernel emits, so here they are a handful of torch element-wise ops on the [B, K] tensor: a learner-free closed loop runs
entirely on the device.

Grid moment vector layout (simulation_quart.cpp:327-330): m[0]=<x>, m[1]=<p>, m[2]=<XX>, m[3]=Re<XP>, m[4]=<PP>, m[5]=<XXX>, m[6]=Re<XXP>, ...
with X = x-<x>, P = p-<p>.  Raw moments used by the reference controllers:
    <x^3>  = <XXX> + 3<x><XX> + <x>^3
    <x p x> = Re<x^2 p> = Re<XXP> + 2<x>Re<XP> + <p>(<XX> + <x>^2)        (x p x = x^2 p - i x)
"""
from math import pi, sqrt


def quantise_force(torch, force, f_max, n_levels=21):
    """clamp to [-f_max, f_max], round to the level grid, return (level index int32, force value) -- main_parallel.py:170-176."""
    half = (n_levels - 1) // 2
    step = f_max / half
    f = torch.clamp(force, -f_max, f_max)
    lvl = torch.round(f / step)                       # round-half-even, like Python's round()
    return (lvl + half).to(torch.int32), lvl * step


def lqg_harmonic(torch, moments, omega, n_con):
    """harmonic oscillator/main_parallel.py:192-203:  F = -((x+p) + (p-x) omega dt)/dt,  force = F/omega."""
    x, p = moments[:, 0], moments[:, 1]
    dt = 1.0 / n_con
    return (-((x + p) + (p - x) * omega * dt) / dt) / omega


def lqg_inverted_harmonic(torch, moments, omega, n_con):
    """inverted harmonic oscillator/main_parallel.py:194-206."""
    x, p = moments[:, 0], moments[:, 1]
    dt = 1.0 / n_con
    return (-(x + p) * (1 + omega * dt + 0.5 * omega * omega * dt * dt) / (dt + omega * dt * dt / 2)) / omega


def _raw(m):
    x, p = m[:, 0], m[:, 1]
    x3 = m[:, 5] + 3 * x * m[:, 2] + x ** 3
    xpx = m[:, 6] + 2 * x * m[:, 3] + p * (m[:, 2] + x * x)
    return x, p, m[:, 2], x3, xpx


def steepest_descent(torch, moments, lambda_, mass, n_con, damping=0.5):
    """quartic oscillator/controllers.py:7-15 ("damping"); returns F/pi (the `force` fed to simulation.step)."""
    x, p, var, x3, xpx = _raw(moments)
    ct = 1.0 / n_con
    p_pred = p - 4. * lambda_ * x3 * ct - ct ** 2 * 2. * lambda_ * 3. * xpx / mass
    return (-p_pred / ct * damping) / pi


def linear_quadratic(torch, moments, lambda_, mass, n_con, k):
    """quartic oscillator/controllers.py:17-21 ("LQG"), k = lambda * con_parameter."""
    x, p, var, x3, xpx = _raw(moments)
    ct = 1.0 / n_con
    x_pred = x + p / mass * ct - ct ** 2 * 2. * lambda_ * x3 / mass
    p_pred = p - 4. * lambda_ * x3 * ct - ct ** 2 * 2. * lambda_ * 3. * xpx / mass
    return (-(sqrt(k * mass) * x_pred + p_pred) / ct) / pi


def gaussian_approx(torch, moments, lambda_, mass, n_con):
    """quartic oscillator/controllers.py:23-29 ("semiclassical")."""
    x, p, var, x3, xpx = _raw(moments)
    ct = 1.0 / n_con
    target_p = -torch.sqrt(2 * mass * (6 * lambda_ * var + lambda_ * x * x)) * x
    return ((target_p - p) / ct) / pi
