"""CPU tests of the oracle (the checker itself): two independent restatements agree, golden vectors from the reference's own
Python code, frozen regression vectors, and physics known-answer tests."""
import os
import numpy as np
import pytest
from math import pi, sqrt

from common import TASKS, oracle_for, initial_states, oracle_control_step, level_force, fock_observation
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
from oracle.sse_oracle import Oracle
from oracle.sse_oracle_np import OracleNP

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def np_oracle_for(params):
    v = params["variant"]
    if "quartic" in v:
        return OracleNP(v, x_max=params["x_max"], grid_size=params["grid_size"], lambda_=params["lambda_"], mass=params["mass"])
    return OracleNP(v, n_max=params["n_max"], omega=params["omega"], herm_mode=params.get("herm_mode", 0))


@pytest.mark.parametrize("task", TASKS)
def test_two_restatements_agree(task):
    """oracle/sse_oracle.c vs oracle/sse_oracle_np.py (scipy.sparse + real LAPACK zgbtrf/zgbtrs): <= 1e-13 per substep."""
    params = configs.PRESETS[task]()
    a, b = oracle_for(params), np_oracle_for(params)
    rng = np.random.default_rng(1)
    p1 = initial_states(params, 1, seed=2)[0]
    p2 = p1.copy()
    for lvl in (10, 20, 3, 0):
        F = level_force(params, lvl)
        for s in range(3):
            r = rng.standard_normal(2)
            o1 = a.step(p1, params["dt"], F, params["gamma"], r)
            o2 = b.step(p2, params["dt"], F, params["gamma"], r)
            assert np.linalg.norm(p1 - p2) / np.linalg.norm(p1) < 1e-13
            assert abs(o1[0] - o2[0]) <= 1e-13 * max(1.0, abs(o1[0])) and abs(o1[1] - o2[1]) < 1e-13 and o1[2] == o2[2]
            p2[:] = p1
    if "quartic" in task:
        assert np.max(np.abs(a.get_moments(p1) - b.get_moments(p1))) < 1e-12


@pytest.mark.parametrize("task", TASKS)
def test_no_pivoting_for_all_force_levels(task):
    """The device solve is pivot-free L D L^T; LAPACK's partial pivoting must be the identity for every force level."""
    params = configs.PRESETS[task]()
    o = oracle_for(params)
    for lvl in range(params["n_levels"]):
        assert np.array_equal(o.ipiv(params["dt"], level_force(params, lvl)), np.arange(o.n))


@pytest.mark.parametrize("task,fixture", [("quartic", "grid_reference_python_quartic.npz"), ("inverted_quartic", "grid_reference_python_inverted_quartic.npz")])
def test_grid_operators_match_reference_python(task, fixture):
    """Golden vectors produced by the reference's own space_def.py / main_parallel.py functions (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, fixture))
    params = configs.PRESETS[task]()
    o = oracle_for(params)
    assert o.n == g["x"].shape[0]
    assert np.max(np.abs(o.x_array() - g["x"])) < 1e-13
    H = o.H0_dense(0.0)
    for k in range(5):
        d = np.concatenate([np.diag(H, k), np.zeros(k)])
        assert np.max(np.abs(d - g["H_bands"][k])) <= 1e-12 * np.max(np.abs(g["H_bands"][k]))
    # p_hat: identical to the Python mirror except for the reference C code's truncated right edge (Q:59-70)
    P = o.p_dense()
    Pref = g["p_hat_dense_untruncated"]
    diff = np.abs(P - Pref)
    n = o.n
    assert np.max(diff[: n - 8, : n - 8]) < 1e-12
    assert np.count_nonzero(diff > 1e-12) == 2 * (1 + 2 + 3 + 4)
    assert np.allclose(P, P.conj().T)
    for b, psi in enumerate(g["states"]):
        psi = np.ascontiguousarray(psi)
        assert abs(o.x_expectation(psi) - g["x_mean"][b]) < 1e-12
        e = float(np.real(np.vdot(psi, H @ psi))) * params["grid_size"]
        assert abs(e - g["energy"][b]) < 1e-11 * max(1.0, abs(g["energy"][b]))
        m = o.get_moments(psi)
        assert abs(m[1] - g["p_mean_untruncated"][b]) < 1e-9          # the truncation only touches ~1e-30 amplitudes here
    if task == "inverted_quartic":
        h = params["grid_size"]
        r = int(np.rint(params["x_threshold"] / h))
        for b, psi in enumerate(g["states"]):
            out = 1.0 - np.sum(np.abs(psi[n // 2 - r: n // 2 + r]) ** 2) * h
            assert abs(out - g["outside"][b]) < 1e-13


def test_fock_observation_matches_reference_python():
    g = np.load(os.path.join(GOLD, "fock_reference_python.npz"))
    for task in ("harmonic", "inverted_harmonic"):
        params = configs.PRESETS[task]()
        n = params["n_max"] + 1
        o = oracle_for(params)
        for b, psi in enumerate(g[task + "_states"]):
            obs, nph = fock_observation(psi, n)
            assert np.max(np.abs(obs.astype(np.float32) - g[task + "_obs_float32"][b])) < 2e-6
            assert abs(o.x_expectation(np.ascontiguousarray(psi)) - g[task + "_x"][b]) < 1e-13
            if task == "harmonic":
                assert abs(nph - g["harmonic_phonon"][b]) < 1e-12


@pytest.mark.parametrize("task", TASKS)
def test_frozen_oracle_vectors(task):
    """Regression: committed outputs of the oracle on seeded PCG64 inputs (also what the GPU tests compare against at full size)."""
    g = np.load(os.path.join(GOLD, "oracle_control_step.npz"))
    params = configs.PRESETS[task]()
    orc = oracle_for(params)
    ref, fails, qs = oracle_control_step(orc, params, g[task + "_psi0"], g[task + "_actions"], g[task + "_noise"], want_q=True)
    assert np.max(np.linalg.norm(ref - g[task + "_psi1"], axis=1)) < 1e-12
    assert np.array_equal(fails, g[task + "_fail"])
    assert np.max(np.abs(np.array([q for q, _ in qs]) - g[task + "_q"]) / np.maximum(1, np.abs(g[task + "_q"]))) < 1e-12
    if "quartic" in task:
        m = np.array([orc.get_moments(ref[b]) for b in range(ref.shape[0])])
        assert np.max(np.abs(m - g[task + "_moments"]) / np.maximum(np.abs(g[task + "_moments"]), 1e-3)) < 1e-9


# ---- physics known-answer tests --------------------------------------------------------------------------------------

def test_harmonic_free_evolution_is_exact_phase_rotation():
    """gamma -> 0, F = 0: psi_n(t) = exp(-i omega (n+1/2) t) psi_n(0).  Checks Crank-Nicolson + the H^2..H^5 correction (O(dt^7))."""
    params = configs.harmonic()
    o = oracle_for(params)
    psi0 = initial_states(params, 1, seed=3)[0]
    psi = psi0.copy()
    steps, dt = 200, params["dt"]
    for s in range(steps):
        o.step(psi, dt, 0.0, 1e-14, np.zeros(2))
    nn = np.arange(o.n)
    exact = np.exp(-1j * params["omega"] * (nn + 0.5) * dt * steps) * psi0
    mask = np.abs(psi0) > 1e-12
    assert np.max(np.abs(psi[mask] - exact[mask])) < 1e-11


@pytest.mark.parametrize("task", TASKS)
def test_norm_and_determinism(task):
    params = configs.PRESETS[task]()
    o = oracle_for(params)
    w = params.get("grid_size", 1.0) if "quartic" in task else 1.0
    rng = np.random.default_rng(3)
    noise = rng.standard_normal((10, 2))
    a = initial_states(params, 1, seed=5)[0]
    b = a.copy()
    for s in range(10):
        o.step(a, params["dt"], level_force(params, 14), params["gamma"], noise[s])
        assert abs(np.sum(np.abs(a) ** 2) * w - 1.0) < 1e-13
    o2 = oracle_for(params)
    for s in range(10):
        o2.step(b, params["dt"], level_force(params, 14), params["gamma"], noise[s])
    assert np.array_equal(a, b)


def test_strong_order_under_step_halving():
    """Shared Brownian path, dt -> dt/2 -> dt/4: the scheme is strong order 1.5 (Platen), errors must shrink by ~2^1.5 per halving."""
    params = configs.quartic(gamma=0.5 * pi)
    T_steps, dt0 = 16, 1.0 / 1440
    rng = np.random.default_rng(9)
    # finest increments (dt0/4): dW, dZ
    nf = T_steps * 4
    dtf = dt0 / 4
    r = rng.standard_normal((nf, 2))
    dW = r[:, 0] * sqrt(dtf)
    dZ = 0.5 * dtf ** 1.5 * (r[:, 0] + r[:, 1] / sqrt(3.))

    def coarsen(dW, dZ, dt):
        # dW_c = dW1 + dW2, dZ_c = dZ1 + dZ2 + dt * dW1  (iterated integral of the Wiener increment)
        return dW[0::2] + dW[1::2], dZ[0::2] + dZ[1::2] + dt * dW[0::2]

    def to_normals(dW, dZ, dt):
        r0 = dW / sqrt(dt)
        r1 = (dZ / (0.5 * dt ** 1.5) - r0) * sqrt(3.)
        return np.stack([r0, r1], axis=1)

    psi0 = initial_states(params, 1, seed=8)[0]
    sols = []
    levels = [(dtf, dW, dZ)]
    dWc, dZc = coarsen(dW, dZ, dtf)
    levels.append((2 * dtf, dWc, dZc))
    dWc2, dZc2 = coarsen(dWc, dZc, 2 * dtf)
    levels.append((4 * dtf, dWc2, dZc2))
    for dt, w_, z_ in levels:
        o = Oracle("quartic", x_max=params["x_max"], grid_size=params["grid_size"], lambda_=params["lambda_"], mass=params["mass"])
        psi = psi0.copy()
        for rr in to_normals(w_, z_, dt):
            o.step(psi, dt, 1.0, params["gamma"], rr)
        sols.append(psi)
    e_fine = np.linalg.norm(sols[1] - sols[0])
    e_coarse = np.linalg.norm(sols[2] - sols[1])
    ratio = e_coarse / e_fine
    assert 1.8 < ratio < 6.0, ratio


def test_gaussian_covariance_is_noise_independent():
    """Harmonic SSE from the vacuum keeps the state Gaussian: (Var x, Var p, Cov) follow a deterministic Riccati flow, independent of
    the measurement record."""
    params = configs.harmonic()
    res = []
    for seed in (1, 2):
        o = oracle_for(params)
        psi = np.zeros(o.n, np.complex128); psi[0] = 1.0
        rng = np.random.default_rng(seed)
        for s in range(160):
            o.step(psi, params["dt"], 0.0, params["gamma"], rng.standard_normal(2))
        obs, _ = fock_observation(psi, o.n)
        res.append(obs)
    assert np.max(np.abs(res[0][2:] - res[1][2:])) < 1e-4            # equal up to the (path dependent) discretisation error
    assert abs(res[0][0] - res[1][0]) > 1e-4        # while the means do follow the noise


def test_inverted_harmonic_hermitian_descriptor_quirk():
    """I:23,551 applies the complex-symmetric correction matrix with a HERMITIAN descriptor; the effect is tiny but not zero."""
    params = configs.inverted_harmonic()
    psi0 = initial_states(params, 1, seed=6)[0]
    outs = []
    for mode in (0, 1, 2):
        o = oracle_for(params, herm_mode=mode)
        psi = psi0.copy()
        for s in range(5):
            o.step(psi, params["dt"], 8.0, params["gamma"], np.array([0.3, -0.2]))
        outs.append(psi)
    d02 = np.linalg.norm(outs[0] - outs[2])
    assert 0 < d02 < 1e-6
    assert np.linalg.norm(outs[0] - outs[1]) < 1e-7


def test_unused_reference_exports():
    """Hamiltonian_dot_psi / solve_ab (H:566-597): A (A^-1 b) = b and H psi for a Fock basis state."""
    params = configs.harmonic()
    o = oracle_for(params)
    rng = np.random.default_rng(0)
    b = (rng.standard_normal(o.n) + 1j * rng.standard_normal(o.n)).astype(np.complex128)
    x = o.solve_ab(params["dt"], 2.0, b.copy())
    assert np.linalg.norm(o.A_dense(params["dt"], 2.0) @ x - b) / np.linalg.norm(b) < 1e-13
    e3 = np.zeros(o.n, np.complex128); e3[3] = 1.0
    assert abs(o.hamiltonian_dot_psi(e3)[3] - params["omega"] * 3.5) < 1e-13
