"""Analytic controllers from the moment vector against the reference's own controllers.py outputs
(tests/golden/controllers_reference_python.npz, produced by running the reference functions unmodified)."""
import os
import numpy as np
import pytest
from math import pi

from common import oracle_for, initial_states, level_force
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, controllers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "controllers_reference_python.npz")


def _check(moments, params, g, torch):
    lam, m = params["lambda_"], params["mass"]
    got = {"damping": controllers.steepest_descent(torch, moments, lam, m, 18, damping=0.5) * pi,
           "lqg": controllers.linear_quadratic(torch, moments, lam, m, 18, k=lam * 2.0) * pi,
           "semiclassical": controllers.gaussian_approx(torch, moments, lam, m, 18) * pi}
    for k, v in got.items():
        ref = g[k]
        err = np.max(np.abs(v.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1.0))
        assert err < 1e-6, (k, err)          # the reference uses the un-truncated FD p_hat and x p x; differences are O(h^8) + boundary
    lvl, f = controllers.quantise_force(torch, got["lqg"] / pi, params["f_max"], 21)
    assert int(lvl.min()) >= 0 and int(lvl.max()) <= 20
    assert np.allclose(f.cpu().numpy(), (lvl.cpu().numpy() - 10) * 0.5)


def test_controllers_from_oracle_moments_match_reference_python():
    import torch
    g = np.load(GOLD)
    params = configs.quartic()
    orc = oracle_for(params)
    mom = np.stack([orc.get_moments(np.ascontiguousarray(s)) for s in g["states"]])
    _check(torch.as_tensor(mom), params, g, torch)


def test_quantisation_matches_reference_rounding():
    import torch
    f = torch.tensor([-7.3, -5.0, -0.25, 0.25, 0.75, 1.25, 4.76, 9.0], dtype=torch.float64)
    lvl, q = controllers.quantise_force(torch, f, 5.0, 21)
    ref = [round(min(max(v, -5.0), 5.0) / 0.5) for v in f.tolist()]          # Python round(): half to even, as in main_parallel.py:174
    assert lvl.tolist() == [r + 10 for r in ref]


@pytest.mark.gpu
def test_controllers_on_gpu_moments():
    import torch
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import BatchedSim
    g = np.load(GOLD)
    params = configs.quartic()
    sim = BatchedSim(params, batch=g["states"].shape[0])
    sim.set_state(g["states"])
    out = sim.get_moments()
    _check(out["moments"], params, g, torch)


@pytest.mark.gpu
def test_config1_harmonic_lqg_closed_loop_matches_oracle():
    """BASELINE configs[0]: harmonic oscillator cooling, LQG controller.  Closed loop on the GPU (moments -> LQG force -> quantised level
    -> control step) against the same loop around the CPU oracle with identical noise: identical actions, psi within 1e-9."""
    import torch
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import BatchedSim
    from common import fock_observation
    params = configs.harmonic()
    B, n_ctrl = 3, 60
    rng = np.random.default_rng(42)
    sim = BatchedSim(params, batch=B)
    sim.init_fock(None)
    orc = oracle_for(params)
    ref = np.zeros((B, sim.n), np.complex128); ref[:, 0] = 1.0
    act_gpu = torch.full((B,), 10, dtype=torch.int32, device="cuda")        # first interval: F = 0 (control skipped at i = 0)
    act_ref = np.full(B, 10)
    for c in range(n_ctrl):
        noise = rng.standard_normal((B, params["n_sub"], 2))
        out = sim.step(act_gpu, noise=torch.as_tensor(noise, device="cuda"))
        for b in range(B):
            st = ref[b].copy()
            orc.run(st, params["dt"], level_force(params, int(act_ref[b])), params["gamma"], noise[b])
            ref[b] = st
        # controller on both sides
        f_gpu = controllers.lqg_harmonic(torch, out["moments"], params["omega"], 18)
        act_gpu, _ = controllers.quantise_force(torch, f_gpu, params["f_max"], 21)
        obs = np.stack([fock_observation(ref[b], sim.n)[0] for b in range(B)])
        f_ref = controllers.lqg_harmonic(torch, torch.as_tensor(obs), params["omega"], 18)
        a_ref, _ = controllers.quantise_force(torch, f_ref, params["f_max"], 21)
        act_ref = a_ref.numpy()
        assert np.array_equal(act_gpu.cpu().numpy(), act_ref), c
    got = sim.get_state()
    assert np.max(np.linalg.norm(got - ref, axis=1)) < 1e-9
    # LQG cools: phonon number stays small
    assert float(out["aux"][:, 0].max()) < 3.0
