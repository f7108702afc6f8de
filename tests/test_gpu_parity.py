"""GPU parity tests: the CUDA path through the C-ABI vs the CPU oracle on identical (psi0, operators, noise).

Tolerances (BASELINE.json north_star / SURVEY.md 8c): <= 1e-10 relative on psi and on every moment after one control
step; <= 1e-6 after a full episode.
"""
import numpy as np
import pytest

from common import TASKS, oracle_for, initial_states, oracle_control_step, fock_observation, level_force
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim, _lib as L

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-10


def _torch():
    import torch
    return torch


def rel_err(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)))


def run_case(task, B, seed=0, n_sub=None, overrides=None, want_q=False):
    torch = _torch()
    params = configs.PRESETS[task](**(overrides or {}))
    if n_sub:
        params["n_sub"] = n_sub
    rng = np.random.default_rng(seed + 100)
    psi0 = initial_states(params, B, seed)
    actions = rng.integers(0, params["n_levels"], B).astype(np.int32)
    noise = rng.standard_normal((B, params["n_sub"], 2))
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    out = sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"), want_q=want_q)
    torch.cuda.synchronize()
    psi_gpu = sim.get_state()
    orc = oracle_for(params)
    psi_ref, fails, qs = oracle_control_step(orc, params, psi0, actions, noise, want_q=want_q)
    return params, sim, out, psi_gpu, orc, psi_ref, fails, qs


@pytest.mark.parametrize("task", TASKS)
def test_control_step_matches_oracle(task):
    B = 24
    params, sim, out, psi_gpu, orc, psi_ref, fails, _ = run_case(task, B)
    err = rel_err(psi_gpu, psi_ref)
    print(task, sim.kernel_info(), "psi rel err", err)
    assert err < TOL_STEP
    flags = out["flags"].cpu().numpy()
    assert np.array_equal((flags & L.QC_FLAG_FAIL) != 0, fails != 0)
    mom = out["moments"].cpu().numpy()
    aux = out["aux"].cpu().numpy()
    assert np.allclose(aux[:, L.QC_AUX_NORM], 1.0, atol=1e-12)
    for b in range(B):
        if task in ("quartic", "inverted_quartic"):
            ref = orc.get_moments(psi_ref[b])
            scale = np.maximum(np.abs(ref), 1e-3)
            assert np.max(np.abs(mom[b] - ref) / scale) < TOL_STEP, (b, mom[b], ref)
            assert abs(aux[b, L.QC_AUX_XMEAN] - orc.x_expectation(psi_ref[b])) < 1e-10
        else:
            obs, nph = fock_observation(psi_ref[b], sim.n)
            assert np.max(np.abs(mom[b] - obs)) < 1e-9, (b, mom[b], obs)
            assert abs(aux[b, L.QC_AUX_ENERGY] - nph) < 1e-9


@pytest.mark.parametrize("task", TASKS)
def test_q_and_xmean_streams(task):
    B = 5
    params, sim, out, psi_gpu, orc, psi_ref, fails, qs = run_case(task, B, n_sub=20, want_q=True)
    q = out["q"].cpu().numpy()
    xm = out["x_mean"].cpu().numpy()
    for b in range(B):
        assert np.max(np.abs(xm[b] - qs[b][1])) < 1e-10
        assert np.max(np.abs(q[b] - qs[b][0]) / np.maximum(1.0, np.abs(qs[b][0]))) < 1e-10
