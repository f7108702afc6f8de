"""Pins of the oracle against the REFERENCE'S OWN C++ (simulation*.cpp compiled unmodified against oracle/mkl_shim/mkl.h):
(1) the committed fixture tests/golden/reference_build_control_step.npz, produced by that build (always runs);
(2) the live build under oracle/_ref/ when it exists (authoring container and GPU box; built by oracle/build_ref.sh)."""
import os
import numpy as np
import pytest

from common import TASKS, oracle_for, oracle_control_step, initial_states, level_force
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_build_control_step.npz")


@pytest.mark.parametrize("task", TASKS)
def test_oracle_matches_reference_build_fixture(task):
    g = np.load(GOLD)
    params = configs.PRESETS[task]()
    orc = oracle_for(params)
    ref, fails, qs = oracle_control_step(orc, params, g[task + "_psi0"], g[task + "_actions"], g[task + "_noise"], want_q=True)
    assert np.max(np.linalg.norm(ref - g[task + "_psi1"], axis=1) / np.linalg.norm(g[task + "_psi1"], axis=1)) < 1e-12
    assert np.array_equal(fails, g[task + "_fail"])
    q = np.array([a for a, _ in qs]); xm = np.array([b for _, b in qs])
    assert np.max(np.abs(q - g[task + "_q"]) / np.maximum(1, np.abs(g[task + "_q"]))) < 1e-12
    assert np.max(np.abs(xm - g[task + "_xmean"])) < 1e-12
    for b in range(ref.shape[0]):
        assert abs(orc.x_expectation(np.ascontiguousarray(g[task + "_psi1"][b])) - g[task + "_xexp"][b]) < 1e-12
    if "quartic" in task:
        m = np.array([orc.get_moments(np.ascontiguousarray(g[task + "_psi1"][b])) for b in range(ref.shape[0])])
        assert np.max(np.abs(m - g[task + "_moments"]) / np.maximum(np.abs(g[task + "_moments"]), 1e-3)) < 1e-9
        assert int(g[task + "_settings"][0]) == orc.n


@pytest.mark.parametrize("task", TASKS)
def test_oracle_matches_live_reference_build(task):
    from oracle.ref_module import RefModule, available
    if not available(task):
        pytest.skip("oracle/_ref not built here (needs /root/reference at build time)")
    params = configs.PRESETS[task]()
    ref, orc = RefModule(task), oracle_for(params)
    rng = np.random.default_rng(3)
    a = initial_states(params, 1, 17)[0]
    b = a.copy()
    for lvl in (10, 0, 0, 20, 7):            # force changes exercise the reference's reset_ab cache (Q:513-518)
        for s in range(3):
            r = rng.standard_normal(2)
            o1 = ref.step(a, params["dt"], level_force(params, lvl), params["gamma"], r)
            o2 = orc.step(b, params["dt"], level_force(params, lvl), params["gamma"], r)
            assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-13
            assert abs(o1[0] - o2[0]) < 1e-12 * max(1, abs(o2[0])) and abs(o1[1] - o2[1]) < 1e-13 and o1[2] == o2[2]
    # the reference's argument checks (Q:288-305) are live in this build
    with pytest.raises(ValueError):
        ref.mod.step(np.zeros(5, np.complex128), params["dt"], 0.0, params["gamma"])
