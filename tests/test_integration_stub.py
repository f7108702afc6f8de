"""INTEGRATION.md's reference-side binding is executable documentation: the ctypes stub a maintainer of the reference would drop next to
main_parallel.py (reference: quartic oscillator/main_parallel.py:120,226,481-509) is extracted from the markdown and run as written.

CPU part: the stub's struct must match the library's qc_config (size and field order against the packaged binding).
GPU part: with a stand-in `arguments` module, one trajectory is stepped through the stub and must reproduce the packaged `simulation`
mirror and the CPU oracle."""
import ctypes as C
import os
import re
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_source():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    src = [b for b in blocks if b.startswith("# simulation.py -- ctypes stub")]
    assert len(src) == 1
    return src[0]


def fake_arguments():
    """The fields the stub reads, with the reference's quartic defaults (quartic oscillator/arguments.py:6-20,46)."""
    ns = types.SimpleNamespace(x_max=8.5, grid_size=0.1, mass=1.0, time_steps=1440, gamma=0.01, n_con=18, F_max=5.0, input_moment_order=5, gpu_id=0)
    ns.__dict__["lambda"] = 0.04
    mod = types.ModuleType("arguments")
    mod.args = ns
    return mod


def test_stub_struct_matches_the_library():
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import _lib as L
    src = stub_source()
    m = re.search(r"class _Cfg\(C\.Structure\):.*?_fields_ = (\[.*?\])\n", src, flags=re.S)
    fields = eval(m.group(1), {"C": C})
    assert [f[0] for f in fields] == [f[0] for f in L.QcConfig._fields_]
    assert [f[1] for f in fields] == [f[1] for f in L.QcConfig._fields_]
    lib = L.load()
    assert lib.qc_config_size() == C.sizeof(L.QcConfig)


def test_out_of_date_struct_is_rejected_without_a_device():
    """struct_size is checked before anything touches CUDA: a binding built against an older header fails loudly with QC_ERR_ARG."""
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import _lib as L
    lib = L.load()
    cfg = L.QcConfig()
    cfg.struct_size = C.sizeof(L.QcConfig) - 8          # the round-1 stub: solve_tol missing
    h = C.c_void_p()
    assert lib.qc_create(C.byref(cfg), C.byref(h)) == L.QC_ERR_ARG
    assert b"struct_size" in lib.qc_last_error()


@pytest.mark.gpu
def test_stub_steps_one_trajectory_like_the_packaged_mirror(monkeypatch):
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import _lib as L, configs, simulation
    from common import initial_states, oracle_for
    monkeypatch.setenv("QCART_LIB", L.LIB_PATH)
    monkeypatch.setitem(sys.modules, "arguments", fake_arguments())
    stub = types.ModuleType("simulation_stub")
    exec(compile(stub_source(), "INTEGRATION.md:simulation.py", "exec"), stub.__dict__)

    params = configs.quartic()
    assert stub.check_settings() == (171, 0.1, params["lambda_"], params["mass"], 5)
    simulation.configure(params)
    assert simulation.check_settings() == stub.check_settings()

    psi0 = initial_states(params, 1, 4)[0]
    a, b = psi0.copy(), psi0.copy()
    stub.set_seed(77); simulation.set_seed(77)
    for k in range(5):
        F = [0.0, 2.5, -5.0, 2.5, 0.5][k]
        ra = stub.step(a, params["dt"], F, params["gamma"])
        rb = simulation.step(b, params["dt"], F, params["gamma"])
        assert ra == rb
    assert np.array_equal(a, b)
    # against the oracle: same Philox normals fed explicitly
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import philox_normals
    orc = oracle_for(params)
    c = psi0.copy()
    for k in range(5):
        F = [0.0, 2.5, -5.0, 2.5, 0.5][k]
        orc.run(c, params["dt"], F, params["gamma"], np.array([philox_normals(77, 0, k)]))
    assert np.linalg.norm(a - c) / np.linalg.norm(c) < 1e-12
    ma, mb = np.empty(20), np.empty(20)
    stub.get_moments(a, ma); simulation.get_moments(b, mb)
    assert np.array_equal(ma, mb)
    assert np.max(np.abs(ma - orc.get_moments(c)) / np.maximum(np.abs(orc.get_moments(c)), 1e-3)) < 1e-10
    assert stub.x_expectation(a) == simulation.x_expectation(b)
    with pytest.raises(ValueError):
        stub.step(a, 2 * params["dt"], 0.0, params["gamma"])          # dt differs from the handle's: QC_ERR_ARG -> ValueError
