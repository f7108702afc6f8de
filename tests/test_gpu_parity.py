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


@pytest.mark.parametrize("task", ["quartic", "inverted_harmonic"])
def test_many_trajectories_per_cta(task):
    """Batch large enough that the planner packs several trajectories per CTA (T > 1) and several CTAs per SM."""
    B = 1500 if task == "quartic" else 1300
    params, sim, out, psi_gpu, orc, psi_ref, fails, _ = run_case(task, B, n_sub=12)
    assert " T=1 " not in sim.kernel_info()
    assert rel_err(psi_gpu, psi_ref) < TOL_STEP


@pytest.mark.parametrize("task", TASKS)
def test_frozen_golden_vectors_through_the_abi(task):
    """tests/golden/oracle_control_step.npz (committed): same psi0 / actions / noise through the CUDA path."""
    import os
    torch = _torch()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_control_step.npz"))
    params = configs.PRESETS[task]()
    sim = BatchedSim(params, batch=3)
    sim.set_state(g[task + "_psi0"])
    out = sim.step(torch.as_tensor(g[task + "_actions"], device="cuda"), noise=torch.as_tensor(g[task + "_noise"], device="cuda"), want_q=True)
    torch.cuda.synchronize()
    assert rel_err(sim.get_state(), g[task + "_psi1"]) < TOL_STEP
    assert np.array_equal((out["flags"].cpu().numpy() & L.QC_FLAG_FAIL) != 0, g[task + "_fail"] != 0)
    q = out["q"].cpu().numpy()
    assert np.max(np.abs(q - g[task + "_q"]) / np.maximum(1, np.abs(g[task + "_q"]))) < 1e-10
    if "quartic" in task:
        m = out["moments"].cpu().numpy()
        assert np.max(np.abs(m - g[task + "_moments"]) / np.maximum(np.abs(g[task + "_moments"]), 1e-3)) < TOL_STEP


@pytest.mark.parametrize("task", ["quartic", "harmonic"])
def test_in_kernel_philox_noise(task):
    """noise = NULL: the kernel draws (r0, r1) from Philox4x32-10 keyed by (seed, global trajectory id, substep counter);
    feeding the host restatement of the same stream to the oracle must reproduce the trajectory."""
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import philox_normals
    torch = _torch()
    params = configs.PRESETS[task](n_sub=10)
    B, seed, off = 6, 1234, 1000
    psi0 = initial_states(params, B, 3)
    sim = BatchedSim(params, batch=B, seed=seed, traj_offset=off)
    sim.set_state(psi0)
    actions = np.array([0, 5, 10, 15, 20, 7], np.int32)
    act = torch.as_tensor(actions, device="cuda")
    orc = oracle_for(params)
    ref = psi0
    for cstep in range(2):          # the substep counter continues across control steps
        sim.step(act)
        noise = np.array([[philox_normals(seed, off + b, cstep * 10 + s) for s in range(10)] for b in range(B)])
        ref, _, _ = oracle_control_step(orc, params, ref, actions, noise)
    torch.cuda.synchronize()
    assert rel_err(sim.get_state(), ref) < TOL_STEP


def test_result_is_independent_of_sharding():
    """Philox is keyed by the GLOBAL trajectory id: one handle with 12 trajectories == two handles with 6 (traj_offset 0 and 6)."""
    torch = _torch()
    params = configs.quartic(n_sub=8)
    psi0 = initial_states(params, 12, 9)
    act = np.arange(12, dtype=np.int32)
    full = BatchedSim(params, batch=12, seed=5)
    full.set_state(psi0)
    full.step(torch.as_tensor(act, device="cuda"))
    parts = []
    for r in range(2):
        s = BatchedSim(params, batch=6, seed=5, traj_offset=6 * r)
        s.set_state(psi0[6 * r: 6 * r + 6])
        s.step(torch.as_tensor(act[6 * r: 6 * r + 6], device="cuda"))
        parts.append(s.get_state())
    assert np.array_equal(full.get_state(), np.concatenate(parts))


def test_per_trajectory_substep_budget():
    torch = _torch()
    params = configs.quartic(n_sub=12)
    B = 5
    psi0 = initial_states(params, B, 2)
    rng = np.random.default_rng(0)
    noise = rng.standard_normal((B, 12, 2))
    budget = np.array([12, 0, 5, 1, 9], np.int32)
    actions = np.full(B, 13, np.int32)
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"), nsub_traj=torch.as_tensor(budget, device="cuda"))
    got = sim.get_state()
    orc = oracle_for(params)
    for b in range(B):
        st = psi0[b].copy()
        if budget[b]:
            orc.run(st, params["dt"], level_force(params, 13), params["gamma"], noise[b, :budget[b]])
        assert np.linalg.norm(got[b] - st) / np.linalg.norm(st) < TOL_STEP


def test_arbitrary_force_values():
    torch = _torch()
    params = configs.quartic(n_sub=6)
    B = 4
    psi0 = initial_states(params, B, 2)
    noise = np.random.default_rng(1).standard_normal((B, 6, 2))
    forces = np.array([0.123, -4.9, 2.5, 0.123])
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    sim.step_forces(forces, noise=torch.as_tensor(noise, device="cuda"))
    got = sim.get_state()
    orc = oracle_for(params)
    for b in range(B):
        st = psi0[b].copy()
        orc.run(st, params["dt"], float(forces[b]), params["gamma"], noise[b])
        assert np.linalg.norm(got[b] - st) / np.linalg.norm(st) < TOL_STEP


@pytest.mark.parametrize("npts,env", [(257, {}), (513, {}), (1025, {}), (2049, {}), (4097, {}), (8193, {})])
def test_grid_size_sweep(npts, env):
    """BASELINE.json config 5: x_max 13, N points, dt ~ h^2; tolerance check vs the oracle at every N of the sweep (257 ... 8193)."""
    torch = _torch()
    params = configs.quartic_sweep(npts, n_sub=3)
    B = 2
    psi0 = initial_states(params, B, 1)
    noise = np.random.default_rng(4).standard_normal((B, 3, 2))
    actions = np.array([2, 19], np.int32)
    sim = BatchedSim(params, batch=B)
    assert sim.n == npts
    sim.set_state(psi0)
    out = sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
    torch.cuda.synchronize()
    orc = oracle_for(params)
    ref, fails, _ = oracle_control_step(orc, params, psi0, actions, noise)
    print(npts, sim.kernel_info())
    assert rel_err(sim.get_state(), ref) < TOL_STEP
    m = out["moments"].cpu().numpy()
    mref = np.array([orc.get_moments(ref[b]) for b in range(B)])
    assert np.max(np.abs(m[:, :5] - mref[:, :5]) / np.maximum(np.abs(mref[:, :5]), 1e-3)) < 1e-9


def _episode_tool():
    import importlib.util, os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "episode_parity.py")
    spec = importlib.util.spec_from_file_location("episode_parity", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("task", ["harmonic", "quartic"])
def test_full_episode_cooling_tasks(task):
    """North-star: <= 1e-6 after a full episode with identical noise.  Both cooling tasks for the full t_max = 100 time units = 1800 control
    steps (144 000 substeps) under a quantised damping feedback; measured 5e-12 / 6e-12 (profiles/parity_episode_r02.jsonl)."""
    rec = _episode_tool().run(task, 1800, every=300)
    print(task, rec["control_steps"], "control steps:", rec["curve"])
    assert rec["control_steps"] == 1800 and not rec["ended_by_failure"]
    assert rec["max_rel_err"] < 1e-6
    assert rec["max_rel_err"] < 1e-9          # far inside the tolerance: the stable systems do not amplify rounding differences


def test_full_episode_inverted_quartic_until_termination():
    """Inverted quartic cartpole under a stabilising quantised feedback until the episode ends (escape or Fail on either side; the reference's
    episodes have no t_max, inverted quartic main_parallel.py:198-221): <= 1e-6 at every checkpoint."""
    rec = _episode_tool().run("inverted_quartic", 450, every=25)
    print(rec["control_steps"], rec["ended_by_failure"], rec["curve"])
    assert rec["control_steps"] >= 100
    assert rec["max_rel_err"] < 1e-6


def test_full_episode_inverted_harmonic_growth_is_the_physical_one():
    """The inverted harmonic oscillator amplifies ANY state difference by e^(omega t) = 1.19 per control step while both sides receive the same
    quantised force (the 21-level feedback does not react to differences below a level): 1e-15 becomes 1e-6 after ~120 control steps whatever
    the implementation -- two CPU builds of the oracle (strict vs the reference's -Ofast) separate at that same rate.  So: <= 1e-6 for the first
    100 control steps (8000 substeps), growth no faster than the physical rate afterwards, and no worse than the CPU pair."""
    tool = _episode_tool()
    rec = tool.run("inverted_harmonic", 150, every=25)
    cpu = tool.run_cpu_pair("inverted_harmonic", 150, every=25)
    print("gpu vs oracle", rec["curve"]); print("cpu strict vs cpu -Ofast", cpu["curve"])
    curve = dict(rec["curve"])
    assert curve[100] < 1e-6
    rate = np.pi / 18                                             # omega * control interval
    for k, e in rec["curve"]:
        assert e < 1e-12 * np.exp(1.08 * rate * k), (k, e)       # 8 % head-room on the exponent for the stochastic part
    cpu_curve = dict(cpu["curve"])
    for k in (75, 100, 125):
        if k in curve and k in cpu_curve:
            assert curve[k] < 30 * cpu_curve[k], (k, curve[k], cpu_curve[k])


def test_full_size_properties_config2():
    """BASELINE configs[1] at full size (1024 trajectories): normalisation, determinism, flags clear, and a random subset against
    the oracle."""
    torch = _torch()
    params = configs.quartic()
    B = 1024
    psi0 = np.tile(initial_states(params, 128, 7), (8, 1))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    act = torch.randint(0, 21, (B,), device="cuda", dtype=torch.int32, generator=g)
    runs = []
    for rep in range(2):
        sim = BatchedSim(params, batch=B, seed=77)
        sim.set_state(psi0)
        out = sim.step(act)
        torch.cuda.synchronize()
        runs.append((sim.get_state(), out["moments"].cpu().numpy(), out["aux"].cpu().numpy(), out["flags"].cpu().numpy()))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])       # bitwise deterministic
    psi, mom, aux, flags = runs[0]
    assert np.max(np.abs(np.sum(np.abs(psi) ** 2, axis=1) * params["grid_size"] - 1)) < 1e-12
    assert np.all(flags == 0) and np.all(np.isfinite(mom))
    assert np.max(np.abs(mom[:, 0] - aux[:, L.QC_AUX_XMEAN])) < 1e-12
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import philox_normals
    orc = oracle_for(params)
    a = act.cpu().numpy()
    for b in (0, 511, 1023):
        noise = np.array([philox_normals(77, b, s) for s in range(params["n_sub"])])
        st = psi0[b].copy()
        orc.run(st, params["dt"], level_force(params, int(a[b])), params["gamma"], noise)
        assert np.linalg.norm(psi[b] - st) / np.linalg.norm(st) < TOL_STEP


@pytest.mark.parametrize("task", TASKS)
def test_reference_build_fixture_through_the_abi(task):
    """tests/golden/reference_build_control_step.npz: outputs of the REFERENCE'S OWN C++ (compiled against the MKL-API shim) for one
    full control step; the CUDA path must reproduce them within the north-star tolerance (1e-10 relative per control step)."""
    import os
    torch = _torch()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_build_control_step.npz"))
    params = configs.PRESETS[task]()
    sim = BatchedSim(params, batch=g[task + "_psi0"].shape[0])
    sim.set_state(g[task + "_psi0"])
    out = sim.step(torch.as_tensor(g[task + "_actions"], device="cuda"), noise=torch.as_tensor(g[task + "_noise"], device="cuda"), want_q=True)
    torch.cuda.synchronize()
    assert rel_err(sim.get_state(), g[task + "_psi1"]) < TOL_STEP
    assert np.array_equal((out["flags"].cpu().numpy() & L.QC_FLAG_FAIL) != 0, g[task + "_fail"] != 0)
    assert np.max(np.abs(out["x_mean"].cpu().numpy() - g[task + "_xmean"])) < 1e-10
    assert np.max(np.abs(out["q"].cpu().numpy() - g[task + "_q"]) / np.maximum(1, np.abs(g[task + "_q"]))) < 1e-10
    if "quartic" in task:
        m = out["moments"].cpu().numpy()
        assert np.max(np.abs(m - g[task + "_moments"]) / np.maximum(np.abs(g[task + "_moments"]), 1e-3)) < TOL_STEP


def test_force_binning_is_bitwise_neutral(monkeypatch):
    """Large batches are grouped by force level so that a CTA stages one factor table for all its trajectories (QCART_BIN).  The grouping
    only changes which trajectories share a CTA: every output must be bitwise identical to the un-binned launch."""
    torch = _torch()
    params = configs.inverted_quartic(n_sub=6)
    B = 300
    psi0 = initial_states(params, B, 3)
    rng = np.random.default_rng(1)
    actions = rng.integers(0, 21, B).astype(np.int32)
    actions[:40] = 7                                    # a bin that is not a multiple of the CTA size, plus sparse bins
    noise = rng.standard_normal((B, 6, 2))
    res = []
    for mode in ("0", "1"):
        monkeypatch.setenv("QCART_BIN", mode)
        sim = BatchedSim(params, batch=B)
        sim.set_state(psi0)
        out = sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
        torch.cuda.synchronize()
        assert ("bin=%s" % mode) in sim.kernel_info()
        res.append((sim.get_state(), out["moments"].cpu().numpy(), out["aux"].cpu().numpy(), out["flags"].cpu().numpy()))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    orc = oracle_for(params)
    ref, _, _ = oracle_control_step(orc, params, psi0[:8], actions[:8], noise[:8])
    assert rel_err(res[1][0][:8], ref) < TOL_STEP


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["quartic", "inverted_harmonic"])
def test_solve_truncation_threshold_is_below_rounding_noise(task):
    """qc_config.solve_tol: the default (2^-48) and a 1000x stricter threshold give the same state to rounding noise, and both are
    within the 1e-10 control-step tolerance of the exact band solve of the oracle."""
    import torch
    params = configs.PRESETS[task]()
    B, n_sub = 6, params["n_sub"]
    rng = np.random.default_rng(17)
    psi0 = initial_states(params, B, 3)
    actions = rng.integers(0, params["n_levels"], B).astype(np.int32)
    noise = rng.standard_normal((B, n_sub, 2))
    ref, _, _ = oracle_control_step(oracle_for(params), params, psi0, actions, noise)
    got = {}
    for tol in (0.0, 1e-18):
        sim = BatchedSim(dict(params, solve_tol=tol), batch=B)
        assert sim.cfg.solve_tol == (2.0 ** -48 if tol == 0.0 else tol)
        sim.set_state(psi0)
        sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
        got[tol] = sim.get_state()
        err = np.max(np.linalg.norm(got[tol] - ref, axis=1) / np.linalg.norm(ref, axis=1))
        assert err < 1e-10, (tol, err)
    diff = np.max(np.linalg.norm(got[0.0] - got[1e-18], axis=1) / np.linalg.norm(ref, axis=1))
    assert diff < 1e-13, diff


@pytest.mark.gpu
def test_bad_solve_tol_is_rejected():
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import QcartError
    with pytest.raises(QcartError):
        BatchedSim(dict(configs.quartic(), solve_tol=1e-3), batch=4)


@pytest.mark.gpu
@pytest.mark.parametrize("M", [2, 3, 4])
def test_moment_orders_below_five(M):
    """MOMENT macro of the reference (Q:325): K = (M+3)M/2 moments in the reference's order; M = 2 is the north-star's 5-moment observation."""
    params, sim, out, psi_gpu, orc, psi_ref, fails, _ = run_case("quartic", 6, seed=3, n_sub=8, overrides={"moment_order": M})
    K = (M + 3) * M // 2
    assert sim.K == K and out["moments"].shape[1] == K
    assert rel_err(psi_gpu, psi_ref) < TOL_STEP
    mom = out["moments"].cpu().numpy()
    for b in range(6):
        ref = orc.get_moments(psi_ref[b])
        assert ref.shape == (K,)
        assert np.max(np.abs(mom[b] - ref) / np.maximum(np.abs(ref), 1e-3)) < TOL_STEP, (b, mom[b], ref)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2])
def test_inverted_harmonic_descriptor_modes(mode):
    """herm_mode 1 (HERMITIAN descriptor with a real diagonal) and 2 (SYMMETRIC, as harmonic/simulation.cpp:532 does) against the oracle run
    in the same mode; mode 0 (the literal reading, default) is covered by test_control_step_matches_oracle."""
    params, sim, out, psi_gpu, orc, psi_ref, fails, _ = run_case("inverted_harmonic", 8, seed=5, n_sub=20, overrides={"herm_mode": mode})
    assert rel_err(psi_gpu, psi_ref) < TOL_STEP


@pytest.mark.gpu
@pytest.mark.parametrize("task,overrides", [
    ("quartic", {"n_levels": 11, "f_max": 3.0, "dt": 1.0 / 720, "gamma": 0.05 * np.pi, "x_max": 6.0, "grid_size": 0.08, "lambda_": 0.1 * np.pi, "mass": 0.7 / np.pi}),
    ("harmonic", {"n_levels": 5, "f_max": 2.0, "dt": 1.0 / 960, "gamma": 0.3 * np.pi, "n_max": 40, "omega": 0.8 * np.pi}),
    ("inverted_harmonic", {"n_levels": 9, "f_max": 6.0, "dt": 1.0 / 2000, "gamma": 1.5 * np.pi, "n_max": 100}),
])
def test_non_default_physics_and_force_grids(task, overrides):
    """The library takes at run time what the reference bakes in with -D macros / arguments.py: other grids, masses, couplings, level counts."""
    params, sim, out, psi_gpu, orc, psi_ref, fails, _ = run_case(task, 7, seed=11, n_sub=12, overrides=overrides)
    assert rel_err(psi_gpu, psi_ref) < TOL_STEP
    flags = out["flags"].cpu().numpy()
    assert np.array_equal((flags & L.QC_FLAG_FAIL) != 0, fails != 0)


@pytest.mark.gpu
@pytest.mark.parametrize("task,B", [("inverted_harmonic", 8192), ("inverted_quartic", 8192)])
def test_full_size_properties_configs_3_and_4(task, B):
    """BASELINE configs[2] (inverted harmonic, 8192 trajectories) and configs[3] per GPU (inverted quartic, 8192 of 65 536) at full size:
    bitwise run-to-run determinism, unit norm, finite moments, shard independence of the in-kernel noise, and three trajectories against
    the oracle driven by the host restatement of the Philox stream."""
    torch = _torch()
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import philox_normals
    params = configs.PRESETS[task]()
    psi0 = np.tile(initial_states(params, 64, 9), (B // 64, 1))
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    act = torch.randint(0, 21, (B,), device="cuda", dtype=torch.int32, generator=g)
    runs = []
    for rep in range(2):
        sim = BatchedSim(params, batch=B, seed=123)
        sim.set_state(psi0)
        out = sim.step(act)
        torch.cuda.synchronize()
        runs.append((sim.get_state(), out["moments"].cpu().numpy(), out["aux"].cpu().numpy()))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    psi, mom, aux = runs[0]
    w = params["grid_size"] if "quartic" in task else 1.0
    assert np.max(np.abs(np.sum(np.abs(psi) ** 2, axis=1) * w - 1)) < 1e-12
    assert np.all(np.isfinite(mom)) and np.allclose(aux[:, L.QC_AUX_NORM], 1.0, atol=1e-12)
    # the second half of the batch as its own shard (different launch geometry): same results
    half = BatchedSim(params, batch=B // 2, seed=123, traj_offset=B // 2)
    half.set_state(psi0[B // 2:])
    half.step(act[B // 2:].contiguous())
    assert np.array_equal(half.get_state(), psi[B // 2:])
    orc = oracle_for(params)
    a = act.cpu().numpy()
    for b in (0, B // 2 + 17, B - 1):
        noise = np.array([philox_normals(123, b, s) for s in range(params["n_sub"])])
        st = psi0[b].copy()
        orc.run(st, params["dt"], level_force(params, int(a[b])), params["gamma"], noise)
        assert np.linalg.norm(psi[b] - st) / np.linalg.norm(st) < TOL_STEP


# ---- the warp-specialised pipeline kernel (csrc/qc_pipe_impl.cuh): multi-warp grid trajectories in force-binned launches ----------------

def _pipe_case(B, n_sub, seed, ragged_budget=False, one_bin=False, want_q=False, task="inverted_quartic"):
    torch = _torch()
    params = configs.PRESETS[task](n_sub=n_sub)
    rng = np.random.default_rng(seed)
    psi0 = initial_states(params, B, seed)
    actions = rng.integers(0, params["n_levels"], B).astype(np.int32)
    if one_bin:
        actions[:] = 13
    noise = rng.standard_normal((B, n_sub, 2))
    budget = rng.integers(0, n_sub + 1, B).astype(np.int32) if ragged_budget else None
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    out = sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"), want_q=want_q,
                   nsub_traj=None if budget is None else torch.as_tensor(budget, device="cuda"))
    torch.cuda.synchronize()
    return params, sim, out, psi0, actions, noise, budget


@pytest.mark.parametrize("B,one_bin", [(1400, False), (1351, False), (1400, True)])
def test_pipeline_kernel_matches_oracle(B, one_bin):
    """Inverted quartic (N = 521, three warps per trajectory) at a batch that selects sse_pipe_kernel: state, moments, flags, q stream of a
    random subset against the CPU oracle (same psi0, force, noise).  B = 1351 leaves ragged bins and a partly filled last CTA; one_bin puts
    every trajectory into one force level."""
    n_sub = 8
    params, sim, out, psi0, actions, noise, _ = _pipe_case(B, n_sub, 11, one_bin=one_bin, want_q=True)
    assert "sse_pipe_kernel" in sim.kernel_info(), sim.kernel_info()
    pick = np.random.default_rng(5).choice(B, 24, replace=False)
    pick[:3] = [0, B - 1, B // 2]
    orc = oracle_for(params)
    ref, fails, qs = oracle_control_step(orc, params, psi0[pick], actions[pick], noise[pick], want_q=True)
    got = sim.get_state()
    assert rel_err(got[pick], ref) < TOL_STEP
    assert np.allclose(out["aux"].cpu().numpy()[:, L.QC_AUX_NORM], 1.0, atol=1e-12)
    mom = out["moments"].cpu().numpy(); flags = out["flags"].cpu().numpy()
    q = out["q"].cpu().numpy(); xm = out["x_mean"].cpu().numpy()
    for k, b in enumerate(pick):
        m_ref = orc.get_moments(ref[k])
        assert np.max(np.abs(mom[b] - m_ref) / np.maximum(np.abs(m_ref), 1e-3)) < TOL_STEP
        assert bool(flags[b] & L.QC_FLAG_FAIL) == bool(fails[k])
        assert np.max(np.abs(xm[b] - qs[k][1])) < 1e-10
        assert np.max(np.abs(q[b] - qs[k][0]) / np.maximum(1.0, np.abs(qs[k][0]))) < 1e-10


def test_pipeline_kernel_agrees_with_the_per_trajectory_kernel(monkeypatch):
    """Same launch with QCART_PIPE=0 (sse_step_kernel, per-trajectory solver): every output agrees to rounding level (the chunking of the
    substitution differs, so not bitwise), flags identical; per-trajectory substep budgets honoured; both kernels deterministic run to run."""
    B, n_sub = 1400, 10
    res = {}
    for mode in ("0", "1", "1"):
        monkeypatch.setenv("QCART_PIPE", mode)
        params, sim, out, psi0, actions, noise, budget = _pipe_case(B, n_sub, 21, ragged_budget=True)
        assert ("sse_pipe_kernel" in sim.kernel_info()) == (mode == "1")
        cur = (sim.get_state(), out["moments"].cpu().numpy(), out["aux"].cpu().numpy(), out["flags"].cpu().numpy())
        if mode in res:
            for a, b in zip(res[mode], cur):
                assert np.array_equal(a, b)                     # deterministic
        res[mode] = cur
    a, b = res["0"], res["1"]
    assert rel_err(b[0], a[0]) < 1e-12
    assert np.max(np.abs(b[1] - a[1]) / np.maximum(np.abs(a[1]), 1e-3)) < TOL_STEP        # 5th-order centred moments amplify rounding
    assert np.array_equal(a[3], b[3])
    idle = np.where(budget == 0)[0]
    assert len(idle) > 0 and rel_err(b[0][idle], psi0[idle]) < 1e-14        # budget 0: state untouched (up to the renormalising store)
    orc = oracle_for(params)
    pick = np.argsort(budget)[-6:]
    for bidx in pick:
        st = psi0[bidx].copy()
        orc.run(st, params["dt"], level_force(params, int(actions[bidx])), params["gamma"], noise[bidx][: budget[bidx]])
        assert rel_err(b[0][bidx][None], st[None]) < TOL_STEP


@pytest.mark.parametrize("task", ["inverted_harmonic", "harmonic"])
def test_pipeline_kernel_fock_systems_match_oracle(task, monkeypatch):
    """The Fock instantiations of sse_pipe_kernel (inverted harmonic: two-warp groups with L = 3 incl. the HERMITIAN-descriptor term;
    harmonic: eight one-warp groups) against the CPU oracle on a random subset, with ragged per-trajectory budgets, and against the
    chunk-Jacobi kernel (QCART_PIPE=0) on everything."""
    B, n_sub = (1400 if task == "inverted_harmonic" else 2800), 10        # (harmonic: 16 trajectories per CTA, the planner wants >= one CTA per SM)
    params, sim, out, psi0, actions, noise, budget = _pipe_case(B, n_sub, 31, ragged_budget=True, want_q=True, task=task)
    assert "sse_pipe_kernel" in sim.kernel_info(), sim.kernel_info()
    got = sim.get_state(); mom = out["moments"].cpu().numpy(); aux = out["aux"].cpu().numpy(); flags = out["flags"].cpu().numpy()
    q = out["q"].cpu().numpy(); xm = out["x_mean"].cpu().numpy()
    orc = oracle_for(params)
    pick = np.random.default_rng(7).choice(B, 24, replace=False)
    pick[:2] = [0, B - 1]
    for b in pick:
        st = psi0[b].copy()
        nb = int(budget[b])
        f, qq, xx = orc.run(st, params["dt"], level_force(params, int(actions[b])), params["gamma"], noise[b][:nb], want_q=True)
        assert rel_err(got[b][None], st[None]) < TOL_STEP
        obs, nph = fock_observation(st, sim.n)
        assert np.max(np.abs(mom[b] - obs)) < 1e-9 and abs(aux[b, L.QC_AUX_ENERGY] - nph) < 1e-9
        assert bool(flags[b] & L.QC_FLAG_FAIL) == bool(f)
        if nb:
            assert np.max(np.abs(xm[b][:nb] - xx)) < 1e-10
            assert np.max(np.abs(q[b][:nb] - qq) / np.maximum(1.0, np.abs(qq))) < 1e-10
    assert np.allclose(aux[:, L.QC_AUX_NORM], 1.0, atol=1e-12)
    monkeypatch.setenv("QCART_PIPE", "0")
    _, sim0, out0, *_ = _pipe_case(B, n_sub, 31, ragged_budget=True, task=task)
    assert "sse_step_kernel" in sim0.kernel_info()
    assert rel_err(got, sim0.get_state()) < 1e-12
    assert np.array_equal(flags, out0["flags"].cpu().numpy())


@pytest.mark.parametrize("task,ne", [("quartic", "4"), ("quartic", "20"), ("harmonic", "8"), ("harmonic", "24")])
def test_one_warp_group_pipeline_variants_match_oracle(task, ne, monkeypatch):
    """One-warp explicit groups with one (NE) or two (NE + 16) solver warps per set -- the shorter recurrence of the second form is the default
    for the harmonic oscillator, and an option (QCART_PIPE=2) for config 2's grid -- against the oracle with ragged budgets, and both forms
    against each other (same chunked substitution up to the chunk length: rounding level)."""
    monkeypatch.setenv("QCART_PIPE", "2"); monkeypatch.setenv("QCART_PIPE_NE", ne)
    B, n_sub = 2800, 6
    params, sim, out, psi0, actions, noise, budget = _pipe_case(B, n_sub, 41, ragged_budget=True, want_q=True, task=task)
    want = "NE=%d,NSW=%d" % (int(ne) & 15, (int(ne) >> 4) + 1)
    assert "sse_pipe_kernel" in sim.kernel_info() and want in sim.kernel_info(), sim.kernel_info()
    got = sim.get_state(); flags = out["flags"].cpu().numpy(); xm = out["x_mean"].cpu().numpy()
    orc = oracle_for(params)
    pick = np.random.default_rng(3).choice(B, 16, replace=False)
    pick[:2] = [0, B - 1]
    for b in pick:
        st = psi0[b].copy()
        nb = int(budget[b])
        f, qq, xx = orc.run(st, params["dt"], level_force(params, int(actions[b])), params["gamma"], noise[b][:nb], want_q=True)
        assert rel_err(got[b][None], st[None]) < TOL_STEP
        assert bool(flags[b] & L.QC_FLAG_FAIL) == bool(f)
        if nb: assert np.max(np.abs(xm[b][:nb] - xx)) < 1e-10
    assert np.allclose(out["aux"].cpu().numpy()[:, L.QC_AUX_NORM], 1.0, atol=1e-12)
    monkeypatch.setenv("QCART_PIPE", "0")
    _, sim0, out0, *_ = _pipe_case(B, n_sub, 41, ragged_budget=True, task=task)
    assert "sse_step_kernel" in sim0.kernel_info()
    assert rel_err(got, sim0.get_state()) < 1e-12
    assert np.array_equal(flags, out0["flags"].cpu().numpy())


@pytest.mark.parametrize("npts,B", [(641, 340), (897, 340), (1281, 340), (1409, 340), (1793, 340), (2049, 340)])
def test_wide_grid_pipeline_matches_oracle(npts, B):
    """Single-group pipeline instances (N = 577 .. 2112): factor table in shared memory where it fits next to the lines (N <= 1536), else
    streamed from the chunk-transposed global copy: a subset of the trajectories against the oracle."""
    torch = _torch()
    params = configs.quartic_sweep(npts, n_sub=4)
    rng = np.random.default_rng(9)
    psi0 = np.tile(initial_states(params, 4, 3), (B // 4, 1))
    actions = rng.integers(0, 21, B).astype(np.int32)
    noise = rng.standard_normal((B, 4, 2))
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    out = sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
    torch.cuda.synchronize()
    # N <= 960: two groups per CTA; above: one group; N > 1536: two solver warps per trajectory
    assert "sse_pipe_kernel" in sim.kernel_info() and "NE=%d,NSW=%d" % (2 if npts <= 960 else 1, 2 if npts > 1536 else 1) in sim.kernel_info(), sim.kernel_info()
    assert ("tab=smem" in sim.kernel_info()) == (npts <= 1536), sim.kernel_info()
    pick = np.array([0, 1, B // 2, B - 1])
    orc = oracle_for(params)
    ref, fails, _ = oracle_control_step(orc, params, psi0[pick], actions[pick], noise[pick])
    assert rel_err(sim.get_state()[pick], ref) < TOL_STEP
    mom = out["moments"].cpu().numpy()
    for k, b in enumerate(pick):
        m_ref = orc.get_moments(ref[k])
        assert np.max(np.abs(mom[b][:5] - m_ref[:5]) / np.maximum(np.abs(m_ref[:5]), 1e-3)) < 1e-9


def test_wide_grid_falls_back_to_the_streamed_table_when_shared_memory_is_short():
    """N = 1409 keeps its factor table in shared memory with 320 bytes to spare at 160 substeps; a longer control step (larger noise block)
    must select the instance that streams the chunk-transposed table instead -- same results against the oracle."""
    torch = _torch()
    npts, B, n_sub = 1409, 340, 200
    params = configs.quartic_sweep(npts, n_sub=n_sub)
    rng = np.random.default_rng(23)
    psi0 = np.tile(initial_states(params, 4, 6), (B // 4, 1))
    actions = rng.integers(0, 21, B).astype(np.int32)
    noise = rng.standard_normal((B, n_sub, 2))
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
    torch.cuda.synchronize()
    assert "sse_pipe_kernel" in sim.kernel_info() and "G=256,NE=1" in sim.kernel_info() and "tab=smem" not in sim.kernel_info(), sim.kernel_info()
    pick = np.array([0, B - 1])
    ref, _, _ = oracle_control_step(oracle_for(params), params, psi0[pick], actions[pick], noise[pick])
    assert rel_err(sim.get_state()[pick], ref) < TOL_STEP


def test_transposed_factor_table_follows_on_demand_forces():
    """N = 1793 streams its factor rows from the chunk-transposed copy of the table (fac_transpose_kernel).  Forces outside the 21 levels are
    factorised on demand (qc_step_forces): the copy must be rebuilt, also when a slot is replaced; checked against the oracle."""
    torch = _torch()
    npts, B, n_sub = 1793, 340, 3
    params = configs.quartic_sweep(npts, n_sub=n_sub)
    rng = np.random.default_rng(19)
    psi0 = np.tile(initial_states(params, 4, 5), (B // 4, 1))
    sim = BatchedSim(params, batch=B)
    sim.set_state(psi0)
    ref = psi0[:4].copy()
    orc = oracle_for(params)
    pick = np.array([0, 1, 2, 3])
    for it in range(3):
        noise = rng.standard_normal((B, n_sub, 2))
        if it == 1:
            act = rng.integers(0, 21, B).astype(np.int32)
            forces = np.array([level_force(params, int(a)) for a in act])
            sim.step(torch.as_tensor(act, device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
        else:
            forces = np.repeat(rng.uniform(-4.0, 4.0, 10), B // 10)      # 10 new force values per call, 34 trajectories each
            sim.step_forces(forces, noise=torch.as_tensor(noise, device="cuda"))
        torch.cuda.synchronize()
        assert "sse_pipe_kernel" in sim.kernel_info() and "tab=smem" not in sim.kernel_info(), sim.kernel_info()
        for k, b in enumerate(pick):
            orc.run(ref[k], params["dt"], float(forces[b]), params["gamma"], noise[b])
        assert rel_err(sim.get_state()[pick], ref) < TOL_STEP * (it + 1)


@pytest.mark.parametrize("npts", [2501, 4097])
def test_cluster_kernel_matches_oracle(npts):
    """One trajectory per thread-block cluster (N > 2112): state, all 20 moments, energy, <x>, q / <x> streams and flags against the oracle,
    with ragged per-trajectory substep budgets; bitwise deterministic run to run; agrees with the single-CTA instance (QCART_CLUSTER=0 in
    the other tests' history) to rounding level."""
    torch = _torch()
    n_sub = 5
    params = configs.quartic_sweep(npts, n_sub=n_sub)
    B = 5
    rng = np.random.default_rng(13)
    psi0 = initial_states(params, B, 8)
    actions = np.array([0, 6, 10, 15, 20], np.int32)
    noise = rng.standard_normal((B, n_sub, 2))
    budget = np.array([5, 0, 3, 5, 1], np.int32)
    runs = []
    for rep in range(2):
        sim = BatchedSim(params, batch=B)
        sim.set_state(psi0)
        out = sim.step(torch.as_tensor(actions, device="cuda"), noise=torch.as_tensor(noise, device="cuda"), want_q=True,
                       nsub_traj=torch.as_tensor(budget, device="cuda"))
        torch.cuda.synchronize()
        assert "sse_cluster_kernel" in sim.kernel_info(), sim.kernel_info()
        runs.append((sim.get_state(), out["moments"].cpu().numpy(), out["aux"].cpu().numpy(), out["flags"].cpu().numpy(), out["q"].cpu().numpy(), out["x_mean"].cpu().numpy()))
    for a, b in zip(runs[0][:4], runs[1][:4]):
        assert np.array_equal(a, b)
    got, mom, aux, flags, q, xm = runs[0]
    orc = oracle_for(params)
    for b in range(B):
        st = psi0[b].copy()
        nb = int(budget[b])
        f, qq, xx = orc.run(st, params["dt"], level_force(params, int(actions[b])), params["gamma"], noise[b][:nb], want_q=True)
        assert rel_err(got[b][None], st[None]) < TOL_STEP
        m_ref = orc.get_moments(st)
        assert np.max(np.abs(mom[b] - m_ref) / np.maximum(np.abs(m_ref), 1e-3)) < TOL_STEP
        assert abs(aux[b, L.QC_AUX_XMEAN] - orc.x_expectation(st)) < 1e-10 and abs(aux[b, L.QC_AUX_NORM] - 1) < 1e-12
        assert bool(flags[b] & L.QC_FLAG_FAIL) == bool(f)
        if nb:
            assert np.max(np.abs(xm[b][:nb] - xx)) < 1e-10
            assert np.max(np.abs(q[b][:nb] - qq) / np.maximum(1.0, np.abs(qq))) < 1e-10
