"""GPU tests of the boundary: the `simulation`-module mirror (reference signatures and error behaviour), host-buffer entry point,
flags, and the env API."""
import numpy as np
import pytest
from math import pi

from common import oracle_for, initial_states, level_force
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs, BatchedSim, QuantumCartpoleEnv, simulation, _lib as L

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def test_simulation_module_mirror_quartic():
    params = configs.quartic()
    simulation.configure(params)
    assert simulation.check_settings() == (171, 0.1, params["lambda_"], params["mass"], 5)       # Q:652-654
    orc = oracle_for(params)
    psi = initial_states(params, 1, 4)[0]
    ref = psi.copy()
    rng = np.random.default_rng(2)
    for F in (0.0, 2.5, 2.5, -5.0, 0.37):
        r = rng.standard_normal(2)
        q, xm, fail = simulation.step(psi, params["dt"], F, params["gamma"], normals=r)           # in place, like the reference
        q2, xm2, fail2 = orc.step(ref, params["dt"], F, params["gamma"], r)
        assert np.linalg.norm(psi - ref) / np.linalg.norm(ref) < 1e-10
        assert abs(q - q2) < 1e-10 * max(1, abs(q2)) and abs(xm - xm2) < 1e-10 and fail == fail2
    out = np.empty(20)
    assert simulation.get_moments(psi, out) is None
    mref = orc.get_moments(ref)
    assert np.max(np.abs(out - mref) / np.maximum(np.abs(mref), 1e-3)) < 1e-9
    assert abs(simulation.x_expectation(psi) - orc.x_expectation(ref)) < 1e-10
    # simulate_10_steps (Q:526-558)
    r10 = rng.standard_normal((10, 2))
    q, xm, fail = simulation.simulate_10_steps(psi, params["dt"], 1.5, params["gamma"], normals=r10)
    f2, qs, xms = orc.run(ref, params["dt"], 1.5, params["gamma"], r10, want_q=True)
    assert np.linalg.norm(psi - ref) / np.linalg.norm(ref) < 1e-10 and abs(q - qs[-1]) < 1e-9 * max(1, abs(qs[-1])) and abs(xm - xms[-1]) < 1e-10
    # seeded internal stream is reproducible
    a, b = initial_states(params, 1, 4)[0], initial_states(params, 1, 4)[0]
    simulation.set_seed(99); simulation.step(a, params["dt"], 0.0, params["gamma"])
    simulation.set_seed(99); simulation.step(b, params["dt"], 0.0, params["gamma"])
    assert np.array_equal(a, b)


def test_simulation_module_error_behaviour():
    """check_type / check_moment_data_array of the reference (Q:288-323): TypeError for non-arrays, ValueError for shape/dtype."""
    params = configs.quartic()
    simulation.configure(params)
    good = initial_states(params, 1, 0)[0]
    with pytest.raises(TypeError):
        simulation.step([0j] * 171, params["dt"], 0.0, params["gamma"])
    with pytest.raises(ValueError, match="one-dimensional"):
        simulation.step(np.zeros((1, 171), np.complex128), params["dt"], 0.0, params["gamma"])
    with pytest.raises(ValueError, match="required size 171"):
        simulation.step(np.zeros(170, np.complex128), params["dt"], 0.0, params["gamma"])
    with pytest.raises(ValueError, match="Complex128"):
        simulation.step(np.zeros(171, np.float64), params["dt"], 0.0, params["gamma"])
    with pytest.raises(ValueError, match="required size 20"):
        simulation.get_moments(good, np.empty(19))
    with pytest.raises(ValueError, match="Float64"):
        simulation.get_moments(good, np.empty(20, np.float32))
    with pytest.raises(TypeError):
        simulation.step(good, "dt", 0.0, params["gamma"])


def test_simulation_module_mirror_fock():
    for task in ("harmonic", "inverted_harmonic"):
        params = configs.PRESETS[task]()
        simulation.configure(params)
        assert simulation.check_settings() == (params["n_max"], params["omega"])                 # H:562-564
        orc = oracle_for(params)
        psi = initial_states(params, 1, 3)[0]
        ref = psi.copy()
        rng = np.random.default_rng(5)
        for F in (0.0, 4.0, -2.5):
            r = rng.standard_normal(2)
            q, xm, fail = simulation.step(psi, params["dt"], F, params["gamma"], normals=r)
            q2, xm2, fail2 = orc.step(ref, params["dt"], F, params["gamma"], r)
            assert np.linalg.norm(psi - ref) / np.linalg.norm(ref) < 1e-10 and abs(xm - xm2) < 1e-10 and fail == fail2
        assert abs(simulation.x_expectation(psi) - orc.x_expectation(ref)) < 1e-10


def test_simulation_module_diagnostics_all_systems():
    """Hamiltonian_dot_psi / solve_ab (harmonic simulation.cpp:566-597, simulation_i.cpp:585-616) through the C-ABI against the oracle
    (pivoted LAPACK-style LU) and, where it was built, the reference's own module.  solve_ab is the exact band substitution: 1e-12."""
    from oracle.ref_module import RefModule, available
    rng = np.random.default_rng(21)
    for task in ("harmonic", "inverted_harmonic", "quartic", "inverted_quartic"):
        params = configs.PRESETS[task]()
        simulation.configure(params)
        orc = oracle_for(params)
        n = orc.n
        psi = initial_states(params, 1, 2)[0]
        with pytest.raises(RuntimeError):
            simulation.solve_ab(psi.copy())                       # no factorisation before the first step
        for F in (level_force(params, 19), 0.0, -1.234567):
            r = rng.standard_normal(2)
            st = psi.copy()
            simulation.step(st, params["dt"], F, params["gamma"], normals=r)
            v = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * np.exp(-0.02 * np.arange(n))
            x_gpu, x_orc = v.copy(), v.copy()
            assert simulation.solve_ab(x_gpu) == 0.0
            orc.solve_ab(params["dt"], F, x_orc)
            assert np.linalg.norm(x_gpu - x_orc) / np.linalg.norm(x_orc) < 1e-12
            assert np.linalg.norm(orc.A_dense(params["dt"], F) @ x_gpu - v) / np.linalg.norm(v) < 1e-12
            if task == "harmonic" and available(task):            # (simulation_i.cpp:613 solves with the wrong band width: not a reference)
                ref = RefModule(task)
                ref.step(psi.copy(), params["dt"], F, params["gamma"], r)
                x_ref = v.copy(); ref.mod.solve_ab(x_ref)
                assert np.linalg.norm(x_gpu - x_ref) / np.linalg.norm(x_ref) < 1e-12
        h_gpu, h_orc = v.copy(), v.copy()
        assert simulation.Hamiltonian_dot_psi(h_gpu) == 0.0
        orc.hamiltonian_dot_psi(h_orc)
        assert np.linalg.norm(h_gpu - h_orc) / np.linalg.norm(h_orc) < 1e-13
        with pytest.raises(ValueError):
            simulation.Hamiltonian_dot_psi(np.zeros(n - 1, np.complex128))


def test_host_buffer_entry_point_equals_device_entry_point():
    torch = _torch()
    params = configs.quartic(n_sub=8)
    B = 64
    psi0 = initial_states(params, B, 1)
    act = np.random.default_rng(0).integers(0, 21, B).astype(np.int32)
    a = BatchedSim(params, batch=B, seed=3); a.set_state(psi0)
    b = BatchedSim(params, batch=B, seed=3); b.set_state(psi0)
    out = a.step(torch.as_tensor(act, device="cuda"))
    mom, aux, flags = b.step_host(act)
    assert np.array_equal(out["moments"].cpu().numpy(), mom) and np.array_equal(out["aux"].cpu().numpy(), aux)
    assert np.array_equal(out["flags"].cpu().numpy(), flags)
    assert np.array_equal(a.get_state(), b.get_state())


def test_host_entry_point_pinned_buffers_equal_pageable_buffers():
    """qc_step_host writes page-locked result buffers from inside the kernel (mapped alias, mirror_row) and copies into pageable ones:
    bitwise the same rows, for the per-trajectory kernel, the pipeline kernel and the cluster kernel, and with only some buffers pinned."""
    torch = _torch()
    cases = [(configs.quartic(n_sub=6), 300), (configs.inverted_quartic(n_sub=4), 1300), (configs.inverted_harmonic(n_sub=5), 1500),
             (configs.quartic_sweep(2501, n_sub=2), 5)]
    for params, B in cases:
        psi0 = initial_states(params, min(B, 64), 1)
        psi0 = np.tile(psi0, ((B + psi0.shape[0] - 1) // psi0.shape[0], 1))[:B]
        act = np.random.default_rng(1).integers(0, params["n_levels"], B).astype(np.int32)
        a = BatchedSim(params, batch=B, seed=5); a.set_state(psi0)
        b = BatchedSim(params, batch=B, seed=5); b.set_state(psi0)
        c = BatchedSim(params, batch=B, seed=5); c.set_state(psi0)
        mom_p = torch.full((B, a.K), float("nan"), dtype=torch.float64).pin_memory()
        aux_p = torch.full((B, L.QC_AUX_COUNT), float("nan"), dtype=torch.float64).pin_memory()
        flg_p = torch.full((B,), 255, dtype=torch.uint8).pin_memory()
        for _ in range(2):
            a.step_host(act, moments=mom_p, aux=aux_p, flags=flg_p)
            mom, aux, flags = b.step_host(act)
            mom_c, aux_c, flg_c = c.step_host(act, aux=aux_p.clone().pin_memory())            # mixed: only aux pinned
            assert np.array_equal(mom_p.numpy(), mom) and np.array_equal(aux_p.numpy(), aux) and np.array_equal(flg_p.numpy(), flags), a.kernel_info()
            assert np.array_equal(mom_c, mom) and np.array_equal(aux_c.numpy(), aux) and np.array_equal(flg_c, flags)
        assert np.array_equal(a.get_state(), b.get_state())
        # flags alone (no moments, no aux): the kernels skip their output stage, the pinned flag buffer is filled by a copy
        flg_only = torch.full((B,), 255, dtype=torch.uint8).pin_memory()
        L.check(a.lib.qc_step_host(a.h, act.ctypes.data, None, 0, None, None, flg_only.data_ptr()))
        L.check(b.lib.qc_step_host(b.h, act.ctypes.data, None, 0, None, None, flags.ctypes.data))
        assert np.array_equal(flg_only.numpy(), flags) and flags.max() < 255


def test_fail_and_escape_flags():
    torch = _torch()
    # numerical Fail: amplitude at the grid boundary (Q:559-565); oracle decides
    params = configs.quartic(n_sub=4)
    orc = oracle_for(params)
    x = orc.x_array()
    psi = np.exp(-(x - 7.9) ** 2 / 0.5).astype(np.complex128)
    psi /= np.sqrt(np.sum(np.abs(psi) ** 2) * params["grid_size"])
    ok = initial_states(params, 1, 0)[0]
    batch = np.stack([psi, ok])
    noise = np.zeros((2, 4, 2))
    sim = BatchedSim(params, batch=2); sim.set_state(batch)
    out = sim.step(torch.as_tensor(np.array([10, 10], np.int32), device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
    st = psi.copy(); f_ref, _, _ = orc.run(st, params["dt"], 0.0, params["gamma"], noise[0])
    flags = out["flags"].cpu().numpy()
    assert f_ref == 1 and (flags[0] & L.QC_FLAG_FAIL) and not (flags[1] & L.QC_FLAG_FAIL)
    # latched until cleared
    out = sim.step(torch.as_tensor(np.array([10, 10], np.int32), device="cuda"), noise=torch.as_tensor(noise, device="cuda"))
    assert out["flags"].cpu().numpy()[0] & L.QC_FLAG_FAIL
    # escape: inverted quartic, packet far outside |x| < x_th = 5 (IQ/main_parallel.py:78-81,199-200)
    p2 = configs.inverted_quartic(n_sub=2)
    sim2 = BatchedSim(p2, batch=2)
    sim2.init_packets(mean=np.array([0.0, 7.0]))
    out2 = sim2.step(torch.zeros(2, dtype=torch.int32, device="cuda") + 10)
    f2 = out2["flags"].cpu().numpy()
    aux2 = out2["aux"].cpu().numpy()
    assert not (f2[0] & L.QC_FLAG_ESCAPED) and (f2[1] & L.QC_FLAG_ESCAPED)
    assert aux2[0, L.QC_AUX_OUTSIDE] < 0.01 and aux2[1, L.QC_AUX_OUTSIDE] > 0.9


def test_initial_state_builders_match_reference_formulas():
    params = configs.quartic()
    sim = BatchedSim(params, batch=3)
    k = np.array([0.0, 0.2, -0.3]); mean = np.array([0.0, 1.0, -0.5])
    sim.init_packets(wavenumber=k, mean=mean, std=1.0)
    got = sim.get_state()
    x = sim.x_grid()
    for b in range(3):       # Gaussian_packet, Q/main_parallel.py:75-76 with wavelength = 1/k
        ref = np.exp(2j * pi * (x - mean[b]) * k[b]) * np.exp(-(x - mean[b]) ** 2 / 4) / np.sqrt(np.sqrt(2 * pi))
        assert np.max(np.abs(got[b] - ref)) < 1e-14
    f = BatchedSim(configs.harmonic(), batch=2)
    f.init_fock(None)
    s = f.get_state()
    assert np.all(s[:, 0] == 1.0) and np.all(s[:, 1:] == 0)


@pytest.mark.parametrize("task", ["quartic", "inverted_quartic", "harmonic", "inverted_harmonic"])
def test_env_reset_step_protocol(task):
    torch = _torch()
    kw = {"batch": 16, "seed": 1}
    env = QuantumCartpoleEnv(task, **kw)
    if task == "quartic":
        # shorten the 15-20 time-unit warm-up for the test by monkeypatching the rng range is not needed: 16 trajectories, ~28 launches
        pass
    obs = env.reset()
    assert obs.shape == (16, env.observation_size()) and obs.dtype == torch.float32 and torch.isfinite(obs).all()
    zero = torch.full((16,), env.zero_action, dtype=torch.int64, device=env.dev)
    for _ in range(3):
        obs, reward, done, info = env.step(zero)
    assert reward.shape == (16,) and done.shape == (16,) and done.dtype == torch.bool
    if task == "quartic":
        assert torch.all(info["energy"] < 7.6 + 5) and torch.allclose(reward.double(), -info["energy"], atol=1e-5)
        assert abs(float(env.t[0]) - 3 * 80 / 1440) < 1e-9
    if task == "harmonic":
        assert torch.allclose(reward.double(), -10 * info["energy"], atol=1e-4)
    if task in ("inverted_quartic", "inverted_harmonic"):
        assert set(reward.cpu().numpy().tolist()) <= {1.0, -1.0}
    # pushing hard in one direction eventually terminates the inverted tasks
    if task == "inverted_quartic":
        push = torch.full((16,), 20, dtype=torch.int64, device=env.dev)
        for _ in range(40):
            obs, reward, done, info = env.step(push)
            if bool(done.all()):
                break
        assert bool(done.all()) and float(reward.max()) == -1.0


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["quartic", "harmonic", "inverted_quartic_pipeline"])
def test_fused_result_exchange_single_process_ranks(task):
    """qc_set_gather: three sims on one device play three ranks; every rank's gather area ends up with all rows, in rank order, equal to
    pack_block of each rank's own outputs, for six consecutive control steps (four buffers, plain and overlapped consumer schedule); the
    bounded wait reports ranks that never publish; unequal batch sizes are refused."""
    import torch
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import dist as qdist
    B, world = 40, 3
    if task == "inverted_quartic_pipeline":                    # a batch that selects sse_pipe_kernel: its epilogue publishes the rows too
        task, B = "inverted_quartic", 1400
    params = configs.PRESETS[task](n_sub=8)
    sims = [BatchedSim(params, batch=B, seed=3, traj_offset=r * B) for r in range(world)]
    for r, s in enumerate(sims):
        s.set_state(np.tile(initial_states(params, 40, seed=10 + r), ((B + 39) // 40, 1))[:B])
    fgs = qdist.FusedGather.local_group(sims)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    expect = {}
    for step in range(1, 7):                                   # six steps: every one of the four buffers is reused
        packed = []
        for s in sims:
            act = torch.randint(0, params["n_levels"], (B,), device="cuda", dtype=torch.int32, generator=g)
            out = s.step(act)
            packed.append(qdist.pack_block(out["moments"], out["aux"], out["flags"]))
        expect[step] = torch.cat(packed, dim=0)
        if B > 1000:
            assert "sse_pipe_kernel" in sims[0].kernel_info()
        for fg in fgs:
            assert fg.seq() == step
            if step % 2 == 1:                                  # plain schedule: wait for the step just enqueued
                fg.wait()
                torch.cuda.synchronize()
                assert torch.equal(fg.block(), expect[step])
            else:                                              # overlapped schedule: consume step k-1 behind step k (still intact in its buffer)
                fg.wait(step - 1)
                torch.cuda.synchronize()
                assert torch.equal(fg.block(step - 1), expect[step - 1])
    assert all(fg.failed_ranks() == [] for fg in fgs)
    fgs[0].wait(fgs[0].seq() + 1)                              # a step nobody will ever publish: the consumer gives up (bounded spin) and says who
    torch.cuda.synchronize()
    assert fgs[0].failed_ranks() == [0, 1, 2]
    with pytest.raises(ValueError):                            # unequal shards would address peer memory out of bounds
        qdist.FusedGather.local_group([sims[0], BatchedSim(params, batch=B + 1)])
    with pytest.raises(L.QcartError):                      # the exchange needs all three output buffers
        L.check(sims[0].lib.qc_step(sims[0].h, act.data_ptr(), None, 8, None, None, None, None, None, None, sims[0]._stream()))
    for fg in fgs:
        fg.close()
    out = sims[0].step(act)                                     # exchange switched off again: plain step works
    assert torch.isfinite(out["moments"]).all()


@pytest.mark.gpu
def test_edge_cases_empty_ragged_and_oversized():
    """Empty batch is an argument error; a batch that does not fill its last CTA gives the same per-trajectory results as any other
    batch size (bitwise, with supplied noise); a grid beyond the resident-kernel limit is reported, not mis-computed."""
    import torch
    params = configs.quartic(n_sub=6)
    empty = BatchedSim(params)
    with pytest.raises(L.QcartError) as e:
        empty.set_batch(0)
    assert e.value.code == L.QC_ERR_ARG
    rng = np.random.default_rng(2)
    Bbig = 1031                                                   # prime: ragged against every CTA packing
    psi0 = initial_states(params, 16, seed=4)
    psi = np.tile(psi0, (Bbig // 16 + 1, 1))[:Bbig]
    act = rng.integers(0, params["n_levels"], Bbig).astype(np.int32)
    noise = rng.standard_normal((Bbig, 6, 2))
    got = {}
    for B in (Bbig, 37):
        sim = BatchedSim(params, batch=B)
        sim.set_state(psi[:B])
        sim.step(torch.as_tensor(act[:B], device="cuda"), noise=torch.as_tensor(noise[:B], device="cuda"))
        got[B] = sim.get_state()
    assert np.array_equal(got[Bbig][:37], got[37])
    # 9473 points: beyond what one CTA can hold (9216) -- a cluster of 8 CTAs takes it (up to 10 752); compare with the oracle
    mid = configs.quartic_sweep(9473, n_sub=3)
    sim = BatchedSim(mid, batch=2)
    p0 = initial_states(mid, 2, seed=6)
    sim.set_state(p0)
    nz = np.random.default_rng(3).standard_normal((2, 3, 2))
    sim.step(torch.as_tensor(np.array([4, 17], np.int32), device="cuda"), noise=torch.as_tensor(nz, device="cuda"))
    assert "sse_cluster_kernel" in sim.kernel_info() and "C=8" in sim.kernel_info()
    from common import oracle_for, oracle_control_step
    ref, _, _ = oracle_control_step(oracle_for(mid), mid, p0, np.array([4, 17], np.int32), nz)
    got = sim.get_state()
    assert float(np.max(np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1))) < 1e-10
    big = configs.quartic_sweep(10801)                             # beyond 8 CTAs x 224 lanes x 6 points: reported, not mis-computed
    with pytest.raises(L.QcartError) as e:
        sim = BatchedSim(big, batch=2)
        sim.step(torch.zeros(2, dtype=torch.int32, device="cuda"))
    assert e.value.code == L.QC_ERR_UNSUPPORTED
