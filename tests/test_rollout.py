"""The device-resident steps either side of the SSE kernel (include/qcart_rollout.h) against the CPU restatement in
oracle/rollout_oracle.py and the golden vectors of the reference's own `direct_DQN` / layers.py (tests/golden/policy_reference_python.npz).

Floating point: the policy is fp32 like the reference's torch modules; summation order differs, so action values are compared to
2e-5 * max|q| (a few fp32 ulps of a 512-term dot product) and argmax is required to agree wherever the top-two gap exceeds that bound.
The experience rows and the measurement record are float32 COPIES / means of float64 values: bit-exact.
"""
import os

import numpy as np
import pytest

from deepreinforcementlearningcontrolofquantumcartpoles_b200 import _lib as L, configs
from deepreinforcementlearningcontrolofquantumcartpoles_b200 import rollout as R
from oracle import rollout_oracle as RO

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "policy_reference_python.npz"))
QTOL = 2e-5


# ---- CPU: oracle against the reference's own classes, host helpers, error paths ---------------------------------------------------------
@pytest.mark.parametrize("tag,n_in,seed", [("grid", 20, 101), ("fock", 5, 102)])
def test_oracle_policy_matches_reference_classes(tag, n_in, seed):
    sd = RO.policy_state_dict(seed, n_in=n_in)
    a, m = RO.direct_dqn_forward(sd, GOLD[tag + "_x"], (GOLD[tag + "_fc31_rand_in"], GOLD[tag + "_fc31_rand_out"]),
                                 (GOLD[tag + "_fc41_rand_in"], GOLD[tag + "_fc41_rand_out"]))
    scale = np.abs(GOLD[tag + "_action_values"]).max()
    assert np.abs(a - GOLD[tag + "_action_values"]).max() <= QTOL * scale
    assert np.abs(m - GOLD[tag + "_mean"]).max() <= QTOL * max(1.0, np.abs(GOLD[tag + "_mean"]).max())
    assert (a.argmax(1) == GOLD[tag + "_argmax"]).all()
    assert np.allclose([RO.convert_to_force(n, 5.0) for n in range(21)], GOLD[tag + "_force_of_action"], rtol=0, atol=1e-15)


def test_fold_weight_norm_and_epsilon_schedule():
    sd = RO.policy_state_dict(5)
    w = R.fold_weight_norm(sd["fc2.weight"], sd["fc2.weight_norm"])
    assert w.dtype == np.float32
    assert abs(np.linalg.norm(w.astype(np.float64)) - float(sd["fc2.weight_norm"])) <= 1e-5 * float(sd["fc2.weight_norm"])
    for n in (0, 1000, 10 ** 6):
        assert R.epsilon_threshold(n, 0.2, 0.004, 18 * 100 * 80) == RO.epsilon_threshold(n, 0.2, 0.004, 18 * 100 * 80)


def test_measurement_record_sizes_follow_the_reference():
    p = configs.harmonic()
    # harmonic main_parallel.py:142-148 at time_steps = 1440, n_con = 18
    read_length = round(1.5 * 2 * 1440)
    assert (read_length, 1440 // 1440, p["n_sub"] // 1) == (4320, 1, 80)


def test_rollout_entry_points_fail_loudly_without_a_device():
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-box check")
    lib = L.load()
    h = C.c_void_p()
    assert lib.qc_policy_create(20, 21, 2, 0, C.byref(h)) == L.QC_ERR_CUDA
    assert b"no usable CUDA device" in lib.qc_last_error()
    assert lib.qc_replay_create(42, 1024, 0, C.byref(h)) == L.QC_ERR_CUDA
    assert lib.qc_record_create(4, 4320, 1, 80, 0, C.byref(h)) == L.QC_ERR_CUDA
    with pytest.raises(L.QcartError):
        R.DirectDQNPolicy(20)


# ---- GPU -------------------------------------------------------------------------------------------------------------------------------
def _policy(seed, n_in=20, n_actions=21):
    pol = R.DirectDQNPolicy(n_in, n_actions)
    sd = RO.policy_state_dict(seed, n_in=n_in, n_actions=n_actions)
    pol.load_state_dict(sd)
    return pol, sd


def _agree(q_gpu, q_ref, greedy):
    scale = max(1.0, np.abs(q_ref).max())
    assert np.abs(q_gpu - q_ref).max() <= QTOL * scale, np.abs(q_gpu - q_ref).max()
    top2 = np.sort(q_ref, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 4 * QTOL * scale
    assert (greedy[clear] == q_ref.argmax(1)[clear]).all()
    assert (greedy == q_gpu.argmax(1)).all()                  # the kernel's argmax is the argmax of the values it reports


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n_in,seed", [("grid", 20, 101), ("fock", 5, 102)])
def test_policy_matches_reference_golden_vectors(tag, n_in, seed):
    import torch
    pol, sd = _policy(seed, n_in)
    x = torch.as_tensor(GOLD[tag + "_x"], device="cuda")
    noise = pol.pack_noise(GOLD[tag + "_fc31_rand_in"], GOLD[tag + "_fc31_rand_out"], GOLD[tag + "_fc41_rand_in"], GOLD[tag + "_fc41_rand_out"])
    out = pol.forward(x, noise=noise, want_value=True)
    _agree(out["q"].cpu().numpy(), GOLD[tag + "_action_values"], out["greedy"].cpu().numpy())
    assert np.abs(out["value"].cpu().numpy() - GOLD[tag + "_mean"]).max() <= QTOL * max(1.0, np.abs(GOLD[tag + "_mean"]).max())
    assert (out["greedy"].cpu().numpy() == GOLD[tag + "_argmax"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 63, 64, 65, 1000])
def test_policy_matches_oracle_ragged_batches(B):
    import torch
    pol, sd = _policy(7)
    rng = np.random.Generator(np.random.PCG64(B))
    x = (rng.standard_normal((B, 20)) * 1.5).astype(np.float32)
    n = [RO.noisy_f(rng.standard_normal(s)).astype(np.float32) for s in ((B, 512), (B, 256), (B, 256), (B, 21))]
    out = pol.forward(torch.as_tensor(x, device="cuda"), noise=pol.pack_noise(*n), want_value=True)
    a, m = RO.direct_dqn_forward(sd, x, (n[0], n[1]), (n[2], n[3]))
    _agree(out["q"].cpu().numpy(), a, out["greedy"].cpu().numpy())
    assert np.abs(out["value"].cpu().numpy() - m).max() <= QTOL * max(1.0, np.abs(m).max())
    # noisy = False: F.linear(x, u_w, u_b)
    out0 = pol.forward(torch.as_tensor(x, device="cuda"), noise=None)
    a0, _ = RO.direct_dqn_forward(sd, x, None, None)
    _agree(out0["q"].cpu().numpy(), a0, out0["greedy"].cpu().numpy())


@pytest.mark.gpu
def test_tensor_core_and_cuda_core_gemm_agree():
    """The tcgen05 3xTF32 kernel and the fp32 CUDA-core kernel give the same action values to fp32 round-off, both within tolerance of the oracle."""
    import torch
    pol, sd = _policy(11)
    B = 700                                                       # ragged: 5 full 128-row tiles + 60 rows
    rng = np.random.Generator(np.random.PCG64(77))
    x = (rng.standard_normal((B, 20)) * 1.5).astype(np.float32)
    n = [RO.noisy_f(rng.standard_normal(s)).astype(np.float32) for s in ((B, 512), (B, 256), (B, 256), (B, 21))]
    a, m = RO.direct_dqn_forward(sd, x, (n[0], n[1]), (n[2], n[3]))
    res = {}
    for kind in ("tcgen05", "simt", "tcgen05_staged", "tcgen05_raw"):
        pol.set_gemm(kind)
        out = pol.forward(torch.as_tensor(x, device="cuda"), noise=pol.pack_noise(*n), want_value=True)
        res[kind] = out["q"].cpu().numpy()
        _agree(res[kind], a, out["greedy"].cpu().numpy())
        assert np.abs(out["value"].cpu().numpy() - m).max() <= QTOL * max(1.0, np.abs(m).max())
    assert np.abs(res["tcgen05"] - res["simt"]).max() <= 0.25 * QTOL * max(1.0, np.abs(a).max()), np.abs(res["tcgen05"] - res["simt"]).max()
    # TMA-fed and thread-staged kernels issue the same tensor-core instructions on the same split operands: bitwise equal
    assert np.array_equal(res["tcgen05"], res["tcgen05_staged"]), np.abs(res["tcgen05"] - res["tcgen05_staged"]).max()
    assert np.array_equal(res["tcgen05_raw"], res["tcgen05_staged"]), np.abs(res["tcgen05_raw"] - res["tcgen05_staged"]).max()


@pytest.mark.gpu
def test_policy_unset_parameter_and_bad_sizes_are_errors():
    pol = R.DirectDQNPolicy(20)
    import torch
    x = torch.zeros((4, 20), dtype=torch.float32, device="cuda")
    with pytest.raises(L.QcartError) as e:
        pol.forward(x)
    assert e.value.code == L.QC_ERR_STATE
    with pytest.raises(L.QcartError) as e:
        pol.set_param("FC2_W", np.zeros(7, np.float32))
    assert e.value.code == L.QC_ERR_ARG


@pytest.mark.gpu
def test_philox_noise_is_shard_independent_and_has_the_right_law():
    import torch
    pol, sd = _policy(3)
    B = 4096
    rng = np.random.Generator(np.random.PCG64(1))
    x = torch.as_tensor(rng.standard_normal((B, 20)).astype(np.float32), device="cuda")
    full = pol.forward(x, noise="philox", seed=11, traj_offset=0, counter=5)["q"]
    lo = pol.forward(x[:1000].contiguous(), noise="philox", seed=11, traj_offset=0, counter=5)["q"]
    hi = pol.forward(x[1000:].contiguous(), noise="philox", seed=11, traj_offset=1000, counter=5)["q"]
    assert torch.equal(full[:1000], lo) and torch.equal(full[1000:], hi)          # keyed by the global trajectory id
    other = pol.forward(x, noise="philox", seed=11, traj_offset=0, counter=6)["q"]
    assert not torch.equal(full, other)
    # law of the noisy output: mean = noiseless value; spread consistent with the oracle under fresh Gaussian noise
    q0 = pol.forward(x, noise=None)["q"]
    xs = x[:1].repeat(B, 1).contiguous()
    qs = pol.forward(xs, noise="philox", seed=2, counter=0)["q"].cpu().numpy()
    n = [RO.noisy_f(rng.standard_normal(s)).astype(np.float32) for s in ((B, 512), (B, 256), (B, 256), (B, 21))]
    qo, _ = RO.direct_dqn_forward(sd, xs.cpu().numpy(), (n[0], n[1]), (n[2], n[3]))
    assert np.abs(qs.mean(0) - qo.mean(0)).max() < 6 * qo.std(0).max() / np.sqrt(B) * 2
    assert np.abs(qs.std(0) / qo.std(0) - 1).max() < 0.1
    assert np.abs(qs.mean(0) - q0[0].cpu().numpy()).max() < 0.2 * max(1.0, np.abs(qs).max())


@pytest.mark.gpu
def test_epsilon_greedy_rule():
    import torch
    pol, _ = _policy(3)
    B = 200000
    greedy = torch.full((B,), 7, dtype=torch.int32, device="cuda")
    a, r = pol.epsilon_greedy(greedy, 0.0, seed=1, counter=0)
    assert (a == 7).all() and (r == 0).all()
    a, r = pol.epsilon_greedy(greedy, 1.0, seed=1, counter=0)
    assert (r == 1).all()
    counts = torch.bincount(a.long(), minlength=21).cpu().numpy()
    assert counts.min() > 0 and np.abs(counts / B - 1 / 21).max() < 5 * np.sqrt((1 / 21) / B)       # uniform over the 21 levels (Q/main_parallel.py:156)
    a, r = pol.epsilon_greedy(greedy, 0.25, seed=1, counter=3)
    frac = r.float().mean().item()
    assert abs(frac - 0.25) < 5 * np.sqrt(0.25 * 0.75 / B)
    assert ((a == 7) | (r == 1)).all()
    a2, r2 = pol.epsilon_greedy(greedy[500:].contiguous(), 0.25, seed=1, traj_offset=500, counter=3)
    assert torch.equal(a[500:], a2) and torch.equal(r[500:], r2)


@pytest.mark.gpu
def test_observation_and_forces():
    import torch
    rng = np.random.Generator(np.random.PCG64(9))
    m = rng.standard_normal((333, 20)) * 10.0 ** rng.integers(-3, 4, (333, 20))
    obs = R.observation(torch.as_tensor(m, device="cuda"), 0.37)
    assert np.array_equal(obs.cpu().numpy(), RO.observation(m, 0.37))
    a = torch.arange(21, dtype=torch.int32, device="cuda")
    f = R.action_forces(a, 21, 5.0).cpu().numpy()
    assert np.allclose(f, GOLD["grid_force_of_action"], rtol=0, atol=1e-15)


@pytest.mark.gpu
def test_experience_rows_are_the_reference_rows_in_trajectory_order_with_wraparound():
    import torch
    K, B, cap = 20, 777, 2000
    ring = R.ReplayRing(K, cap)
    rng = np.random.Generator(np.random.PCG64(4))
    expected = []
    for step in range(5):
        last, cur = rng.standard_normal((B, K)).astype(np.float32), rng.standard_normal((B, K)).astype(np.float32)
        act = rng.integers(0, 21, B).astype(np.int32)
        aux = rng.standard_normal((B, 4)) * 3
        keep = (rng.uniform(size=B) < 0.8).astype(np.uint8) if step != 2 else None
        t = lambda v: torch.as_tensor(v, device="cuda")
        ring.push(t(last), t(cur), t(act), t(aux)[:, 0], reward_scale=-1.5, keep=None if keep is None else t(keep))
        for b in range(B):
            if keep is None or keep[b]:
                expected.append(RO.experience_row(last[b], cur[b], act[b], -aux[b, 0] * 1.5))
    total = ring.total()
    assert total == len(expected) and total > cap
    got = ring.read(total - cap, cap)                              # the last `cap` rows, oldest first (wraps around the ring)
    assert np.array_equal(got, np.array(expected[-cap:], np.float32))


@pytest.mark.gpu
def test_measurement_record_matches_the_reference_list_bookkeeping():
    import torch
    B, RL, cg, CL, scale = 3, 40, 2, 5, 0.7
    rec = R.MeasurementRecord(B, RL, cg, CL)
    lists = [RO.MeasurementLists(RL, cg, CL, scale) for _ in range(B)]
    rng = np.random.Generator(np.random.PCG64(8))
    for step in range(2 * (RL // CL) + 3):                          # long enough for both rings to wrap more than once
        q = rng.standard_normal((B, CL * cg))
        force = rng.integers(-10, 11, B) * 0.5
        rec.push(torch.as_tensor(q, device="cuda"), torch.as_tensor(force, device="cuda"), scale)
        exp_gpu, win_gpu = rec.experience().cpu().numpy(), rec.window().cpu().numpy()
        for b in range(B):
            for s in range(CL * cg):
                lists[b].substep(q[b, s], force[b])
            e, w = lists[b].control_step(force[b])
            assert np.array_equal(exp_gpu[b], e), (step, b)
            assert np.array_equal(win_gpu[b], w), (step, b)
    mask = torch.as_tensor(np.array([0, 1, 0], np.uint8), device="cuda")
    rec.reset(mask)
    w = rec.window().cpu().numpy()
    assert (w[1] == 0).all() and (w[0] != 0).any()


@pytest.mark.gpu
def test_measurement_record_consumes_the_q_stream_of_the_sse_kernel():
    """q[B, n_sub] written by qc_step feeds the record directly (harmonic oscillator, reference sizes 4320 / 1 / 80)."""
    import torch
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import BatchedSim
    p = configs.harmonic()
    B = 8
    sim = BatchedSim(p, batch=B, seed=3)
    sim.init_fock(None)
    rec = R.MeasurementRecord.for_params(p, B)
    assert (rec.read_length, rec.coarse_grain, rec.control_len) == (4320, 1, 80)
    lists = [RO.MeasurementLists(4320, 1, 80, 1.0) for _ in range(B)]
    out = sim.alloc_outputs(want_q=True)
    rng = np.random.Generator(np.random.PCG64(0))
    for step in range(3):
        act = torch.as_tensor(rng.integers(0, 21, B).astype(np.int32), device="cuda")
        sim.step(act, out=out, want_q=True)
        force = R.action_forces(act, 21, p["f_max"])
        rec.push(out["q"], force, 1.0)
        q, f = out["q"].cpu().numpy(), force.cpu().numpy()
        win = rec.window().cpu().numpy()
        for b in range(B):
            for s in range(80):
                lists[b].substep(q[b, s], f[b])
            _, w = lists[b].control_step(f[b])
            assert np.array_equal(win[b], w)


@pytest.mark.gpu
def test_device_actor_closed_loop_writes_consistent_rows():
    """obs -> policy -> epsilon-greedy -> SSE step -> experience row, all on the device; rows chain (row t's `data` is row t+1's `last_data`)."""
    import torch
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import QuantumCartpoleEnv
    B = 64
    env = QuantumCartpoleEnv("inverted_quartic", batch=B, seed=5)      # no warm-up phase: fast reset
    pol, sd = _policy(21, n_in=env.K)
    ring = R.ReplayRing(env.K, 4096)
    actor = R.DeviceActor(env, pol, ring, eps=(0.5, 0.5, 1.0), noise="philox", seed=9)
    obs0 = actor.reset().cpu().numpy()
    steps, acts, rewards, alive = 4, [], [], np.ones(B, bool)
    per_step_rows = []
    for t in range(steps):
        a, r, done, info = actor.step()
        keep = alive.copy()                         # inverted quartic: the failing transition is stored too (inverted quartic main_parallel.py:203-213)
        per_step_rows.append(np.flatnonzero(keep))
        acts.append(a.cpu().numpy()); rewards.append(r.cpu().numpy())
        alive &= ~done.cpu().numpy()
    total = ring.total()
    assert total == sum(len(k) for k in per_step_rows) and total > 0
    rows = ring.read(0, total)
    K = env.K
    off = 0
    prev = {b: obs0[b] for b in range(B)}
    for t in range(steps):
        for b in per_step_rows[t]:
            row = rows[off]; off += 1
            assert np.array_equal(row[:K], prev[b])
            assert row[2 * K] == acts[t][b] and row[2 * K + 1] == rewards[t][b]
            prev[b] = row[K:2 * K]
    assert pol.launch_count() >= 4 * steps
