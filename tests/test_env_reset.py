"""Episode protocol of the environment mirror: device-side rejection sampling of the quartic initial state,
auto-reset of finished trajectories, and which transitions become experience rows in each task.

Reference: quartic main_parallel.py:177-198 (init_state + unbounded retry), :236-249 (a finished episode is followed by a new one at once);
inverted quartic main_parallel.py:198-221, inverted harmonic :242-258 (failing transition stored with reward -1)."""
import numpy as np
import pytest

from deepreinforcementlearningcontrolofquantumcartpoles_b200 import QuantumCartpoleEnv, BatchedSim, configs, rollout as R, _lib as L

pytestmark = pytest.mark.gpu


def _boundary_ok(psi, thr=5e-3, k=6):
    lo = np.linalg.norm(psi[:, :k], axis=1); hi = np.linalg.norm(psi[:, -k:], axis=1)
    return (lo <= thr) & (hi <= thr)


def test_quartic_reset_is_rejection_sampling_on_the_device():
    """Every returned initial state satisfies the reference's acceptance rule; with a cut-off that rejects most candidates the loop keeps
    going (several rounds) instead of falling back to rejected states; nothing round-trips through NumPy (states are CUDA tensors)."""
    import torch
    env = QuantumCartpoleEnv("quartic", batch=48, seed=3)
    obs = env.reset()
    assert obs.is_cuda and obs.shape == (48, 20) and torch.isfinite(obs).all()
    out = env.sim.get_moments()
    assert float(out["aux"][:, L.QC_AUX_ENERGY].max()) < env.params["init_energy_cutoff"]
    assert _boundary_ok(env.sim.get_state()).all()
    assert env.reset_attempts >= 1
    base_rounds = env.reset_attempts

    # how selective the default rule is on this seed, then a much stricter one: more rounds, same guarantee
    e_default = out["aux"][:, L.QC_AUX_ENERGY].cpu().numpy()
    strict = float(np.percentile(e_default, 35))
    env2 = QuantumCartpoleEnv("quartic", batch=48, seed=3, init_energy_cutoff=strict)
    states = env2.draw_initial_states()
    assert states.is_cuda and states.dtype == torch.complex128
    env2.sim.set_state(states)
    e2 = env2.sim.get_moments()["aux"][:, L.QC_AUX_ENERGY]
    assert float(e2.max()) < strict
    assert env2.reset_attempts > base_rounds and env2.reset_attempts >= 3
    assert _boundary_ok(states.cpu().numpy()).all()
    nrm = (states.abs() ** 2).sum(1) * env2.params["grid_size"]
    assert torch.allclose(nrm, torch.ones_like(nrm), atol=1e-12)


def test_reset_accept_uses_the_final_state_not_the_latched_flag():
    """init_state() returns the Fail of its LAST step (quartic main_parallel.py:183-187): a candidate whose flag latched earlier but whose final
    state passes the boundary test is accepted; one that fails on the final state is not."""
    import torch
    params = configs.quartic()
    sim = BatchedSim(params, batch=3)
    n = sim.n
    x = sim.x_grid()
    good = np.exp(-x ** 2 / 4) / (2 * np.pi) ** 0.25
    bad = good.copy().astype(np.complex128); bad[:6] += 0.01                 # ||psi[0:6]|| > 5e-3
    psi = np.stack([good.astype(np.complex128), bad, good.astype(np.complex128)])
    sim.set_state(psi)
    aux = torch.zeros((3, 4), dtype=torch.float64, device="cuda"); aux[:, L.QC_AUX_ENERGY] = torch.tensor([1.0, 1.0, 9.0], dtype=torch.float64)
    pending = torch.ones(3, dtype=torch.uint8, device="cuda")
    store = torch.zeros((3, n), dtype=torch.complex128, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    sim.reset_accept(aux, 7.5, pending, store, cnt)
    assert pending.cpu().tolist() == [0, 1, 1] and int(cnt.item()) == 2       # boundary failure and energy >= cut-off stay pending
    assert np.array_equal(store[0].cpu().numpy(), psi[0]) and float(store[1:].abs().max()) == 0.0


def test_auto_reset_restarts_finished_trajectories_inside_step():
    import torch
    B = 64
    env = QuantumCartpoleEnv("inverted_quartic", batch=B, seed=11, auto_reset=True)
    obs0 = env.reset()
    init_obs = env.pool_obs[0].cpu().numpy()
    psi_init = env.pool[0].cpu().numpy()
    push = torch.full((B,), 20, dtype=torch.int64, device=env.dev)           # full force in one direction: every trajectory escapes within ~40 steps
    seen = np.zeros(B, bool)
    was_fresh_prev = np.zeros(B, bool)
    for step in range(60):
        obs, reward, done, info = env.step(push)
        d = done.cpu().numpy()
        assert np.array_equal(info["finished"].cpu().numpy(), d)
        if d.any():
            assert np.all(reward.cpu().numpy()[d] == -1.0)
            assert np.array_equal(obs.cpu().numpy()[d], np.tile(init_obs, (d.sum(), 1)))          # first observation of the new episode
            assert not np.array_equal(info["terminal_observation"].cpu().numpy()[d][0], init_obs)
            assert np.all(env.t.cpu().numpy()[d] == 0.0) and not bool(env.done.any())
            got = env.sim.get_state()
            assert np.array_equal(got[d], np.tile(psi_init, (d.sum(), 1)))
            assert np.array_equal(env.fresh.cpu().numpy(), d)
        # trajectories restarted in the previous step run this interval with F = 0 whatever the policy asked for (inverted quartic main_parallel.py:201)
        applied = info["applied_action"].cpu().numpy()
        assert np.all(applied[was_fresh_prev] == env.zero_action) and np.all(applied[~was_fresh_prev] == 20)
        assert np.array_equal(info["was_fresh"].cpu().numpy(), was_fresh_prev)
        was_fresh_prev = d.copy()
        seen |= d
        if seen.all() and step > 45:
            break
    assert seen.all()
    assert int(env.episodes.min()) >= 1


def test_quartic_auto_reset_draws_from_the_pool_and_refills():
    import torch
    B = 32
    env = QuantumCartpoleEnv("quartic", batch=B, seed=5, auto_reset=True, energy_cutoff=3.0)     # low cut-off: episodes end quickly under a hard push
    env.reset()
    assert env.pool.shape[0] == 2 * B and env.pool_tail == 2 * B and int(env.pool_head.item()) == B
    push = torch.full((B,), 20, dtype=torch.int64, device=env.dev)
    finished_total = 0
    for step in range(70):
        obs, reward, done, info = env.step(push)
        finished_total += int(done.sum().item())
    assert finished_total > B                                              # more restarts than the first batch of spares: the pool was refilled
    assert env.pool_tail > 2 * B and int(env.pool_head.item()) == B + finished_total
    e = env.sim.get_moments()["aux"][:, L.QC_AUX_ENERGY]
    assert torch.isfinite(e).all()
    # every state in the pool obeys the acceptance rule
    warm = env._warm_sim(); warm.set_state(env.pool[:B].contiguous())
    assert float(warm.get_moments()["aux"][:, L.QC_AUX_ENERGY].max()) < env.params["init_energy_cutoff"]


@pytest.mark.parametrize("task", ["inverted_quartic", "quartic"])
def test_experience_rows_follow_the_reference_rule_per_task(task):
    """inverted tasks store the failing transition (reward -1); cooling tasks drop it (and the one that reaches t_max)."""
    import torch
    from oracle import rollout_oracle
    B = 48
    kw = dict(energy_cutoff=3.0) if task == "quartic" else {}
    env = QuantumCartpoleEnv(task, batch=B, seed=2, **kw)
    pol = R.DirectDQNPolicy(env.K, 21)
    pol.load_state_dict(rollout_oracle.policy_state_dict(3, n_in=env.K))
    ring = R.ReplayRing(env.K, 8192)
    actor = R.DeviceActor(env, pol, ring, eps=(1.0, 1.0, 1.0), noise="philox", seed=4)       # eps = 1: random actions, episodes end
    actor.reset()
    alive = np.ones(B, bool)
    expect_rows, expect_rewards = 0, []
    for t in range(40):
        a, r, done, info = actor.step()
        bad = env.last_bad.cpu().numpy(); timeout = info["timeout"].cpu().numpy()
        keep = alive & ~bad & ~timeout if task == "quartic" else alive.copy()
        expect_rows += int(keep.sum()); expect_rewards += list(r.cpu().numpy()[keep])
        alive &= ~done.cpu().numpy()
    assert ring.total() == expect_rows
    rows = ring.read(0, expect_rows)
    assert np.array_equal(rows[:, 2 * env.K + 1], np.array(expect_rewards, np.float32))
    if task == "inverted_quartic":
        assert (~alive).any() and (rows[:, 2 * env.K + 1] == -1.0).sum() == int((~alive).sum())     # one failing row per finished trajectory
    else:
        assert (~alive).any() and np.all(rows[:, 2 * env.K + 1] > -3.0)                               # no row at or beyond the energy cut-off
