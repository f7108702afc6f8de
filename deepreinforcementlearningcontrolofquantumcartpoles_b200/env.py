"""Batched reset/step environment API with the reference's observation, reward and termination rules.

The reference has no env class: its episode protocol lives in `Control()` of each task's main_parallel.py
(quartic :119-249, inverted quartic :123-239, harmonic :172-312, inverted harmonic likewise).  This class restates that
protocol over a batch axis; the physics is entirely the CUDA kernel behind BatchedSim.step.
"""
import numpy as np

from . import _lib as L
from . import configs
from .sim import BatchedSim


class QuantumCartpoleEnv:
    """reset()/step() over a batch of trajectories.

    auto_reset=True reproduces the actor loop of the reference, where a finished episode is followed at once by a new one
    (quartic main_parallel.py:236-249): trajectories that finish in step() restart inside the same call from a fresh initial state, the
    observation returned for them is the first observation of the new episode and info["terminal_observation"] / info["finished"] carry
    the old one.  Fresh initial states of the quartic task (rejection sampling with a 15-20 time-unit warm-up, :177-198) come from a
    device-resident pool that a second simulator handle refills in bulk, so that a single finished trajectory never stalls the batch."""

    def __init__(self, task="quartic", batch=1024, device=0, seed=0, traj_offset=0, train=True, auto_reset=False, **overrides):
        import torch
        self.torch = torch
        self.task = task
        self.params = configs.PRESETS[task](**overrides)
        self.sim = BatchedSim(self.params, batch=batch, device=device, seed=seed, traj_offset=traj_offset)
        self.B, self.K = batch, self.sim.K
        self.device = device
        self.dev = "cuda:%d" % device
        self.seed, self.traj_offset = int(seed), int(traj_offset)
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed((self.seed * 1000003 + 7919 * (self.traj_offset + 1)) & 0x7FFFFFFFFFFF)
        self.n_levels = self.params.get("n_levels", 21)
        self.zero_action = (self.n_levels - 1) // 2
        self.input_scaling = float(self.params.get("input_scaling", 1.0))
        self.train = train
        self.auto_reset = bool(auto_reset)
        self.t = torch.zeros(batch, dtype=torch.float64, device=self.dev)
        self.done = torch.zeros(batch, dtype=torch.bool, device=self.dev)
        self.fresh = torch.zeros(batch, dtype=torch.bool, device=self.dev)     # restarted in the last step(): first interval runs with F = 0
        self.episodes = torch.zeros(batch, dtype=torch.int64, device=self.dev)
        self.reset_attempts = 0                 # rejection-loop rounds of the last pool fill (quartic)
        self._warm = None                       # second handle: evolves initial-state candidates without touching the live batch
        self.pool = self.pool_obs = None
        self.pool_head = torch.zeros((), dtype=torch.int64, device=self.dev)   # next unused pool slot (device scalar: no host sync in step())
        self.pool_tail = 0                      # states produced so far (host)
        self._steps_since_check = 0

    # -- helpers ------------------------------------------------------------------------------------------------
    def _actions(self, value):
        return self.torch.full((self.B,), int(value), dtype=self.torch.int32, device=self.dev)

    def _obs(self, out):
        # get_data_xp returns float32, then `* args.input_scaling` in float32 (quartic main_parallel.py:128-131,210)
        return out["moments"].to(self.torch.float32) * self.input_scaling

    def observation_size(self):
        return self.K

    def action_to_force(self, action):
        """convert_to_force (quartic RL.py:108-112)."""
        half = (self.n_levels - 1) // 2
        return (action.to(self.torch.float64) - half) * (self.params["f_max"] / half)

    # -- initial states -------------------------------------------------------------------------------------------
    def _warm_sim(self):
        if self._warm is None:
            self._warm = BatchedSim(self.params, batch=self.B, device=self.device, seed=self.seed ^ 0x5DEECE66D, traj_offset=self.traj_offset)
        return self._warm

    def draw_initial_states(self, sim=None, max_rounds=10000):
        """B accepted initial states of the quartic task as a CUDA tensor [B, N], drawn like init_state()/do_episode() do (quartic
        main_parallel.py:177-198): Gaussian packet with wavenumber k ~ U(-0.3, 0.3), free (F = 0) SSE evolution for U(15, 20) time units,
        retried -- without an upper bound, like the reference -- while <H> >= init_energy_cutoff or the boundary test fails on the final state.
        Everything stays on the device: the candidates evolve with per-trajectory substep budgets (accepted ones get budget 0), and
        qc_reset_accept moves accepted states into the result and counts the rest; one 4-byte read per round decides whether to go on."""
        torch, p = self.torch, self.params
        sim = sim or self.sim
        B = self.B
        store = torch.zeros((B, sim.n), dtype=torch.complex128, device=self.dev)
        pending = torch.ones(B, dtype=torch.uint8, device=self.dev)
        n_pending = torch.zeros(1, dtype=torch.int32, device=self.dev)
        act = self._actions(self.zero_action)
        max_steps = int(np.ceil(20.0 / p["dt"])) + 1
        chunk = 1440
        rounds = 0
        while True:
            rounds += 1
            k = torch.rand(B, generator=self.gen, device=self.dev, dtype=torch.float64) * 0.6 - 0.3
            init_time = torch.rand(B, generator=self.gen, device=self.dev, dtype=torch.float64) * 5.0 + 15.0
            # `while t < init_time: step; t += time_step` (:183-187) = ceil(init_time / dt) substeps
            remaining = torch.ceil(init_time / p["dt"]).to(torch.int32) * pending.to(torch.int32)
            sim.init_packets(wavenumber=k, mean=None, std=1.0)
            out = None
            for _ in range((max_steps + chunk - 1) // chunk):
                budget = torch.clamp(remaining, 0, chunk).contiguous()
                out = sim.step(act, n_sub=chunk, nsub_traj=budget, out=out)
                remaining = remaining - budget
            n_pending.zero_()
            sim.reset_accept(out["aux"], p["init_energy_cutoff"], pending, store, n_pending)
            if int(n_pending.item()) == 0:
                break
            if rounds >= max_rounds:
                raise RuntimeError("initial-state rejection loop did not finish in %d rounds (%d trajectories pending): check init_energy_cutoff / grid" % (rounds, int(n_pending.item())))
        self.reset_attempts = rounds
        return store

    def _fill_pool(self, first):
        """(Re)fill the pool of fresh initial states and their first observations."""
        torch = self.torch
        if self.task == "quartic":
            P = 2 * self.B
            if first:
                self.pool = torch.zeros((P, self.sim.n), dtype=torch.complex128, device=self.dev)
                self.pool_obs = torch.zeros((P, self.K), dtype=torch.float32, device=self.dev)
                self.pool_tail = 0
                self.pool_head.zero_()
            warm = self._warm_sim()
            states = self.draw_initial_states(warm)
            warm.set_state(states)
            obs = self._obs(warm.get_moments())
            lo = self.pool_tail % P                       # B divides P: a refill never wraps inside one block
            self.pool[lo:lo + self.B] = states
            self.pool_obs[lo:lo + self.B] = obs
            self.pool_tail += self.B
        elif first:
            if self.task == "inverted_quartic":
                self.sim.init_packets(wavenumber=None, mean=None, std=1.0)        # inverted quartic main_parallel.py:182-183
            else:
                self.sim.init_fock(None)                                          # harmonic main_parallel.py:226-227
            self.pool = self.sim.get_state(numpy=False)[:1].clone().contiguous()   # one deterministic initial state
            self.pool_obs = self._obs(self.sim.get_moments())[:1].clone()
            self.pool_tail = 1

    # -- reset --------------------------------------------------------------------------------------------------
    def reset(self):
        torch, p, sim = self.torch, self.params, self.sim
        self.t.zero_()
        self.done.zero_()
        self.fresh.zero_()
        self.episodes.zero_()
        self._fill_pool(first=True)
        if self.task == "quartic":
            sim.set_state(self.pool[:self.B].contiguous())
            self.pool_head.fill_(self.B)
            if self.auto_reset:
                self._fill_pool(first=False)           # a full batch of spares
            return self.pool_obs[:self.B].clone()      # control starts at i = 0 (quartic main_parallel.py:208)
        sim.set_state(self.pool.expand(self.B, -1).contiguous())
        # the first control interval runs with F = 0: control is skipped at i = 0 (harmonic :238, inverted quartic :201)
        out = sim.step(self._actions(self.zero_action))
        self.t += p["n_sub"] * p["dt"]
        self._update_done(out)
        return self._obs(out)

    def evolve_free(self, steps, chunk=1440):
        """Advance trajectory b by steps[b] substeps with zero force (variable-length warm-up)."""
        torch = self.torch
        remaining = torch.as_tensor(np.asarray(steps, np.int32), device=self.dev)
        act = self._actions(self.zero_action)
        out = None
        for _ in range(max(1, (int(np.max(steps)) + chunk - 1) // chunk)):
            budget = torch.clamp(remaining, 0, chunk).to(torch.int32).contiguous()
            out = self.sim.step(act, n_sub=chunk, nsub_traj=budget, out=out)
            remaining = remaining - budget
        return out

    # -- step ---------------------------------------------------------------------------------------------------
    def _update_done(self, out):
        torch, p = self.torch, self.params
        flags, aux = out["flags"], out["aux"]
        failed = (flags & L.QC_FLAG_FAIL) != 0
        if self.task == "quartic":
            cutoff = p["energy_cutoff"] if self.train else p.get("test_energy_cutoff", 100.0)
            bad = (aux[:, L.QC_AUX_ENERGY] >= cutoff) | failed                      # quartic main_parallel.py:211,216-217
            timeout = self.t >= p["t_max"] - 0.01 * p["dt"]
        elif self.task == "harmonic":
            bad = (aux[:, L.QC_AUX_ENERGY] > p["phonon_cutoff"]) | failed           # harmonic main_parallel.py:239-246
            timeout = self.t >= p["t_max"]
        elif self.task == "inverted_harmonic":
            bad = (aux[:, L.QC_AUX_XMEAN].abs() > p["f_max"]) | failed              # inverted harmonic main_parallel.py:246
            timeout = torch.zeros_like(bad)
        else:
            bad = ((flags & L.QC_FLAG_ESCAPED) != 0) | failed                        # inverted quartic main_parallel.py:199-213
            timeout = torch.zeros_like(bad)
        self.last_bad, self.last_failed = bad, failed
        self.done |= bad | timeout
        return bad, timeout

    def step(self, action):
        """action: int tensor [B] in [0, n_levels).  Returns (obs float32 [B,K], reward float32 [B], done bool [B], info).
        With auto_reset, `done` marks the trajectories that finished in THIS step; they have already been restarted (see class docstring)."""
        torch, p = self.torch, self.params
        a = action.to(device=self.dev, dtype=torch.int32).contiguous()
        if self.auto_reset and self.task != "quartic":
            a = torch.where(self.fresh, torch.full_like(a, self.zero_action), a)        # no control in the first interval of an episode (:201 / :238)
        out = self.sim.step(a)
        self.t += p["n_sub"] * p["dt"]
        bad, timeout = self._update_done(out)
        aux = out["aux"]
        if self.task == "quartic":
            reward = -aux[:, L.QC_AUX_ENERGY] * p["reward_scale"]                    # quartic main_parallel.py:215
        elif self.task == "harmonic":
            reward = -aux[:, L.QC_AUX_ENERGY] * p["reward_scale"]                    # harmonic main_parallel.py:245
        else:
            reward = torch.where(bad, -torch.ones_like(aux[:, 0]), torch.ones_like(aux[:, 0]))   # :207,212
        obs = self._obs(out)
        info = {"energy": aux[:, L.QC_AUX_ENERGY], "x_mean": aux[:, L.QC_AUX_XMEAN], "outside": aux[:, L.QC_AUX_OUTSIDE],
                "numerical_failure": self.last_failed, "timeout": timeout, "t": self.t.clone(), "flags": out["flags"], "was_fresh": self.fresh.clone(), "applied_action": a}
        done = self.done.clone()
        if self.auto_reset:
            fin = done
            rank = torch.cumsum(fin.to(torch.int64), 0) - 1
            slot = self.pool_head + rank
            P = self.pool.shape[0]
            self.sim.reset_scatter(fin.to(torch.uint8).contiguous(), slot.contiguous(), self.pool)
            info["terminal_observation"], info["finished"] = obs, fin
            obs = torch.where(fin[:, None], self.pool_obs[torch.remainder(slot, P)], obs)
            if P > 1:
                self.pool_head += fin.sum()
            self.t = torch.where(fin, torch.zeros_like(self.t), self.t)
            self.episodes += fin.to(torch.int64)
            self.fresh = fin if self.task != "quartic" else torch.zeros_like(fin)
            self.done = torch.zeros_like(self.done)
            self._steps_since_check += 1
            if self.task == "quartic" and self._steps_since_check >= 32:
                # one small read every 32 control steps: refill in bulk while at least one batch of spares is left.  (If more than a whole
                # batch finishes between two checks the ring hands out states twice -- a correlation, not an error.)
                self._steps_since_check = 0
                if self.pool_tail - int(self.pool_head.item()) < self.B:
                    self._fill_pool(first=False)
        return obs, reward.to(torch.float32), done, info
