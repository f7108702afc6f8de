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
    def __init__(self, task="quartic", batch=1024, device=0, seed=0, traj_offset=0, train=True, **overrides):
        import torch
        self.torch = torch
        self.task = task
        self.params = configs.PRESETS[task](**overrides)
        self.sim = BatchedSim(self.params, batch=batch, device=device, seed=seed, traj_offset=traj_offset)
        self.B, self.K = batch, self.sim.K
        self.dev = "cuda:%d" % device
        self.rng = np.random.RandomState(seed + 7919 * (traj_offset + 1))
        self.n_levels = self.params.get("n_levels", 21)
        self.zero_action = (self.n_levels - 1) // 2
        self.input_scaling = float(self.params.get("input_scaling", 1.0))
        self.train = train
        self.t = torch.zeros(batch, dtype=torch.float64, device=self.dev)
        self.done = torch.zeros(batch, dtype=torch.bool, device=self.dev)

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

    # -- reset --------------------------------------------------------------------------------------------------
    def reset(self):
        torch, p, sim = self.torch, self.params, self.sim
        self.t.zero_()
        self.done.zero_()
        if self.task == "quartic":
            # Gaussian packet with k ~ U(-0.3, 0.3), free (F=0) SSE evolution for U(15,20) time units, retry while
            # <H> >= 7.5 or Fail (quartic main_parallel.py:177-198)
            pending = np.ones(self.B, bool)
            state = np.zeros((self.B, sim.n), np.complex128)
            x = sim.x_grid()
            for _ in range(8):
                k = self.rng.uniform(-0.3, 0.3, self.B)
                sim.init_packets(wavenumber=k, mean=None, std=1.0)
                steps = np.ceil(self.rng.uniform(15., 20., self.B) / p["dt"]).astype(np.int32)
                out = self.evolve_free(steps)
                ok = ((out["aux"][:, L.QC_AUX_ENERGY] < p["init_energy_cutoff"]) & ((out["flags"] & L.QC_FLAG_FAIL) == 0)).cpu().numpy()
                new = sim.get_state()
                take = pending & ok
                state[take] = new[take]
                pending &= ~ok
                if not pending.any():
                    break
            if pending.any():     # extremely unlikely; fall back to the last attempt for the remainder
                state[pending] = new[pending]
            sim.set_state(state)
            out = sim.get_moments()
            return self._obs(out)
        if self.task == "inverted_quartic":
            sim.init_packets(wavenumber=None, mean=None, std=1.0)        # inverted quartic main_parallel.py:182-183
        else:
            sim.init_fock(None)                                          # harmonic main_parallel.py:226-227
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
        while True:
            m = int(remaining.max().item())
            if m <= 0 and out is not None:
                break
            ns = max(1, min(chunk, m))
            budget = torch.clamp(remaining, 0, ns).to(torch.int32).contiguous()
            out = self.sim.step(act, n_sub=ns, nsub_traj=budget)
            remaining = remaining - budget
            if m <= 0:
                break
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
        """action: int tensor [B] in [0, n_levels).  Returns (obs float32 [B,K], reward float32 [B], done bool [B], info)."""
        torch, p = self.torch, self.params
        a = action.to(device=self.dev, dtype=torch.int32).contiguous()
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
        info = {"energy": aux[:, L.QC_AUX_ENERGY], "x_mean": aux[:, L.QC_AUX_XMEAN], "outside": aux[:, L.QC_AUX_OUTSIDE],
                "numerical_failure": self.last_failed, "timeout": timeout, "t": self.t.clone(), "flags": out["flags"]}
        return self._obs(out), reward.to(torch.float32), self.done.clone(), info
