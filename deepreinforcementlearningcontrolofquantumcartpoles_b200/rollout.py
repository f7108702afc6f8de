"""Device-resident steps either side of the SSE kernel (C-ABI: include/qcart_rollout.h).

  DirectDQNPolicy      the reference's `direct_DQN` (quartic RL.py:81-112, layers.py) evaluated on the whole batch by CUDA kernels
  ReplayRing           experience rows (last_obs, obs, action, reward) of quartic main_parallel.py:212-215 written on the device
  MeasurementRecord    sliding window of coarse-grained measurement outcomes, `--input measurements` (harmonic main_parallel.py:259-292)
  DeviceActor          the actor <-> manager loop (quartic main_parallel.py:150-165,199-233,331-360) for a batch of trajectories without
                       a host round trip per control step
PyTorch only owns the tensors; all arithmetic is in libqcart.so.  There is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L


def _torch():
    import torch
    return torch


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def fold_weight_norm(weight, weight_norm):
    """Effective weight of `Linear_weight_normalize`: weight / ||weight||_F * weight_norm, in float32 (layers.py:97-103)."""
    w = np.asarray(_np(weight), np.float32)
    g = np.float32(_np(weight_norm))
    return (w / np.float32(np.sqrt(np.sum(w.astype(np.float32) ** 2, dtype=np.float32))) * g).astype(np.float32)


def epsilon_threshold(steps_done, eps_start, eps_end, eps_decay):
    """quartic main_parallel.py:152-153."""
    return (eps_start - eps_end) * math.exp(-1.0 * steps_done / eps_decay) + eps_end


class DirectDQNPolicy:
    HIDDEN = (512, 512, 256, 128)

    def __init__(self, n_in, n_actions=21, noisy_layers=2, device=0):
        self.lib = L.load()
        self.n_in, self.n_actions, self.noisy_layers, self.device = int(n_in), int(n_actions), int(noisy_layers), int(device)
        h = C.c_void_p()
        L.check(self.lib.qc_policy_create(self.n_in, self.n_actions, self.noisy_layers, self.device, C.byref(h)))
        self.h = h
        self.noise_width = int(self.lib.qc_policy_noise_width(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.qc_policy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def set_param(self, name, array):
        a = np.ascontiguousarray(_np(array), np.float32).reshape(-1)
        L.check(self.lib.qc_policy_set_param(self.h, L.POLICY_PARAMS.index(name), a.ctypes.data, a.size))

    def load_state_dict(self, sd):
        """`sd`: the reference module's state_dict (torch tensors or numpy arrays), keys as in quartic RL.py:87-98."""
        def plain(prefix, wname, bname):          # Linear_weight_normalize
            self.set_param(wname, fold_weight_norm(sd[prefix + ".weight"], sd[prefix + ".weight_norm"]))
            self.set_param(bname, sd[prefix + ".bias"])

        def noisy(prefix, tag, is_noisy):
            if is_noisy:                           # FactorizedNoisy
                for src, dst in (("u_w", "UW"), ("sigma_w", "SW"), ("u_b", "UB"), ("sigma_b", "SB")):
                    self.set_param("%s_%s" % (tag, dst), sd["%s.%s" % (prefix, src)])
            else:
                plain(prefix, tag + "_UW", tag + "_UB")
        plain("fc1", "FC1_W", "FC1_B")
        plain("fc2", "FC2_W", "FC2_B")
        noisy("fc31", "FC31", self.noisy_layers >= 2)
        noisy("fc41", "FC41", self.noisy_layers >= 1)
        plain("fc32", "FC32_W", "FC32_B")
        plain("fc42", "FC42_W", "FC42_B")

    def pack_noise(self, rand_in31, rand_out31, rand_in41, rand_out41):
        """Per-sample factorised noise -> the [B, noise_width] float32 CUDA tensor qc_policy_forward reads (QC_NOISE_GIVEN)."""
        torch = _torch()
        B = len(rand_in31)
        out = torch.zeros((B, self.noise_width), dtype=torch.float32, device="cuda:%d" % self.device)
        o = 0
        for part in (rand_in31, rand_out31, rand_in41, rand_out41):
            t = torch.as_tensor(np.asarray(_np(part), np.float32).reshape(B, -1), device=out.device)
            out[:, o:o + t.shape[1]] = t
            o += t.shape[1]
        return out

    def forward(self, obs, noise=None, seed=0, traj_offset=0, counter=0, want_q=True, want_value=False, want_greedy=True):
        """obs: float32 CUDA tensor [B, n_in].  noise: None (noisy=False), "philox", or a packed tensor from pack_noise().
        Returns dict(q=[B, n_actions], value=[B], greedy=int32 [B])."""
        torch = _torch()
        assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and obs.shape[1] == self.n_in
        B = obs.shape[0]
        dev = obs.device
        out = {}
        if want_q:
            out["q"] = torch.empty((B, self.n_actions), dtype=torch.float32, device=dev)
        if want_value:
            out["value"] = torch.empty((B,), dtype=torch.float32, device=dev)
        if want_greedy:
            out["greedy"] = torch.empty((B,), dtype=torch.int32, device=dev)
        if noise is None:
            mode, nptr = L.QC_NOISE_OFF, None
        elif isinstance(noise, str):
            assert noise == "philox"
            mode, nptr = L.QC_NOISE_PHILOX, None
        else:
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous() and tuple(noise.shape) == (B, self.noise_width)
            mode, nptr = L.QC_NOISE_GIVEN, noise.data_ptr()
        ptr = lambda k: out[k].data_ptr() if k in out else None
        L.check(self.lib.qc_policy_forward(self.h, obs.data_ptr(), B, mode, nptr, int(seed), int(traj_offset), int(counter),
                                           ptr("q"), ptr("value"), ptr("greedy"), self._stream()))
        return out

    def epsilon_greedy(self, greedy, eps, seed=0, traj_offset=0, counter=0):
        torch = _torch()
        B = greedy.numel()
        action = torch.empty((B,), dtype=torch.int32, device=greedy.device)
        rnd = torch.empty((B,), dtype=torch.uint8, device=greedy.device)
        L.check(self.lib.qc_epsilon_greedy(greedy.data_ptr(), B, self.n_actions, float(eps), int(seed), int(traj_offset), int(counter),
                                           action.data_ptr(), rnd.data_ptr(), self.device, self._stream()))
        return action, rnd

    def launch_count(self):
        return int(self.lib.qc_policy_launch_count(self.h))

    def set_gemm(self, kind):
        """"tcgen05" (default: 3xTF32 on the tensor cores, pre-split operands fed by TMA), "simt" (fp32 FMA on the CUDA cores, the cross-check)
        "tcgen05_staged" (the same tensor-core arithmetic with operands split and staged by the CTA's threads) or "tcgen05_raw" (TMA-fed
        from the plain fp32 matrices, lo parts derived in shared memory)."""
        L.check(self.lib.qc_policy_set_gemm(self.h, {"tcgen05": 0, "simt": 1, "tcgen05_staged": 2, "tcgen05_raw": 3}[kind]))


def observation(moments, input_scaling=1.0):
    """float32(moments) * input_scaling on the device (`get_data(state)*args.input_scaling`, quartic main_parallel.py:128-131,210)."""
    torch = _torch()
    assert moments.is_cuda and moments.dtype == torch.float64 and moments.is_contiguous()
    obs = torch.empty(moments.shape, dtype=torch.float32, device=moments.device)
    stream = C.c_void_p(torch.cuda.current_stream(moments.device).cuda_stream)
    L.check(L.load().qc_obs_f32(moments.data_ptr(), moments.numel(), float(input_scaling), obs.data_ptr(), stream))
    return obs


def action_forces(action, n_levels, f_max):
    """`convert_to_force` (quartic RL.py:108-112) on the device: float64 forces of int32 actions."""
    torch = _torch()
    force = torch.empty(action.shape, dtype=torch.float64, device=action.device)
    stream = C.c_void_p(torch.cuda.current_stream(action.device).cuda_stream)
    L.check(L.load().qc_action_forces(action.data_ptr(), action.numel(), int(n_levels), float(f_max), force.data_ptr(), action.device.index or 0, stream))
    return force


class ReplayRing:
    """Device ring of float32 experience rows [capacity, 2K+2]."""

    def __init__(self, K, capacity, device=0):
        self.lib = L.load()
        self.K, self.row_len, self.capacity, self.device = int(K), 2 * int(K) + 2, int(capacity), int(device)
        h = C.c_void_p()
        L.check(self.lib.qc_replay_create(self.row_len, self.capacity, self.device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.qc_replay_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def push(self, last_obs, obs, last_action, reward_src, reward_scale=1.0, keep=None, reward_stride=None):
        """reward_src: float64 CUDA tensor; element b*reward_stride is trajectory b's reward term (e.g. the aux block's energy column)."""
        torch = _torch()
        B = last_action.numel()
        assert last_obs.dtype == torch.float32 and obs.dtype == torch.float32 and last_obs.is_contiguous() and obs.is_contiguous()
        assert last_action.dtype == torch.int32 and reward_src.dtype == torch.float64
        if reward_stride is None:
            reward_stride = reward_src.stride(0) if reward_src.dim() >= 1 else 1
        if keep is not None:
            assert keep.dtype == torch.uint8 and keep.numel() == B and keep.is_contiguous()
        L.check(self.lib.qc_replay_push(self.h, last_obs.data_ptr(), obs.data_ptr(), self.K, last_action.data_ptr(), reward_src.data_ptr(),
                                        int(reward_stride), float(reward_scale), None if keep is None else keep.data_ptr(), B, self._stream()))

    def total(self):
        t = C.c_int64()
        L.check(self.lib.qc_replay_total(self.h, C.byref(t), self._stream()))
        return t.value

    def read(self, first, count):
        out = np.empty((count, self.row_len), np.float32)
        L.check(self.lib.qc_replay_read(self.h, int(first), int(count), out.ctypes.data, self._stream()))
        return out


class MeasurementRecord:
    """Per-trajectory sliding window of coarse-grained measurement outcomes and applied forces (harmonic main_parallel.py:142-150,259-292)."""

    def __init__(self, batch, read_length, coarse_grain, control_len, device=0):
        self.lib = L.load()
        self.B, self.read_length, self.coarse_grain, self.control_len, self.device = int(batch), int(read_length), int(coarse_grain), int(control_len), int(device)
        h = C.c_void_p()
        L.check(self.lib.qc_record_create(self.B, self.read_length, self.coarse_grain, self.control_len, self.device, C.byref(h)))
        self.h = h
        self.row_len = int(self.lib.qc_record_row_len(self.h))

    @classmethod
    def for_params(cls, params, batch, n_periods_to_read=1.5, num_of_data_per_time_unit=1440, device=0):
        """Sizes as the reference derives them (harmonic main_parallel.py:142-148)."""
        time_steps = int(round(1.0 / params["dt"]))
        assert time_steps % num_of_data_per_time_unit == 0
        coarse = time_steps // num_of_data_per_time_unit
        read_length = int(round(n_periods_to_read * 2 * num_of_data_per_time_unit))
        return cls(batch, read_length, coarse, params["n_sub"] // coarse, device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.qc_record_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def reset(self, mask=None):
        L.check(self.lib.qc_record_reset(self.h, None if mask is None else mask.data_ptr(), self._stream()))

    def push(self, q, force, input_scaling=1.0):
        torch = _torch()
        assert q.is_cuda and q.dtype == torch.float64 and q.is_contiguous() and q.shape[0] == self.B
        assert force.is_cuda and force.dtype == torch.float64 and force.numel() == self.B
        L.check(self.lib.qc_record_push(self.h, q.data_ptr(), q.shape[1], force.data_ptr(), float(input_scaling), self._stream()))

    def window(self):
        torch = _torch()
        out = torch.empty((self.B, 2, self.read_length), dtype=torch.float32, device="cuda:%d" % self.device)
        L.check(self.lib.qc_record_window(self.h, out.data_ptr(), self._stream()))
        return out

    def experience(self):
        torch = _torch()
        out = torch.empty((self.B, self.row_len), dtype=torch.float32, device="cuda:%d" % self.device)
        L.check(self.lib.qc_record_experience(self.h, out.data_ptr(), self._stream()))
        return out


class DeviceActor:
    """The reference's actors + manager for one batch: observation -> policy -> epsilon-greedy -> SSE control step -> experience row, every
    control step, with all tensors resident on the GPU (quartic main_parallel.py:150-165,199-233,331-360).

    eps = (EPS_START, EPS_END, EPS_DECAY) follows :137-142; `steps_done` advances by the number of trajectories per control step (:154)."""

    def __init__(self, env, policy, replay=None, eps=(0.2, 0.004, None), noise="philox", seed=0, train=True):
        self.env, self.policy, self.replay = env, policy, replay
        p = env.params
        decay = eps[2] if eps[2] is not None else (1.0 / (p["n_sub"] * p["dt"])) * p.get("t_max", 100.0) * 80
        self.eps_start, self.eps_end, self.eps_decay = eps[0], eps[1], decay
        self.noise, self.seed, self.train = noise, int(seed), train
        self.steps_done = 0
        self.counter = 0
        self.obs = None

    def reset(self):
        self.obs = self.env.reset().contiguous()
        return self.obs

    def step(self):
        torch = _torch()
        env = self.env
        assert self.obs is not None, "call reset() first"
        alive_before = ~env.done
        pol = self.policy.forward(self.obs, noise=self.noise, seed=self.seed, traj_offset=env.sim.traj_offset, counter=self.counter, want_q=False)
        eps = epsilon_threshold(self.steps_done, self.eps_start, self.eps_end, self.eps_decay) if self.train else 0.0
        self.steps_done += env.B
        action, rnd = self.policy.epsilon_greedy(pol["greedy"], eps, seed=self.seed, traj_offset=env.sim.traj_offset, counter=self.counter)
        self.counter += 1
        obs, reward, done, info = env.step(action)
        obs = obs.contiguous()
        if self.replay is not None and self.train:
            # Which transitions become experience rows depends on the task:
            #   quartic / harmonic: only while the new state is alive -- the failing transition and the one that reaches t_max are dropped
            #     (quartic main_parallel.py:207,211-217; harmonic :237-246);
            #   inverted harmonic / inverted quartic: the failing transition IS stored, with failing_reward = -1 (inverted harmonic :247-258,
            #     inverted quartic :203-213) -- the only negative-reward samples the learner ever sees.
            # A restarted episode's first (uncontrolled) interval is never stored (`i != control_interval`, :204 / :249).
            if env.task in ("quartic", "harmonic"):
                keep = alive_before & ~env.last_bad & ~info["timeout"]
            else:
                keep = alive_before & ~info["was_fresh"]
            if env.auto_reset:                                   # the next observation of a finished trajectory already belongs to the new episode
                nxt = torch.where(info["finished"][:, None], info["terminal_observation"], obs).contiguous()
            else:
                nxt = obs
            self.replay.push(self.obs, nxt, action, reward.to(torch.float64), keep=keep.to(torch.uint8).contiguous())     # float32 -> float64 -> float32 is lossless
        self.obs = obs
        return action, reward, done, info
