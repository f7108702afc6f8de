"""CPU restatement (NumPy, float32) of the reference code either side of the SSE step.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and tests/golden/make_golden.py); the product package never imports it.
Pinned by tests/golden/policy_reference_python.npz, which holds outputs of the reference's own `direct_DQN` / `layers.py` classes run
in the authoring container (tests/golden/make_golden.py:make_policy_fixture), and by literal list emulation for the record format.

Reference paths are relative to /root/reference/implementation codes/.
"""
import math

import numpy as np

H1, H2, H3, HV = 512, 512, 256, 128          # quartic oscillator/RL.py:87-98


def policy_state_dict(seed, n_in=20, n_actions=21, noisy_layers=2):
    """A deterministic synthetic state_dict with the reference module's keys and shapes (RL.py:87-98, layers.py:16-19,100).
    Scales follow the reference's initialisers (kaiming-uniform-like weights, sigma = 0.5/sqrt(fan_in), layers.py:23-31); weight_norm is
    deliberately NOT the initial ||W|| so that the folding is exercised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}

    def plain(name, fan_out, fan_in):
        bound = 1.0 / math.sqrt(fan_in)
        w = rng.uniform(-bound, bound, (fan_out, fan_in)).astype(np.float32)
        sd[name + ".weight"] = w
        sd[name + ".bias"] = rng.uniform(-bound, bound, fan_out).astype(np.float32)
        sd[name + ".weight_norm"] = np.float32(np.linalg.norm(w) * rng.uniform(0.7, 1.4))

    def noisy(name, fan_out, fan_in):
        bound = math.sqrt(6.0 / fan_in)
        sd[name + ".u_w"] = rng.uniform(-bound, bound, (fan_out, fan_in)).astype(np.float32)
        sd[name + ".sigma_w"] = (0.5 / math.sqrt(fan_in) * rng.uniform(0.5, 1.5, (fan_out, fan_in))).astype(np.float32)
        sd[name + ".u_b"] = rng.uniform(-0.1, 0.1, fan_out).astype(np.float32)
        sd[name + ".sigma_b"] = (0.5 / math.sqrt(fan_in) * rng.uniform(0.5, 1.5, fan_out)).astype(np.float32)

    plain("fc1", H1, n_in)
    plain("fc2", H2, H1)
    (noisy if noisy_layers >= 2 else plain)("fc31", H3, H2)
    plain("fc32", HV, H2)
    (noisy if noisy_layers >= 1 else plain)("fc41", n_actions, H3)
    plain("fc42", 1, HV)
    return sd


def noisy_f(x):
    """layers.py:82-83."""
    return np.sign(x) * np.sqrt(np.abs(x))


def _linear_wn(sd, name, x):
    """Linear_weight_normalize.forward (layers.py:101-103)."""
    w = sd[name + ".weight"].astype(np.float32)
    w = w / np.float32(np.linalg.norm(w)) * np.float32(sd[name + ".weight_norm"])
    return x @ w.T + sd[name + ".bias"]


def _noisy(sd, name, x, noise):
    """FactorizedNoisy.forward, per-sample branch (layers.py:42-57): w_b = u_w + sigma_w * (rand_out_b rand_in_b^T), out_b = b_b + x_b w_b^T.
    noise = (rand_in [B, in], rand_out [B, out]) or None for `noisy = False` (layers.py:38-40)."""
    u_w, u_b = sd[name + ".u_w"], sd[name + ".u_b"]
    if noise is None:
        return x @ u_w.T + u_b
    rand_in, rand_out = (np.asarray(t, np.float32) for t in noise)
    out = np.empty((x.shape[0], u_w.shape[0]), np.float32)
    for b in range(x.shape[0]):
        eps_w = np.outer(rand_out[b], rand_in[b]).astype(np.float32)
        w = u_w + sd[name + ".sigma_w"] * eps_w
        bias = u_b + sd[name + ".sigma_b"] * rand_out[b]
        out[b] = bias + x[b] @ w.T
    return out


def direct_dqn_forward(sd, x, noise31=None, noise41=None, noisy_layers=2):
    """direct_DQN.forward (RL.py:99-105) -> (action values [B, A], mean prediction [B]) in float32."""
    relu = lambda t: np.maximum(t, np.float32(0))
    x = np.asarray(x, np.float32)
    x = relu(_linear_wn(sd, "fc1", x))
    x = relu(_linear_wn(sd, "fc2", x))
    a = _noisy(sd, "fc31", x, noise31) if noisy_layers >= 2 else _linear_wn(sd, "fc31", x)
    a = relu(a)
    action = _noisy(sd, "fc41", a, noise41) if noisy_layers >= 1 else _linear_wn(sd, "fc41", a)
    mean = relu(_linear_wn(sd, "fc32", x))
    mean = _linear_wn(sd, "fc42", mean)[:, 0]
    return action.astype(np.float32), mean.astype(np.float32)


def observation(moments, input_scaling):
    """`get_data(state)*args.input_scaling` with get_data returning float32 (quartic main_parallel.py:128-131,210)."""
    return np.asarray(moments, np.float64).astype(np.float32) * np.float32(input_scaling)


def experience_row(last_data, data, last_action, reward):
    """quartic main_parallel.py:212-215: np.hstack((last_data, data, [last_action], [reward])) in float32."""
    return np.hstack((np.asarray(last_data, np.float32), np.asarray(data, np.float32), np.array([last_action], dtype=np.float32),
                      np.array([reward], dtype=np.float32)))


def epsilon_threshold(steps_done, eps_start, eps_end, eps_decay):
    """quartic main_parallel.py:152-153."""
    return (eps_start - eps_end) * math.exp(-1. * steps_done / eps_decay) + eps_end


def convert_to_force(n, f_max, oneside=10):
    """direct_DQN.convert_to_force (RL.py:108-112)."""
    return round(n - oneside) * (f_max / oneside)


class MeasurementLists:
    """The list bookkeeping of one actor under `--input measurements` (harmonic oscillator/main_parallel.py:259-292), kept literal."""

    def __init__(self, read_length, coarse_grain, read_control_step_length, input_scaling):
        self.read_length, self.coarse_grain, self.rcsl, self.input_scaling = read_length, coarse_grain, read_control_step_length, input_scaling
        self.measurements_cache = []
        self.measurements_input = list(np.zeros(read_length))                                  # :260
        self.forces_along_measurements_input = list(np.zeros(read_length))                     # :261
        self.forces_to_store = list(np.zeros(read_length // read_control_step_length))         # :261

    def substep(self, q, force):
        """:284-290"""
        self.measurements_cache.append(q)
        if len(self.measurements_cache) == self.coarse_grain:
            self.measurements_input.append(sum(self.measurements_cache) / self.coarse_grain * self.input_scaling)
            self.measurements_cache.clear()
            self.forces_along_measurements_input.append(force * self.input_scaling)

    def control_step(self, force):
        """:268-283 -> (measurement part of the experience row, network input [2, read_length])."""
        self.forces_to_store.append(force * self.input_scaling)
        experience = np.hstack((np.array(self.measurements_input, dtype=np.float32)[::-1], np.array(self.forces_to_store, dtype=np.float32)[::-1]))
        self.measurements_input = self.measurements_input[self.rcsl:]
        self.forces_along_measurements_input = self.forces_along_measurements_input[self.rcsl:]
        self.forces_to_store = self.forces_to_store[1:]
        window = np.array([self.measurements_input[::-1], self.forces_along_measurements_input[::-1]], dtype=np.float32)
        return experience, window
