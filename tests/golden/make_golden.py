#!/usr/bin/env python3
"""Generates the committed golden fixtures under tests/golden/ (run in the authoring container, where /root/reference exists).

Two kinds of vectors:
 (A) outputs of the REFERENCE'S OWN PYTHON code, executed here:
     * `space_def.set_global` (quartic oscillator/space_def.py:5-88) is imported as is -> grid x, Hamiltonian bands, p_hat;
     * the pure functions of the task scripts that define the observation / reward (`cal_energy`, `Gaussian_packet`,
       `calculate_outside_probability`, `adjust_n_max`, `get_data_xp`, `phonon_number`, ...) are extracted from main_parallel.py by
       AST (the scripts themselves run argparse / CUDA set-up at import and cannot be imported) and executed unmodified.
 (B) frozen outputs of the CPU oracle (oracle/sse_oracle.c) on seeded inputs (numpy PCG64), so that any later change of the oracle or
     of the CUDA path is caught on the GPU box, where /root/reference does not exist.
 (C) if oracle/_ref/ was built (the reference .cpp compiled against the MKL-API shim), its step()/get_moments() outputs on the same
     seeded inputs -- the strongest pin of the oracle.
"""
import ast
import warnings
import os
import sys
import types

import numpy as np
from math import pi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/implementation codes"


def extract_functions(path, names, namespace):
    """Compile the named top-level (or nested-in-nothing) function definitions of a reference script into `namespace`."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tree = ast.parse(open(path).read())
    picked = [node for node in tree.body if isinstance(node, ast.FunctionDef) and node.name in names]
    mod = ast.Module(body=picked, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    missing = set(names) - set(n.name for n in picked)
    assert not missing, missing
    return namespace


def grid_reference(task_dir, x_max, x_n, lambda_, mass, xth=None):
    sys.path.insert(0, os.path.join(REF, task_dir))
    for m in ("space_def",):
        sys.modules.pop(m, None)
    import space_def
    g = space_def.set_global(x_max=x_max, x_n_=x_n, lambda_=lambda_, mass=mass)
    sys.path.pop(0)
    ns = {"np": np, "pi": pi, "sqrt": np.sqrt, "x": g["x"], "grid_size": g["grid_size"], "x_n": x_n, "linalg": __import__("scipy.linalg").linalg}
    names = ["probability", "x_expct", "cal_energy", "Gaussian_packet", "p_expct"]
    ns["p_hat"] = g["p_hat"]
    if xth is not None:
        names.append("calculate_outside_probability")
    extract_functions(os.path.join(REF, task_dir, "main_parallel.py"), names, ns)
    return g, ns


def make_grid_fixture(name, task_dir, params):
    n = 2 * int(params["x_max"] / params["grid_size"] + 0.5) + 1
    g, ns = grid_reference(task_dir, params["x_max"], n, params["lambda_"], params["mass"], params.get("x_threshold"))
    H = g["quartic_Hamil"].toarray()
    rng = np.random.default_rng(11)
    states, energies, xs, ps, outs = [], [], [], [], []
    for b in range(4):
        k = rng.uniform(-0.3, 0.3)
        mean = rng.uniform(-2, 2)
        psi = ns["Gaussian_packet"](wavelength=(float("inf") if b == 0 else 1. / k), mean=(0. if b == 0 else mean), std=1.)
        states.append(psi)
        energies.append(ns["cal_energy"](psi, g["Hamil"]))
        xs.append(ns["x_expct"](psi))
        ps.append(ns["p_expct"](psi))
        if params.get("x_threshold"):
            outs.append(ns["calculate_outside_probability"](psi, params["x_threshold"]))
    np.savez_compressed(os.path.join(HERE, name), x=g["x"], grid_size=g["grid_size"],
                        H_bands=np.stack([np.concatenate([np.diag(H, k), np.zeros(k)]) for k in range(5)]),
                        p_hat_dense_untruncated=g["p_hat"].toarray(), states=np.array(states), energy=np.array(energies),
                        x_mean=np.array(xs), p_mean_untruncated=np.array(ps), outside=np.array(outs))
    print("wrote", name)


def make_fock_fixture():
    from common import initial_states
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
    out = {}
    for task, task_dir in (("harmonic", "harmonic oscillator"), ("inverted_harmonic", "inverted harmonic oscillator")):
        params = configs.PRESETS[task]()
        from scipy.sparse import csr_matrix
        from math import sqrt, factorial
        ns = {"np": np, "csr": csr_matrix, "sqrt": sqrt, "factorial": factorial, "pi": pi, "omega": params["omega"], "n_max": params["n_max"],
              "__name__": "reference_fragment", "linalg": __import__("scipy.linalg").linalg}
        names = ["probability", "adjust_n_max", "x_expct", "p_expct", "expct", "get_data_xp"] + (["phonon_number"] if task == "harmonic" else [])
        extract_functions(os.path.join(REF, task_dir, "main_parallel.py"), names, ns)
        ns["adjust_n_max"](params["n_max"])
        psi = initial_states(params, 6, seed=21)
        rng = np.random.default_rng(5)
        psi = psi + 1e-3 * (rng.standard_normal(psi.shape) + 1j * rng.standard_normal(psi.shape)) * np.exp(-np.arange(psi.shape[1]) / 8.0)
        psi /= np.linalg.norm(psi, axis=1, keepdims=True)
        out[task + "_states"] = psi
        out[task + "_obs_float32"] = np.array([ns["get_data_xp"](s) for s in psi])
        if task == "harmonic":
            out[task + "_phonon"] = np.array([ns["phonon_number"](s) for s in psi])
        out[task + "_x"] = np.array([ns["x_expct"](s) for s in psi])
    np.savez_compressed(os.path.join(HERE, "fock_reference_python.npz"), **out)
    print("wrote fock_reference_python.npz")


def make_oracle_fixture():
    from common import TASKS, oracle_for, initial_states, oracle_control_step, level_force
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
    out = {}
    for task in TASKS:
        params = configs.PRESETS[task]()
        B = 3
        rng = np.random.Generator(np.random.PCG64(2024))
        psi0 = initial_states(params, B, seed=4)
        actions = np.array([0, 10, 17], np.int32)
        noise = rng.standard_normal((B, params["n_sub"], 2))
        orc = oracle_for(params)
        ref, fails, qs = oracle_control_step(orc, params, psi0, actions, noise, want_q=True)
        out[task + "_psi0"] = psi0
        out[task + "_actions"] = actions
        out[task + "_noise"] = noise
        out[task + "_psi1"] = ref
        out[task + "_fail"] = fails
        out[task + "_q"] = np.array([q for q, _ in qs])
        out[task + "_xmean"] = np.array([xm for _, xm in qs])
        if "quartic" in task:
            out[task + "_moments"] = np.array([orc.get_moments(ref[b]) for b in range(B)])
    np.savez_compressed(os.path.join(HERE, "oracle_control_step.npz"), **out)
    print("wrote oracle_control_step.npz")


def make_reference_build_fixture():
    """(C) the reference's own simulation*.cpp (compiled against oracle/mkl_shim by oracle/build_ref.sh): one full control step per task."""
    from common import TASKS, initial_states, level_force
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
    from oracle.ref_module import RefModule, available
    out = {}
    for task in TASKS:
        if not available(task):
            print("oracle/_ref/%s missing: run oracle/build_ref.sh first" % task)
            return
        params = configs.PRESETS[task]()
        ref = RefModule(task)
        B = 2
        rng = np.random.Generator(np.random.PCG64(777))
        psi0 = initial_states(params, B, seed=31)
        actions = np.array([4, 18], np.int32)
        noise = rng.standard_normal((B, params["n_sub"], 2))
        psi1 = psi0.copy()
        q = np.zeros((B, params["n_sub"])); xm = np.zeros((B, params["n_sub"])); fail = np.zeros(B, np.int32)
        for b in range(B):
            st = psi1[b].copy()
            for s in range(params["n_sub"]):         # the reference's call pattern: one Python->C call per substep (Q/main_parallel.py:226)
                q[b, s], xm[b, s], f = ref.step(st, params["dt"], level_force(params, int(actions[b])), params["gamma"], noise[b, s])
                fail[b] |= f
            psi1[b] = st
        out[task + "_psi0"], out[task + "_actions"], out[task + "_noise"] = psi0, actions, noise
        out[task + "_psi1"], out[task + "_q"], out[task + "_xmean"], out[task + "_fail"] = psi1, q, xm, fail
        out[task + "_xexp"] = np.array([ref.x_expectation(np.ascontiguousarray(psi1[b])) for b in range(B)])
        out[task + "_settings"] = np.array(ref.check_settings(), dtype=np.float64)
        if "quartic" in task:
            mom = np.zeros((B, 20))
            for b in range(B):
                ref.get_moments(np.ascontiguousarray(psi1[b]), mom[b])
            out[task + "_moments"] = mom
    np.savez_compressed(os.path.join(HERE, "reference_build_control_step.npz"), **out)
    print("wrote reference_build_control_step.npz")


def make_controller_fixture():
    """Outputs of the reference's analytic controllers (quartic oscillator/controllers.py), extracted by AST and run unmodified."""
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
    from common import initial_states
    params = configs.quartic()
    n = 171
    g, ns = grid_reference("quartic oscillator", params["x_max"], n, params["lambda_"], params["mass"])
    extract_functions(os.path.join(REF, "quartic oscillator", "main_parallel.py"), ["x_2_expct", "xpx_expct"], ns)
    ns.update({"x_2": g["x_2"], "lambda_": params["lambda_"], "mass": params["mass"], "control_time": 1.0 / 18, "controls_per_unit_time": 18,
               "sqrt": __import__("math").sqrt})
    extract_functions(os.path.join(REF, "quartic oscillator", "controllers.py"), ["steepest_descent", "LinearQuadratic", "Gaussian_approx"], ns)
    rng = np.random.default_rng(3)
    psi = initial_states(params, 8, seed=13)
    psi = psi * (1 + 0.3 * np.cos(0.7 * g["x"])[None, :])            # make them non-Gaussian (non-zero third moments)
    psi /= np.sqrt(np.sum(np.abs(psi) ** 2, axis=1, keepdims=True) * params["grid_size"])
    out = {"states": psi,
           "damping": np.array([ns["steepest_descent"](s, damping=0.5) for s in psi]),
           "lqg": np.array([ns["LinearQuadratic"](s, k=params["lambda_"] * 2.0) for s in psi]),
           "semiclassical": np.array([ns["Gaussian_approx"](s) for s in psi])}
    np.savez_compressed(os.path.join(HERE, "controllers_reference_python.npz"), **out)
    print("wrote controllers_reference_python.npz")


def extract_classes(path, names, namespace):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tree = ast.parse(open(path).read())
    picked = [node for node in tree.body if isinstance(node, ast.ClassDef) and node.name in names]
    assert len(picked) == len(names), names
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), namespace)
    return namespace


def make_policy_fixture():
    """The reference's own `direct_DQN` (quartic oscillator/RL.py:81-112; RL.py itself imports numba/termcolor and cannot be imported, so
    the class is extracted by AST) on top of its own layers.py (imported as is), evaluated in float32 on the CPU with the synthetic
    state_dict of oracle/rollout_oracle.py:policy_state_dict.  Stores inputs, the noise the reference drew, and its outputs."""
    import importlib.util
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from oracle import rollout_oracle as RO
    spec = importlib.util.spec_from_file_location("layers", os.path.join(REF, "quartic oscillator", "layers.py"))
    layers = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(layers)
    ns = {"nn": nn, "F": F, "torch": torch, "layers": layers, "F_max": 5., "print": lambda *a, **k: None}
    extract_classes(os.path.join(REF, "quartic oscillator", "RL.py"), ["direct_DQN"], ns)
    out = {}
    for tag, n_in, noisy_layers, seed in (("grid", 20, 2, 101), ("fock", 5, 2, 102)):     # noisy_layers < 2 does not run in the reference (RL.py:103 unpacks a tuple)
        torch.manual_seed(seed)
        net = ns["direct_DQN"](n_in, noisy_layers=noisy_layers)
        sd = RO.policy_state_dict(seed, n_in=n_in, noisy_layers=noisy_layers)
        net.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
        net.eval()
        rng = np.random.Generator(np.random.PCG64(seed + 1))
        B = 6                                                        # not a multiple of 128 -> per-sample noise branch (layers.py:41-57)
        x = (rng.standard_normal((B, n_in)) * 2.0).astype(np.float32)
        with torch.no_grad():
            action, mean, _ = net(torch.as_tensor(x))
        out[tag + "_x"] = x
        out[tag + "_action_values"] = action.numpy()
        out[tag + "_mean"] = mean.numpy()
        out[tag + "_argmax"] = action.max(1)[1].numpy()
        if noisy_layers == 2:
            for name in ("fc31", "fc41"):
                layer = getattr(net, name)
                i = layer.randbuffer_pointer - B
                out["%s_%s_rand_in" % (tag, name)] = layer.randbuffer_in[i:i + B, 0, :].numpy()
                out["%s_%s_rand_out" % (tag, name)] = layer.randbuffer_out[i:i + B, :, 0].numpy()
            # (`noisy = False` makes FactorizedNoisy return a bare tensor, which RL.py:103 cannot unpack: the reference always acts with noise on)
        out[tag + "_force_of_action"] = np.array([net.convert_to_force(int(a)) for a in range(21)])
    np.savez_compressed(os.path.join(HERE, "policy_reference_python.npz"), **out)
    print("wrote policy_reference_python.npz")


if __name__ == "__main__":
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
    make_grid_fixture("grid_reference_python_quartic.npz", "quartic oscillator", configs.quartic())
    make_grid_fixture("grid_reference_python_inverted_quartic.npz", "inverted quartic oscillator", configs.inverted_quartic())
    make_fock_fixture()
    make_oracle_fixture()
    make_reference_build_fixture()
    make_controller_fixture()
    make_policy_fixture()
