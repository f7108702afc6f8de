"""Drop-in mirror of the reference's compiled `simulation` CPython module, backed by the CUDA library.

The reference bakes its physics in at compile time (`setupC.py -D...`) and imports the result by name
(`simulation = __import__('simulation')`, quartic main_parallel.py:120).  Here `configure(**params)` plays the role of
setupC.py, and the module-level functions keep the reference's names, argument meaning and error behaviour:

    step(state, dt, F, gamma) -> (q, x_mean, Fail)        quartic simulation_quart.cpp:493-525 (state mutated in place)
    simulate_10_steps(state, dt, F, gamma)                 :526-558
    get_moments(state, out)                                :363-388
    x_expectation(state)                                   :244-258
    set_seed(seed)                                         :645-650
    check_settings()                                       :652-654 / harmonic simulation.cpp:562-564
    Hamiltonian_dot_psi(state) -> 0.0                      harmonic simulation.cpp:566-582 (state <- H state, in place)
    solve_ab(state) -> 0.0                                 harmonic simulation.cpp:584-597 (state <- A^-1 state with the force of the last step)

Every call runs on the GPU (batch of one); there is no CPU path.
"""
import numpy as np

from .sim import BatchedSim

_params = None
_device = 0
_sims = {}
_seed = 0
_last = None        # (dt, gamma, F) of the most recent step: the factorisation the reference's solve_ab would use


def configure(params, device=0):
    """Select the system (a dict from configs.py).  Equivalent of compiling the reference module with setupC.py."""
    global _params, _device, _sims, _last
    _params, _device, _sims, _last = dict(params), device, {}, None


def _sim(dt=None, gamma=None):
    if _params is None:
        raise RuntimeError("simulation.configure(params) has not been called")
    dt = _params["dt"] if dt is None else float(dt)
    gamma = _params["gamma"] if gamma is None else float(gamma)
    key = (dt, gamma)
    if key not in _sims:
        p = dict(_params)
        p["dt"], p["gamma"] = dt, gamma
        s = BatchedSim(p, batch=None, device=_device, seed=_seed)
        _sims[key] = s
    return _sims[key]


def _check_state(state, n):
    # check_type, quartic simulation_quart.cpp:288-305
    if not isinstance(state, np.ndarray):
        raise TypeError("The input object cannot be identified as a Numpy array")
    if state.ndim != 1:
        raise ValueError("The state array is not one-dimensional")
    if state.shape[0] != n:
        raise ValueError("The state array does not match the required size %d" % n)
    if state.dtype != np.complex128:
        raise ValueError("The state array does not match the required datatype: Complex128")
    if not state.flags.c_contiguous:
        raise ValueError("The state array must be contiguous")


def step(state, dt, F, gamma, normals=None):
    """One SSE substep in place.  `normals` (extension): the two N(0,1) draws, else the seeded Philox stream."""
    if not all(isinstance(v, (int, float, np.floating)) for v in (dt, F, gamma)):
        raise TypeError("The input does not match the required input signature (state (numpy array), dt (double), F (double), \\gamma (double))")
    global _last
    s = _sim(dt, gamma)
    _check_state(state, s.n)
    _last = (float(dt), float(gamma), float(F))
    return s.step1(state, float(dt), float(F), float(gamma), normals)


def simulate_10_steps(state, dt, F, gamma, normals=None):
    global _last
    s = _sim(dt, gamma)
    _check_state(state, s.n)
    _last = (float(dt), float(gamma), float(F))
    return s.simulate_10_steps1(state, float(dt), float(F), float(gamma), normals)


def get_moments(state, out):
    s = _sim()
    _check_state(state, s.n)
    # check_moment_data_array, quartic simulation_quart.cpp:306-323
    if not isinstance(out, np.ndarray):
        raise TypeError("The input moment data array (argument 2) cannot be identified as a Numpy array")
    if out.ndim != 1:
        raise ValueError("The moment data array is not one-dimensional")
    if out.shape[0] != s.K:
        raise ValueError("The moment data array does not match the required size %d" % s.K)
    if out.dtype != np.float64:
        raise ValueError("The moment data array does not match the required datatype: Float64")
    s.get_moments1(state, out)
    return None


def x_expectation(state):
    s = _sim()
    _check_state(state, s.n)
    return s.x_expectation1(state)


def Hamiltonian_dot_psi(state):
    s = _sim()
    _check_state(state, s.n)
    s.hamiltonian_dot_psi1(state)
    return 0.0


def solve_ab(state):
    """Solves with the factorisation of the most recent step()'s (dt, gamma, F), like the reference's cached LU; before any step the
    reference's LU is uninitialised -- here that is a RuntimeError."""
    if _last is None:
        raise RuntimeError("solve_ab before the first step(): no factorisation has been set")
    s = _sim(_last[0], _last[1])
    _check_state(state, s.n)
    s.solve_ab1(state, _last[2])
    return 0.0


def set_seed(seed):
    global _seed
    _seed = int(seed)
    for s in _sims.values():
        s.set_seed(_seed)


def check_settings():
    return _sim().check_settings()
