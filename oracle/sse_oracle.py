"""ctypes front-end of oracle/sse_oracle.c (TEST INFRASTRUCTURE ONLY).

Mirrors the reference `simulation` module's function set (Q = quartic simulation_quart.cpp:656-668,
H = harmonic simulation.cpp:599-612) with the two Gaussian normals per substep passed in explicitly
(the MKL VSL stream of Q:572,648 is not reproducible without MKL).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class _Cfg(C.Structure):
    _fields_ = [("variant", C.c_int), ("n", C.c_int), ("x_max", C.c_double), ("grid_size", C.c_double),
                ("lambda_", C.c_double), ("mass", C.c_double), ("omega", C.c_double),
                ("moment_order", C.c_int), ("herm_mode", C.c_int)]


def build(force=False):
    """Compile the oracle shared libraries (gcc).  Building the checker is not using it."""
    so = os.path.join(_HERE, "_build", "libsse_oracle.so")
    src = os.path.join(_HERE, "sse_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "all"])
    return so


_libs = {}


def _lib(fast=False):
    if fast not in _libs:
        build()
        path = os.path.join(_HERE, "_build", "libsse_oracle_fast.so" if fast else "libsse_oracle.so")
        try:
            lib = C.CDLL(path)
        except OSError:
            build(force=True)
            lib = C.CDLL(path)
        lib.sse_oracle_create.restype = C.c_void_p
        lib.sse_oracle_create.argtypes = [C.POINTER(_Cfg)]
        lib.sse_oracle_destroy.argtypes = [C.c_void_p]
        lib.sse_oracle_n.argtypes = [C.c_void_p]
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        lib.sse_oracle_step.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double, vp, dp, dp, ip]
        lib.sse_oracle_run.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double, vp, C.c_int, vp, vp, ip]
        lib.sse_oracle_x_expectation.restype = C.c_double
        lib.sse_oracle_x_expectation.argtypes = [vp, vp]
        lib.sse_oracle_get_moments.argtypes = [vp, vp, vp]
        lib.sse_oracle_get_A_dense.argtypes = [vp, C.c_double, C.c_double, vp]
        lib.sse_oracle_get_C_dense.argtypes = [vp, C.c_double, C.c_double, vp]
        lib.sse_oracle_get_H_dense.argtypes = [vp, C.c_double, vp]
        lib.sse_oracle_get_p_dense.argtypes = [vp, vp]
        lib.sse_oracle_get_x.argtypes = [vp, vp]
        lib.sse_oracle_get_ipiv.argtypes = [vp, C.c_double, C.c_double, vp]
        lib.sse_oracle_solve_ab.argtypes = [vp, C.c_double, C.c_double, vp]
        lib.sse_oracle_hamiltonian_dot_psi.argtypes = [vp, vp]
        lib.sse_oracle_reset_count.restype = C.c_long
        lib.sse_oracle_reset_count.argtypes = [vp]
        _libs[fast] = lib
    return _libs[fast]


VARIANTS = {"harmonic": 0, "inverted_harmonic": 1, "quartic": 2, "inverted_quartic": 2}


class Oracle:
    """One single-trajectory simulator, like one imported reference `simulation` module."""

    def __init__(self, variant, *, n_max=None, omega=np.pi, x_max=None, grid_size=None, lambda_=None,
                 mass=None, moment_order=5, herm_mode=0, fast=False):
        self.lib = _lib(fast)
        v = VARIANTS[variant] if isinstance(variant, str) else int(variant)
        cfg = _Cfg(v, 0 if n_max is None else n_max + 1, x_max or 0., grid_size or 0., lambda_ or 0.,
                   mass or 0., omega or 0., moment_order, herm_mode)
        self.cfg = cfg
        self.variant = v
        self.h = self.lib.sse_oracle_create(C.byref(cfg))
        self.n = self.lib.sse_oracle_n(self.h)
        self.moment_order = moment_order

    def __del__(self):
        try:
            self.lib.sse_oracle_destroy(self.h)
        except Exception:
            pass

    @staticmethod
    def _chk(state, n):
        assert isinstance(state, np.ndarray) and state.dtype == np.complex128 and state.shape == (n,) \
            and state.flags.c_contiguous

    def step(self, state, dt, F, gamma, r):
        """step(state, dt, F, gamma) of the reference (Q:493-525), in place; r = the two normals."""
        self._chk(state, self.n)
        r = np.ascontiguousarray(r, dtype=np.float64)
        q, xm, fail = C.c_double(), C.c_double(), C.c_int()
        rc = self.lib.sse_oracle_step(self.h, state.ctypes.data, dt, F, gamma, r.ctypes.data,
                                      C.byref(q), C.byref(xm), C.byref(fail))
        if rc:
            raise RuntimeError("oracle reset_ab failed: %d" % rc)
        return q.value, xm.value, fail.value

    def run(self, state, dt, F, gamma, noise, want_q=False):
        """nsub substeps at one force; returns (latched Fail, q[nsub] or None, x_mean[nsub] or None)."""
        self._chk(state, self.n)
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        nsub = noise.shape[0]
        q = np.empty(nsub) if want_q else None
        xm = np.empty(nsub) if want_q else None
        fail = C.c_int()
        rc = self.lib.sse_oracle_run(self.h, state.ctypes.data, dt, F, gamma, noise.ctypes.data, nsub,
                                     q.ctypes.data if want_q else None, xm.ctypes.data if want_q else None,
                                     C.byref(fail))
        if rc:
            raise RuntimeError("oracle reset_ab failed: %d" % rc)
        return fail.value, q, xm

    def x_expectation(self, state):
        self._chk(state, self.n)
        return self.lib.sse_oracle_x_expectation(self.h, state.ctypes.data)

    def get_moments(self, state, out=None):
        self._chk(state, self.n)
        M = self.moment_order
        if out is None:
            out = np.empty((2 + M + 1) * M // 2)
        rc = self.lib.sse_oracle_get_moments(self.h, state.ctypes.data, out.ctypes.data)
        if rc:
            raise RuntimeError("get_moments is grid-only")
        return out

    # ---- introspection -------------------------------------------------------------------------
    def A_dense(self, dt, F):
        out = np.empty((self.n, self.n), np.complex128)
        self.lib.sse_oracle_get_A_dense(self.h, dt, F, out.ctypes.data)
        return out

    def C_dense(self, dt, F):
        out = np.empty((self.n, self.n), np.complex128)
        self.lib.sse_oracle_get_C_dense(self.h, dt, F, out.ctypes.data)
        return out

    def H0_dense(self, F):
        out = np.empty((self.n, self.n), np.float64)
        self.lib.sse_oracle_get_H_dense(self.h, F, out.ctypes.data)
        return out

    def p_dense(self):
        out = np.empty((self.n, self.n), np.complex128)
        self.lib.sse_oracle_get_p_dense(self.h, out.ctypes.data)
        return out

    def x_array(self):
        out = np.empty(self.n)
        self.lib.sse_oracle_get_x(self.h, out.ctypes.data)
        return out

    def ipiv(self, dt, F):
        out = np.empty(self.n, np.int32)
        self.lib.sse_oracle_get_ipiv(self.h, dt, F, out.ctypes.data)
        return out

    def solve_ab(self, dt, F, state):
        self._chk(state, self.n)
        self.lib.sse_oracle_solve_ab(self.h, dt, F, state.ctypes.data)
        return state

    def hamiltonian_dot_psi(self, state):
        self._chk(state, self.n)
        self.lib.sse_oracle_hamiltonian_dot_psi(self.h, state.ctypes.data)
        return state

    def reset_count(self):
        return self.lib.sse_oracle_reset_count(self.h)
