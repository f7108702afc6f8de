"""Loader of the reference's own `simulation` modules built by oracle/build_ref.sh (reference .cpp + MKL-API shim).
TEST INFRASTRUCTURE ONLY; exists only where /root/reference was available at build time (oracle/_ref/ travels as built .so)."""
import ctypes
import importlib.machinery
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def available(task):
    return os.path.exists(os.path.join(_HERE, "_ref", task, "simulation.so"))


class RefModule:
    """The reference's compiled module for one task + the shim's noise hook."""

    def __init__(self, task):
        path = os.path.join(_HERE, "_ref", task, "simulation.so")
        loader = importlib.machinery.ExtensionFileLoader("simulation", path)
        spec = importlib.util.spec_from_loader("simulation", loader)
        self.mod = importlib.util.module_from_spec(spec)
        loader.exec_module(self.mod)
        self._c = ctypes.CDLL(path)
        self._c.qc_shim_set_normals.argtypes = [ctypes.c_void_p, ctypes.c_longlong]

    def set_normals(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64).ravel()
        self._c.qc_shim_set_normals(r.ctypes.data, r.size)

    def step(self, state, dt, F, gamma, normals):
        """reference step(state, dt, F, gamma) with the two normals of this substep injected into the shim's VSL stand-in."""
        self.set_normals(normals)
        return self.mod.step(state, dt, F, gamma)

    def __getattr__(self, name):
        return getattr(self.mod, name)
