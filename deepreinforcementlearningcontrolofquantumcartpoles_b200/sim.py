"""BatchedSim: thin Python owner of one libqcart handle.  PyTorch only holds the action / moment / flag tensors
(device memory + streams); all arithmetic happens in the hand-written CUDA kernels behind the C-ABI."""
import ctypes as C
import numpy as np

from . import _lib as L
from . import configs

_VARIANT = {"harmonic": L.QC_HARMONIC, "inverted_harmonic": L.QC_INV_HARMONIC, "quartic": L.QC_QUARTIC,
            "inverted_quartic": L.QC_QUARTIC}


def make_config(params, device=0):
    """dict (see configs.py) -> QcConfig.  Mirrors what the reference fixes through setupC.py -D macros and arguments.py."""
    v = params["variant"]
    c = L.QcConfig()
    c.struct_size = C.sizeof(L.QcConfig)
    c.variant = _VARIANT[v]
    if c.variant == L.QC_QUARTIC:
        c.n = 0
        c.x_max, c.grid_size = params["x_max"], params["grid_size"]
        c.lambda_, c.mass = params["lambda_"], params["mass"]
        c.moment_order = params.get("moment_order", 5)
        c.x_threshold = params.get("x_threshold", 0.0)
    else:
        c.n = params["n_max"] + 1
        c.omega = params["omega"]
        c.moment_order = 2
        c.herm_mode = params.get("herm_mode", 0)
    c.dt, c.gamma, c.n_sub = params["dt"], params["gamma"], params["n_sub"]
    c.f_max, c.n_levels = params["f_max"], params.get("n_levels", 21)
    c.device = device
    c.solve_tol = float(params.get("solve_tol", 0.0))      # 0 = library default (2^-48)
    return c


def _torch():
    import torch
    return torch


class BatchedSim:
    """B independent trajectories of one system on one GPU.

    step(action) = one control step: n_sub substeps of the reference's `simulation.step` (quartic
    simulation_quart.cpp:493-525) + get_moments (363-388) + the Fail latch, in ONE kernel launch.
    """

    def __init__(self, params, batch=None, device=0, seed=0, traj_offset=0):
        if isinstance(params, str):
            params = configs.PRESETS[params]()
        self.params = dict(params)
        self.lib = L.load()
        self.device = device
        self.cfg = make_config(self.params, device)
        h = C.c_void_p()
        L.check(self.lib.qc_create(C.byref(self.cfg), C.byref(h)))
        self.h = h
        rc = L.QcConfig()
        L.check(self.lib.qc_get_config(self.h, C.byref(rc)))
        self.cfg = rc
        self.n = self.lib.qc_state_len(self.h)
        self.K = self.lib.qc_num_moments(self.h)
        self.n_sub = self.cfg.n_sub
        self.B = 0
        self.seed = seed
        self.traj_offset = traj_offset
        if batch:
            self.set_batch(batch)

    def close(self):
        if getattr(self, "h", None):
            self.lib.qc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration ------------------------------------------------------------------------------------
    def check_settings(self):
        """check_settings() of the reference: (x_n, grid_size, lambda, mass, moment_order) or (n_max, omega)."""
        c = self.cfg
        if c.variant == L.QC_QUARTIC:
            return (c.n, c.grid_size, c.lambda_, c.mass, c.moment_order)
        return (c.n - 1, c.omega)

    def level_force(self, level):
        return self.lib.qc_level_force(self.h, int(level))

    def kernel_info(self):
        return self.lib.qc_kernel_info(self.h).decode()

    def launch_count(self):
        return self.lib.qc_launch_count(self.h)

    def x_grid(self):
        c = self.cfg
        half = (self.n - 1) // 2
        return c.grid_size * (np.arange(self.n) - half).astype(np.float64)

    # ---- state --------------------------------------------------------------------------------------------
    def _stream(self):
        torch = _torch()
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_batch(self, B):
        L.check(self.lib.qc_set_batch(self.h, int(B)))
        self.B = int(B)
        L.check(self.lib.qc_set_seed(self.h, int(self.seed), int(self.traj_offset)))

    def set_seed(self, seed, traj_offset=None):
        self.seed = int(seed)
        if traj_offset is not None:
            self.traj_offset = int(traj_offset)
        L.check(self.lib.qc_set_seed(self.h, self.seed, int(self.traj_offset)))

    def set_state(self, psi):
        """psi: [B, N] complex128 numpy array or CUDA torch tensor."""
        torch = _torch()
        if isinstance(psi, np.ndarray):
            a = np.ascontiguousarray(psi, dtype=np.complex128)
            assert a.shape == (self.B, self.n), (a.shape, (self.B, self.n))
            L.check(self.lib.qc_set_state(self.h, a.ctypes.data, 0, self._stream()))
        else:
            assert psi.is_cuda and psi.dtype == torch.complex128 and tuple(psi.shape) == (self.B, self.n)
            psi = psi.contiguous()
            L.check(self.lib.qc_set_state(self.h, psi.data_ptr(), 1, self._stream()))

    def get_state(self, numpy=True):
        torch = _torch()
        if numpy:
            out = np.empty((self.B, self.n), np.complex128)
            L.check(self.lib.qc_get_state(self.h, out.ctypes.data, 0, self._stream()))
            return out
        out = torch.empty((self.B, self.n), dtype=torch.complex128, device="cuda:%d" % self.device)
        L.check(self.lib.qc_get_state(self.h, out.data_ptr(), 1, self._stream()))
        return out

    def clear_flags(self):
        L.check(self.lib.qc_clear_flags(self.h, self._stream()))

    def init_packets(self, wavenumber=None, mean=None, std=1.0):
        """Gaussian_packet of quartic main_parallel.py:75-76 for every trajectory (numpy arrays or CUDA float64 tensors of length B, or None)."""
        torch = _torch()
        on_dev = any(isinstance(v, torch.Tensor) for v in (wavenumber, mean))
        if on_dev:
            k = None if wavenumber is None else wavenumber.to(torch.float64).contiguous()
            m = None if mean is None else mean.to(torch.float64).contiguous()
            assert all(v is None or (v.is_cuda and v.numel() == self.B) for v in (k, m))
            L.check(self.lib.qc_init_packets(self.h, None if k is None else k.data_ptr(), None if m is None else m.data_ptr(), float(std), 1, self._stream()))
            return
        k = None if wavenumber is None else np.ascontiguousarray(wavenumber, np.float64)
        m = None if mean is None else np.ascontiguousarray(mean, np.float64)
        L.check(self.lib.qc_init_packets(self.h, None if k is None else k.ctypes.data, None if m is None else m.ctypes.data,
                                         float(std), 0, self._stream()))

    def reset_accept(self, aux, energy_cutoff, pending, store, n_pending):
        """qc_reset_accept: device-side accept / keep-pending step of the rejection loop of the quartic reset (all CUDA tensors)."""
        torch = _torch()
        assert aux.is_cuda and aux.dtype == torch.float64 and pending.dtype == torch.uint8 and store.dtype == torch.complex128 and n_pending.dtype == torch.int32
        assert tuple(store.shape) == (self.B, self.n) and store.is_contiguous() and pending.numel() == self.B
        L.check(self.lib.qc_reset_accept(self.h, aux.data_ptr(), float(energy_cutoff), pending.data_ptr(), store.data_ptr(), n_pending.data_ptr(), self._stream()))

    def reset_scatter(self, mask, slot, pool):
        """qc_reset_scatter: trajectories with mask != 0 restart from pool[slot mod len(pool)] (CUDA tensors: uint8 [B], int64 [B], complex128 [P, N])."""
        torch = _torch()
        assert mask.is_cuda and mask.dtype == torch.uint8 and slot.dtype == torch.int64 and pool.dtype == torch.complex128 and pool.is_contiguous() and pool.shape[1] == self.n
        L.check(self.lib.qc_reset_scatter(self.h, mask.data_ptr(), slot.data_ptr(), pool.data_ptr(), int(pool.shape[0]), self._stream()))

    def init_fock(self, alpha=None):
        """Fock vacuum (harmonic main_parallel.py:226-227) or coherent states alpha[B] (complex)."""
        if alpha is None:
            L.check(self.lib.qc_init_fock(self.h, None, 0, self._stream()))
        else:
            a = np.ascontiguousarray(alpha, np.complex128)
            assert a.shape == (self.B,)
            L.check(self.lib.qc_init_fock(self.h, a.ctypes.data, 0, self._stream()))
            _torch().cuda.current_stream(self.device).synchronize()

    # ---- the hot path -------------------------------------------------------------------------------------
    def alloc_outputs(self, n_sub=None, want_q=False):
        torch = _torch()
        dev = "cuda:%d" % self.device
        out = {"moments": torch.empty((self.B, self.K), dtype=torch.float64, device=dev),
               "aux": torch.empty((self.B, L.QC_AUX_COUNT), dtype=torch.float64, device=dev),
               "flags": torch.empty((self.B,), dtype=torch.uint8, device=dev)}
        if want_q:
            ns = n_sub or self.n_sub
            out["q"] = torch.empty((self.B, ns), dtype=torch.float64, device=dev)
            out["x_mean"] = torch.empty((self.B, ns), dtype=torch.float64, device=dev)
        return out

    def step(self, action, noise=None, n_sub=None, nsub_traj=None, out=None, want_q=False):
        """One control step.  action: int32 CUDA tensor [B] of force levels.  noise: float64 CUDA tensor
        [B, n_sub, 2] of standard normals (verification) or None (in-kernel Philox).  Returns dict of CUDA tensors."""
        torch = _torch()
        assert action.is_cuda and action.dtype == torch.int32 and action.numel() == self.B and action.is_contiguous()
        ns = int(n_sub or self.n_sub)
        if noise is not None:
            assert noise.is_cuda and noise.dtype == torch.float64 and tuple(noise.shape) == (self.B, ns, 2) and noise.is_contiguous()
        if nsub_traj is not None:
            assert nsub_traj.is_cuda and nsub_traj.dtype == torch.int32 and nsub_traj.numel() == self.B
        if out is None:
            out = self.alloc_outputs(ns, want_q)
        q = out.get("q")
        xm = out.get("x_mean")
        L.check(self.lib.qc_step(self.h, action.data_ptr(), None if noise is None else noise.data_ptr(), ns,
                                 None if nsub_traj is None else nsub_traj.data_ptr(),
                                 out["moments"].data_ptr(), out["aux"].data_ptr(), out["flags"].data_ptr(),
                                 None if q is None else q.data_ptr(), None if xm is None else xm.data_ptr(), self._stream()))
        return out

    def step_forces(self, forces, noise=None, n_sub=None, out=None, want_q=False):
        """Same with arbitrary per-trajectory force values (host array of doubles)."""
        f = np.ascontiguousarray(forces, np.float64)
        assert f.shape == (self.B,)
        ns = int(n_sub or self.n_sub)
        if out is None:
            out = self.alloc_outputs(ns, want_q)
        q = out.get("q")
        xm = out.get("x_mean")
        L.check(self.lib.qc_step_forces(self.h, f.ctypes.data, None if noise is None else noise.data_ptr(), ns, None,
                                        out["moments"].data_ptr(), out["aux"].data_ptr(), out["flags"].data_ptr(),
                                        None if q is None else q.data_ptr(), None if xm is None else xm.data_ptr(), self._stream()))
        return out

    def step_host(self, action, noise=None, n_sub=None, moments=None, aux=None, flags=None):
        """End-to-end call with HOST buffers (numpy or pinned torch tensors): H2D action(+noise), kernel, D2H results."""
        ns = int(n_sub or self.n_sub)
        a = action if isinstance(action, np.ndarray) else action.numpy()
        assert a.dtype == np.int32 and a.size == self.B
        moments = np.empty((self.B, self.K)) if moments is None else moments
        aux = np.empty((self.B, L.QC_AUX_COUNT)) if aux is None else aux
        flags = np.empty((self.B,), np.uint8) if flags is None else flags
        ptr = lambda t: t.ctypes.data if isinstance(t, np.ndarray) else t.data_ptr()
        L.check(self.lib.qc_step_host(self.h, ptr(a), None if noise is None else ptr(noise), ns, ptr(moments), ptr(aux), ptr(flags)))
        return moments, aux, flags

    def get_moments(self):
        out = self.alloc_outputs()
        L.check(self.lib.qc_get_moments(self.h, out["moments"].data_ptr(), out["aux"].data_ptr(), self._stream()))
        return out

    # ---- single-trajectory shims (reference signatures) ---------------------------------------------------
    def step1(self, state, dt, F, gamma, normals=None):
        q, xm, f = C.c_double(), C.c_double(), C.c_int()
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float64)
        L.check(self.lib.qc_step1(self.h, state.ctypes.data, dt, F, gamma, None if nrm is None else nrm.ctypes.data,
                                  C.byref(q), C.byref(xm), C.byref(f)))
        return q.value, xm.value, f.value

    def simulate_10_steps1(self, state, dt, F, gamma, normals=None):
        q, xm, f = C.c_double(), C.c_double(), C.c_int()
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float64)
        L.check(self.lib.qc_simulate_10_steps1(self.h, state.ctypes.data, dt, F, gamma, None if nrm is None else nrm.ctypes.data,
                                               C.byref(q), C.byref(xm), C.byref(f)))
        return q.value, xm.value, f.value

    def get_moments1(self, state, out):
        L.check(self.lib.qc_get_moments1(self.h, state.ctypes.data, out.ctypes.data))

    def x_expectation1(self, state):
        v = C.c_double()
        L.check(self.lib.qc_x_expectation1(self.h, state.ctypes.data, C.byref(v)))
        return v.value

    def hamiltonian_dot_psi1(self, state):
        """state <- H state (zero force), in place (harmonic simulation.cpp:566-582)."""
        L.check(self.lib.qc_hamiltonian_dot_psi1(self.h, state.ctypes.data))

    def solve_ab1(self, state, F):
        """state <- (I + i dt/2 (H - kappa F x))^-1 state, in place, exact band substitution (harmonic simulation.cpp:584-597)."""
        L.check(self.lib.qc_solve_ab1(self.h, state.ctypes.data, float(F)))


def philox_normals(seed, traj, step):
    """Host restatement of the in-kernel noise: the (r0, r1) of (seed, global trajectory id, substep counter)."""
    out = np.empty(2)
    L.load().qc_philox_normals(int(seed), int(traj), int(step), out.ctypes.data)
    return out


def measure_peaks(device=0):
    """Measured FP64 FLOP/s and shared-memory bytes/s of the device (micro-benchmarks in the library)."""
    lib = L.load()
    a, b = C.c_double(), C.c_double()
    L.check(lib.qc_measure_fp64_peak(device, C.byref(a)))
    L.check(lib.qc_measure_smem_peak(device, C.byref(b)))
    return a.value, b.value
