"""ctypes binding of libqcart.so (include/qcart.h).  No CPU fallback: a missing library or device raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QCART_LIB") or os.path.join(_HERE, "libqcart.so")   # QCART_LIB: A/B-test another build of the same ABI

QC_HARMONIC, QC_INV_HARMONIC, QC_QUARTIC = 0, 1, 2
QC_FLAG_FAIL, QC_FLAG_ESCAPED = 1, 2
QC_AUX_ENERGY, QC_AUX_XMEAN, QC_AUX_OUTSIDE, QC_AUX_NORM, QC_AUX_COUNT = 0, 1, 2, 3, 4
QC_ERR_ARG, QC_ERR_CUDA, QC_ERR_PIVOT, QC_ERR_UNSUPPORTED, QC_ERR_STATE = -1, -2, -3, -4, -5
QC_NOISE_OFF, QC_NOISE_GIVEN, QC_NOISE_PHILOX = 0, 1, 2
# qc_policy parameter tensors (include/qcart_rollout.h), in ABI order
POLICY_PARAMS = ["FC1_W", "FC1_B", "FC2_W", "FC2_B", "FC31_UW", "FC31_SW", "FC31_UB", "FC31_SB", "FC41_UW", "FC41_SW", "FC41_UB", "FC41_SB",
                 "FC32_W", "FC32_B", "FC42_W", "FC42_B"]


class QcConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("variant", C.c_int32), ("n", C.c_int32), ("x_max", C.c_double), ("grid_size", C.c_double),
                ("lambda_", C.c_double), ("mass", C.c_double), ("omega", C.c_double), ("dt", C.c_double),
                ("gamma", C.c_double), ("n_sub", C.c_int32), ("f_max", C.c_double), ("n_levels", C.c_int32),
                ("moment_order", C.c_int32), ("x_threshold", C.c_double), ("herm_mode", C.c_int32),
                ("device", C.c_int32), ("solve_tol", C.c_double)]


class QcartError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libqcart error %d: %s" % (code, msg))
        self.code = code


# every symbol include/qcart.h and include/qcart_rollout.h declare: (name, restype, argtypes)
_vp, _dp, _ip, _i64 = C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64
SYMBOLS = [
    ("qc_last_error", C.c_char_p, []),
    ("qc_version", C.c_char_p, []),
    ("qc_config_size", C.c_uint32, []),
    ("qc_create", C.c_int, [C.POINTER(QcConfig), C.POINTER(_vp)]),
    ("qc_destroy", C.c_int, [_vp]),
    ("qc_get_config", C.c_int, [_vp, C.POINTER(QcConfig)]),
    ("qc_state_len", C.c_int, [_vp]),
    ("qc_num_moments", C.c_int, [_vp]),
    ("qc_num_aux", C.c_int, [_vp]),
    ("qc_set_batch", C.c_int, [_vp, _i64]),
    ("qc_batch", _i64, [_vp]),
    ("qc_set_state", C.c_int, [_vp, _dp, C.c_int, _vp]),
    ("qc_get_state", C.c_int, [_vp, _dp, C.c_int, _vp]),
    ("qc_state_ptr", C.c_void_p, [_vp]),
    ("qc_set_seed", C.c_int, [_vp, C.c_uint64, _i64]),
    ("qc_clear_flags", C.c_int, [_vp, _vp]),
    ("qc_init_packets", C.c_int, [_vp, _dp, _dp, C.c_double, C.c_int, _vp]),
    ("qc_init_fock", C.c_int, [_vp, _dp, C.c_int, _vp]),
    ("qc_reset_accept", C.c_int, [_vp, _dp, C.c_double, _vp, _dp, _vp, _vp]),
    ("qc_reset_scatter", C.c_int, [_vp, _vp, _vp, _dp, _i64, _vp]),
    ("qc_step", C.c_int, [_vp, _ip, _dp, C.c_int, _ip, _dp, _dp, _vp, _dp, _dp, _vp]),
    ("qc_step_forces", C.c_int, [_vp, _dp, _dp, C.c_int, _ip, _dp, _dp, _vp, _dp, _dp, _vp]),
    ("qc_step_host", C.c_int, [_vp, _ip, _dp, C.c_int, _dp, _dp, _vp]),
    ("qc_get_moments", C.c_int, [_vp, _dp, _dp, _vp]),
    ("qc_level_force", C.c_double, [_vp, C.c_int]),
    ("qc_step1", C.c_int, [_vp, _dp, C.c_double, C.c_double, C.c_double, _dp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    ("qc_simulate_10_steps1", C.c_int, [_vp, _dp, C.c_double, C.c_double, C.c_double, _dp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    ("qc_get_moments1", C.c_int, [_vp, _dp, _dp]),
    ("qc_x_expectation1", C.c_int, [_vp, _dp, C.POINTER(C.c_double)]),
    ("qc_hamiltonian_dot_psi1", C.c_int, [_vp, _dp]),
    ("qc_solve_ab1", C.c_int, [_vp, _dp, C.c_double]),
    ("qc_philox_normals", None, [C.c_uint64, C.c_uint64, C.c_uint64, _dp]),
    ("qc_measure_fp64_peak", C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    ("qc_measure_smem_peak", C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    ("qc_launch_count", _i64, [_vp]),
    ("qc_kernel_info", C.c_char_p, [_vp]),
    ("qc_peer_alloc", C.c_int, [C.c_int32, C.c_uint64, C.POINTER(_vp), _vp]),
    ("qc_peer_free", C.c_int, [C.c_int32, _vp]),
    ("qc_peer_open", C.c_int, [C.c_int32, _vp, C.POINTER(_vp)]),
    ("qc_peer_close", C.c_int, [C.c_int32, _vp]),
    ("qc_set_gather", C.c_int, [_vp, C.c_int32, C.c_int32, C.POINTER(_vp), C.POINTER(_vp)]),
    ("qc_gather_seq", C.c_uint64, [_vp]),
    ("qc_gather_wait", C.c_int, [_vp, C.c_uint64, _vp]),
    ("qc_gather_error", C.c_int, [_vp, C.POINTER(C.c_uint32)]),
    # ---- include/qcart_rollout.h ----
    ("qc_obs_f32", C.c_int, [_dp, _i64, C.c_double, _vp, _vp]),
    ("qc_policy_create", C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    ("qc_policy_destroy", C.c_int, [_vp]),
    ("qc_policy_param_size", _i64, [_vp, C.c_int32]),
    ("qc_policy_set_param", C.c_int, [_vp, C.c_int32, _vp, _i64]),
    ("qc_policy_noise_width", _i64, [_vp]),
    ("qc_policy_forward", C.c_int, [_vp, _vp, _i64, C.c_int32, _vp, C.c_uint64, _i64, C.c_uint64, _vp, _vp, _vp, _vp]),
    ("qc_epsilon_greedy", C.c_int, [_vp, _i64, C.c_int32, C.c_double, C.c_uint64, _i64, C.c_uint64, _vp, _vp, C.c_int32, _vp]),
    ("qc_action_forces", C.c_int, [_vp, _i64, C.c_int32, C.c_double, _vp, C.c_int32, _vp]),
    ("qc_policy_launch_count", _i64, [_vp]),
    ("qc_policy_set_gemm", C.c_int, [_vp, C.c_int32]),
    ("qc_replay_create", C.c_int, [C.c_int32, _i64, C.c_int32, C.POINTER(_vp)]),
    ("qc_replay_destroy", C.c_int, [_vp]),
    ("qc_replay_push", C.c_int, [_vp, _vp, _vp, C.c_int32, _vp, _vp, _i64, C.c_double, _vp, _i64, _vp]),
    ("qc_replay_total", C.c_int, [_vp, C.POINTER(_i64), _vp]),
    ("qc_replay_data", C.c_void_p, [_vp]),
    ("qc_replay_read", C.c_int, [_vp, _i64, _i64, _vp, _vp]),
    ("qc_record_create", C.c_int, [_i64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    ("qc_record_destroy", C.c_int, [_vp]),
    ("qc_record_reset", C.c_int, [_vp, _vp, _vp]),
    ("qc_record_push", C.c_int, [_vp, _vp, C.c_int32, _vp, C.c_double, _vp]),
    ("qc_record_window", C.c_int, [_vp, _vp, _vp]),
    ("qc_record_experience", C.c_int, [_vp, _vp, _vp]),
    ("qc_record_row_len", _i64, [_vp]),
]

_lib = None


def load():
    """Load libqcart.so; raises if it has not been built (python __graft_entry__.py / csrc/build.sh)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libqcart.so is not built (run deepreinforcementlearningcontrolofquantumcartpoles_b200/csrc/build.sh); "
                              "there is no CPU fallback for the SSE hot path")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.qc_config_size() != C.sizeof(QcConfig):
            raise ImportError("libqcart.so was built from a different include/qcart.h: qc_config has %d bytes there, %d here" % (lib.qc_config_size(), C.sizeof(QcConfig)))
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise QcartError(rc, load().qc_last_error().decode())
