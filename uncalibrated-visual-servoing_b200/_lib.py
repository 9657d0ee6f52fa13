"""ctypes binding of libuvs_b200.so (include/uvs_b200.h).

The library is the product: if it is missing this module raises — there is no
CPU or PyTorch fallback anywhere in the package.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UVS_B200_LIB") or os.path.join(HERE, "libuvs_b200.so")   # override: tuning builds only

UVS_MAX_SWEEP = 64

# experiment.py:6-11 / noise.py:7-12 numeric values
METHOD_ANALYTICAL, METHOD_KF, METHOD_MCKF, METHOD_IMCCKF, METHOD_GMCKF = 1, 2, 3, 4, 5
NOISE_NONE, NOISE_WHITE, NOISE_MIXTURE, NOISE_BIMODAL, NOISE_ALPHA_STABLE, NOISE_UNIFORM = 0, 1, 2, 3, 4, 5
MODE_AUTO, MODE_DENSE, MODE_BLOCKDIAG = 0, 1, 2
P_DENSE, P_BLOCKS = 0, 1
FLAG_SKIP_SINGULAR, FLAG_SKIP_MAXEPOCH, FLAG_CX_SINGULAR, FLAG_NOT_SPD = 1, 2, 4, 8
FLAG_RANK_DEFICIENT, FLAG_NONFINITE, FLAG_OUT_OF_VIEW = 16, 32, 64
FLAG_EPOCH_SHIFT = 8
# tests/-script pipeline (uvs_twist_cfg)
X0_ZEROS, X0_IMAGE_JACOBIAN, X0_RANDOM = 0, 1, 2
DEPTH_DISTANCE, DEPTH_HEIGHT = 0, 1

_dp = C.c_void_p


class FilterParams(C.Structure):
    _fields_ = [("method", C.c_int32), ("nf", C.c_int32), ("mode", C.c_int32), ("joseph", C.c_int32),
                ("mckf_skip", C.c_int32), ("fpi_epoch_max", C.c_int32), ("annealing", C.c_int32),
                ("reserved", C.c_int32), ("kernel_bw", C.c_double), ("fpi_threshold", C.c_double),
                ("Q", _dp), ("R", _dp)]


class NoiseParams(C.Structure):
    _fields_ = [("type", C.c_int32), ("hold_cnt", C.c_int32), ("std", C.c_double), ("mean", C.c_double),
                ("rho", C.c_double), ("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double),
                ("delta", C.c_double), ("seed", C.c_uint64)]


class ClosedLoopCfg(C.Structure):
    _fields_ = [("filt", FilterParams), ("noise", NoiseParams), ("T", C.c_int32), ("initial_guess", C.c_int32),
                ("change_q_start", C.c_int32), ("fov_check", C.c_int32), ("sweep_len", C.c_int32),
                ("sweep_epoch", C.c_int32), ("sweep_param", C.c_int32), ("trial_stride", C.c_int32),
                ("dt", C.c_double), ("ibvs_gain", C.c_double), ("q_start", C.c_double * 6),
                ("desired_f", C.c_double * 8), ("targets", C.c_double * 12), ("pose_seed", C.c_uint64),
                ("sweep", C.c_double * UVS_MAX_SWEEP), ("P0", _dp), ("v_max", C.c_double), ("a_max", C.c_double), ("servo_gain", C.c_double),
                ("sensor_offset", C.c_double * 3), ("disc_radius", C.c_double), ("servo_kp", C.c_double),
                ("servo_substeps", C.c_int32), ("reserved2", C.c_int32)]


class ClosedLoopIO(C.Structure):
    _fields_ = [("stats", _dp), ("status", _dp), ("frames", _dp), ("flags", _dp), ("log", _dp), ("X_out", _dp),
                ("P_out", _dp), ("noise_in", _dp), ("q0_in", _dp), ("X0_in", _dp), ("p_layout", C.c_int32),
                ("reserved", C.c_int32), ("dq_in", _dp), ("X_log", _dp), ("P_log", _dp), ("dq_log", _dp),
                ("kappa_log", _dp), ("flags_log", _dp)]


class TwistCfg(C.Structure):
    _fields_ = [("x0_mode", C.c_int32), ("depth_mode", C.c_int32), ("dp_zero_mask", C.c_int32),
                ("stale_fk", C.c_int32), ("focal_x0", C.c_double), ("focal", C.c_double),
                ("err_threshold", C.c_double)]


# int sink(void* user, uint64_t first_trial, int64_t n, const double* log, const double* stats, const int32_t* status,
#          const int32_t* frames)  — include/uvs_b200.h: uvs_log_sink
LOG_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double),
                       C.POINTER(C.c_int32), C.POINTER(C.c_int32))

_PROTOTYPES = {
    "uvs_version": ([], C.c_int),
    "uvs_last_error": ([C.c_char_p, C.c_size_t], C.c_int),
    "uvs_device_info": ([C.POINTER(C.c_int)] * 4, C.c_int),
    "uvs_filter_step": ([C.POINTER(FilterParams), C.c_int64, _dp, _dp, C.c_int, _dp, _dp, C.c_double, _dp, _dp, _dp,
                         _dp], C.c_int),
    "uvs_replay": ([C.POINTER(FilterParams), C.c_int64, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp,
                    _dp, _dp, _dp, _dp], C.c_int),
    "uvs_ibvs": ([C.c_int32, C.c_int64, _dp, _dp, _dp, C.c_double, _dp, _dp, _dp], C.c_int),
    "uvs_ur10_fk": ([C.c_int64, _dp, _dp, _dp, _dp], C.c_int),
    "uvs_project": ([C.c_int32, C.c_int64, _dp, C.POINTER(C.c_double), _dp, _dp, _dp], C.c_int),
    "uvs_initial_guess": ([C.c_int32, C.c_int64, _dp, C.POINTER(C.c_double), _dp, _dp, _dp], C.c_int),
    "uvs_noise_fill": ([C.POINTER(NoiseParams), C.c_uint64, C.c_int64, C.c_int32, C.c_int32, _dp, _dp], C.c_int),
    "uvs_start_poses": ([C.POINTER(C.c_double), C.c_uint64, C.c_int32, C.c_uint64, C.c_int64, _dp, _dp], C.c_int),
    "uvs_closed_loop": ([C.POINTER(ClosedLoopCfg), C.c_uint64, C.c_int64, C.POINTER(ClosedLoopIO), _dp], C.c_int),
    "uvs_twist_loop": ([C.POINTER(ClosedLoopCfg), C.POINTER(TwistCfg), C.c_uint64, C.c_int64, C.POINTER(ClosedLoopIO),
                        _dp], C.c_int),
    "uvs_closed_loop_host": ([C.POINTER(ClosedLoopCfg), C.c_uint64, C.c_int64, C.POINTER(C.c_double),
                              C.POINTER(C.c_int32), C.POINTER(C.c_int32)], C.c_int),
    "uvs_closed_loop_stream": ([C.POINTER(ClosedLoopCfg), C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_double), C.POINTER(C.c_double)], C.c_int),
    "uvs_reduce_stats": ([C.c_int64, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp], C.c_int),
    "uvs_fp64_peak": ([C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)], C.c_int),
}

EXPORTED = tuple(_PROTOTYPES)
_lib = None


class UvsError(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UvsError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def last_error():
    buf = C.create_string_buffer(512)
    load().uvs_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc, what):
    if rc != 0:
        raise UvsError(f"{what} failed ({rc}): {last_error()}")
