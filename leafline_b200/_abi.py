"""ctypes binding of the C ABI declared in include/leafline_b200.h.

The shared library is built in-tree by __graft_entry__.build() (nvcc, sm_100a).  There is no
fallback of any kind: if the library is missing, or no CUDA device is present, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LEAFLINE_B200_LIB") or os.path.join(_HERE, "libleafline_b200.so")  # the env override is for A/B builds

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_CAPACITY, ERR_DOMAIN, ERR_NO_DEVICE, ERR_OOM, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6, -7

WORLD_DTYPE = np.dtype(
    [("plane", "<u8", (12,)), ("initiative", "u1"), ("service_eligibility", "u1"), ("passing_by", "u1"),
     ("pad", "u1", (5,))]
)
COMMIT_DTYPE = np.dtype(
    [("team", "u1"), ("job", "u1"), ("whence", "u1"), ("whither", "u1"), ("hospitalization", "u1"),
     ("ascension", "u1"), ("pad", "u1", (2,)), ("tree", WORLD_DTYPE)]
)
MAX_VARIATION = 16
PATCH_DTYPE = np.dtype([("team", "u1"), ("job", "u1"), ("whence", "u1"), ("whither", "u1")])
FORECAST_DTYPE = np.dtype(
    [("commit", COMMIT_DTYPE), ("score", "<f4"), ("variation_len", "<u4"), ("variation", PATCH_DTYPE, (MAX_VARIATION,))]
)
assert WORLD_DTYPE.itemsize == 104 and COMMIT_DTYPE.itemsize == 112 and FORECAST_DTYPE.itemsize == 184


class Soa(C.Structure):
    _fields_ = [("planes", C.c_void_p), ("meta", C.c_void_p), ("stride", C.c_size_t), ("n", C.c_size_t)]


class LeaflineError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"leafline_b200 error {code}: {message}")
        self.code = code


# every symbol include/leafline_b200.h declares, with its ctypes signature
_vp, _i, _sz, _u64 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64
SIGNATURES = {
    "ll_init": ([_i, C.POINTER(_vp)], _i),
    "ll_shutdown": ([_vp], _i),
    "ll_last_error": ([], C.c_char_p),
    "ll_set_stream": ([_vp, _vp], _i),
    "ll_synchronize": ([_vp], _i),
    "ll_device_sm_count": ([_vp], _i),
    "ll_score_batch": ([_vp, _vp, _sz, _vp], _i),
    "ll_lookahead_batch": ([_vp, _vp, _sz, _i, _vp, _vp, _sz], _i),
    "ll_in_critical_endangerment_batch": ([_vp, _vp, _vp, _sz, _vp], _i),
    "ll_perft": ([_vp, _vp, _i, _i, _i, C.POINTER(_u64), _vp, C.c_uint32, C.POINTER(C.c_uint32)], _i),
    "ll_perft_device": ([_vp, _vp, _i, _i, _i, _vp], _i),
    "ll_horizon_batch": ([_vp, _vp, _sz, _i, _i, _vp], _i),
    "ll_kickoff": ([_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(_u64)], _i),
    "ll_multi_unique_id": ([_vp], _i),
    "ll_init_multi": ([C.POINTER(_i), _i, C.POINTER(_vp)], _i),
    "ll_init_multi_rank": ([_i, _i, _i, _vp, C.POINTER(_vp)], _i),
    "ll_shutdown_multi": ([_vp], _i),
    "ll_multi_world": ([_vp], _i),
    "ll_multi_local_count": ([_vp], _i),
    "ll_multi_ctx": ([_vp, _i], _vp),
    "ll_multi_reduce_path": ([_vp], _i),
    "ll_multi_set_shard_min_depth": ([_vp, _i, _i], _i),
    "ll_perft_multi": ([_vp, _vp, _i, C.POINTER(_u64), _vp, C.c_uint32, C.POINTER(C.c_uint32)], _i),
    "ll_perft_batch_multi": ([_vp, _vp, C.c_uint32, _i, _vp], _i),
    "ll_kickoff_multi": ([_vp, _vp, _i, _i, _i, _vp, _vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(_u64)], _i),
    "ll_forecast": ([_vp, _vp, _i, _i, _i, _vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(_u64)], _i),
    "ll_alpha_beta_forecast": ([_vp, _vp, _i, _i, _i, _i, _vp, C.c_uint32, C.POINTER(C.c_uint32), _vp], _i),
    "ll_alpha_beta_forecast_multi": ([_vp, _vp, _i, _i, _i, _i, _vp, C.c_uint32, C.POINTER(C.c_uint32), _vp], _i),
    "ll_set_deja_vu_bound": ([_vp, C.c_double], _i),
    "ll_set_search_threads": ([_vp, _i], _i),
    "ll_fixed_depth_sequence_forecast": ([_vp, _vp, _vp, C.c_uint32, _i, _vp, C.c_uint32, C.POINTER(C.c_uint32)], _i),
    "ll_iterative_deepening_forecast": ([_vp, _vp, C.c_double, _i, _vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)], _i),
    "ll_soa_alloc": ([_vp, _sz, C.POINTER(Soa)], _i),
    "ll_soa_free": ([_vp, C.POINTER(Soa)], _i),
    "ll_soa_upload": ([_vp, _vp, _sz, C.POINTER(Soa)], _i),
    "ll_soa_download": ([_vp, C.POINTER(Soa), _sz, _sz, _vp], _i),
    "ll_score_soa": ([_vp, C.POINTER(Soa), _vp], _i),
    "ll_horizon_soa": ([_vp, C.POINTER(Soa), _i, _i, _vp, C.POINTER(_u64)], _i),
    "ll_playout_soa": ([_vp, _u64, _u64, _sz, C.POINTER(Soa)], _i),
    "ll_timing_enable": ([_vp, _i], _i),
    "ll_timing_reset": ([_vp], _i),
    "ll_timing_query": ([_vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(_u64)], _i),
    "ll_launch_count": ([_vp], _u64),
    "ll_debug_force_synced": ([_vp, _i], _i),
    "ll_debug_selfcheck_soa": ([_vp, C.POINTER(Soa), _vp], _i),
    "ll_debug_tables": ([_vp, _vp, _vp], _i),
    "ll_debug_horizon_overflowed": ([_vp, _vp], _i),
    "ll_measure_int_peak": ([_vp, _i, C.POINTER(C.c_double)], _i),
}

_lib = None


def load():
    """Load libleafline_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  leafline_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


def check(rc):
    if rc != OK:
        raise LeaflineError(rc, load().ll_last_error().decode(errors="replace"))
