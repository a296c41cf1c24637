"""ctypes binding of the C-ABI in include/mogpe_b200.h.

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a handle is
created, the product path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmogpe_b200.so")

MOGPE_OK = 0
BOUND_IDS = {"further_gating": 0, "tight": 1, "further": 2}
FLAVOUR_IDS = {"keras": 0, "legacy": 1}
PRECISION_IDS = {"float64": 0, "tf32": 1, "int8": 2, "int8x6": 3}
MAX_SAMPLES = 16
MAX_LATENTS = 64
MAX_GROUPS = 64


class MogpeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"mogpe_b200 error {status}: {message}")
        self.status = status


class Desc(C.Structure):
    _fields_ = [
        ("input_dim", C.c_int32), ("output_dim", C.c_int32), ("num_experts", C.c_int32),
        ("num_gating", C.c_int32), ("num_groups", C.c_int32), ("num_latents", C.c_int32),
        ("max_batch", C.c_int32), ("flavour", C.c_int32), ("device", C.c_int32), ("precision", C.c_int32),
        ("jitter", C.c_double),
        ("group_num_inducing", C.POINTER(C.c_int32)), ("latent_group", C.POINTER(C.c_int32)),
    ]


class Layout(C.Structure):
    _fields_ = [
        ("lengthscales", C.c_int64), ("variance", C.c_int64), ("Z", C.c_int64), ("q_mu", C.c_int64),
        ("q_sqrt", C.c_int64), ("mean_c", C.c_int64), ("noise", C.c_int64), ("total", C.c_int64),
        ("Mp", C.c_int32), ("reserved", C.c_int32),
    ]


_P = C.c_void_p
_PROTOS = {
    "mogpe_create": (C.c_int, [C.POINTER(Desc), C.POINTER(_P)]),
    "mogpe_destroy": (None, [_P]),
    "mogpe_last_error": (C.c_char_p, [_P]),
    "mogpe_param_layout": (C.c_int, [_P, C.POINTER(Layout)]),
    "mogpe_elbo_fwd_bwd": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_double, _P, _P, _P]),
    "mogpe_elbo_fwd_bwd_host": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_double, _P, _P, _P]),
    "mogpe_elbo_fwd_bwd_host_async": (C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_double, _P, _P, _P, _P]),
    "mogpe_predict": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "mogpe_predict_f_inducing_samples": (C.c_int, [_P, _P, C.c_int32, _P, _P, C.c_int32, _P, _P, _P]),
    "mogpe_prior_kl": (C.c_int, [_P, _P, _P, _P]),
    "mogpe_fill_normal": (C.c_int, [_P, _P, C.c_int64, C.c_uint64, C.c_uint64, _P]),
    "mogpe_check_numerics": (C.c_int, [_P, _P]),
    "mogpe_set_mc": (C.c_int, [_P, _P, C.c_int32]),
    "mogpe_set_seed": (C.c_int, [_P, C.c_uint64]),
    "mogpe_set_kl_weight": (C.c_int, [_P, C.c_double]),
    "mogpe_profile_enable": (C.c_int, [_P, C.c_int32]),
    "mogpe_train_step": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, C.c_double, _P, _P, _P, _P,
                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int32]),
    "mogpe_train_step_rows": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_double, _P, _P, _P, _P,
                              C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int32]),
    "mogpe_opt_state": (C.c_int, [_P, _P, _P, C.c_int32]),
    "mogpe_graph_stats": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mogpe_set_dp_overlap": (C.c_int, [_P, C.c_int32]),
    "mogpe_set_int8_backward": (C.c_int, [_P, C.c_int32]),
    "mogpe_set_int8_step": (C.c_int, [_P, C.c_int32]),
    "mogpe_profile_read_int8": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "mogpe_set_int8_grad_slices": (C.c_int, [_P, C.c_int32]),
    "mogpe_set_int8_forward_slices": (C.c_int, [_P, C.c_int32]),
    "mogpe_profile_read": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mogpe_profile_read_streaming": (C.c_int, [_P, _P, _P]),
    "mogpe_opt_setup": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_double]),
    "mogpe_opt_transform": (C.c_int, [_P, _P, _P, _P]),
    "mogpe_opt_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, _P]),
    "mogpe_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "mogpe_comm_init": (C.c_int, [_P, C.c_char_p, C.c_int32, C.c_int32]),
    "mogpe_allreduce_sum": (C.c_int, [_P, _P, C.c_int64, _P]),
    "mogpe_gather_rows": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P, _P]),
    "mogpe_malloc": (C.c_int, [C.c_int32, C.c_int64, C.POINTER(_P)]),
    "mogpe_free": (C.c_int, [C.c_int32, _P]),
    "mogpe_malloc_host": (C.c_int, [C.c_int64, C.POINTER(_P)]),
    "mogpe_free_host": (C.c_int, [_P]),
    "mogpe_memcpy_h2d": (C.c_int, [C.c_int32, _P, _P, C.c_int64, _P]),
    "mogpe_memcpy_d2h": (C.c_int, [C.c_int32, _P, _P, C.c_int64, _P]),
    "mogpe_memcpy_d2h_async": (C.c_int, [C.c_int32, _P, _P, C.c_int64, _P]),
    "mogpe_memset_zero": (C.c_int, [C.c_int32, _P, C.c_int64, _P]),
    "mogpe_stream_sync": (C.c_int, [C.c_int32, _P]),
    "mogpe_probe_fp64_tensor_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    "mogpe_probe_int8_tensor_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    "mogpe_debug_copy": (C.c_int, [_P, C.c_char_p, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "mogpe_debug_gemm": (C.c_int, [C.c_int32, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, _P]),
    "mogpe_debug_set_gemm_cfg": (C.c_int, [C.c_int32]),
    "mogpe_version": (C.c_char_p, []),
}
EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None


def load():
    """Load libmogpe_b200.so (built by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). mogpe_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, handle=None):
    if status != MOGPE_OK:
        msg = load().mogpe_last_error(handle)
        raise MogpeError(status, msg.decode() if msg else "?")
