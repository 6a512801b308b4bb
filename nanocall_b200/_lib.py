"""ctypes loader for nanocall_b200/libnanocall_b200.so (the C ABI in include/nanocall_b200.h).

Fails loudly when the library is missing: there is no Python/NumPy fallback for any compute call.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# NC_LIB_PATH: load another build of the same library (kernel A/B experiments); never a different implementation
LIB_PATH = os.environ.get("NC_LIB_PATH") or os.path.join(_HERE, "libnanocall_b200.so")

NC_OK, NC_ERR_ARG, NC_ERR_CUDA, NC_ERR_NOMEM, NC_ERR_STATE = 0, -1, -2, -3, -4
NC_MEM_HOST, NC_MEM_DEVICE = 0, 1
NC_VIT_AUTO, NC_VIT_BACKPOINTER = 0, 2


class PmParams(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("scale", "shift", "drift", "var", "scale_sd", "var_sd")]


class StParams(C.Structure):
    _fields_ = [("p_stay", C.c_float), ("p_skip", C.c_float)]


class VitJob(C.Structure):
    _fields_ = [("mean", C.c_void_p), ("stdv", C.c_void_p), ("start", C.c_void_p),
                ("n_events", C.c_uint32), ("model_id", C.c_int32), ("pm", PmParams), ("st", StParams)]


class VitOut(C.Structure):
    _fields_ = [("path_logprob", C.c_float), ("states", C.c_void_p), ("moves", C.c_void_p),
                ("bases", C.c_void_p), ("bases_cap", C.c_uint32), ("n_bases", C.c_uint32)]


class TrainIn(C.Structure):
    _fields_ = [("model_id", C.c_int32 * 2), ("pm", PmParams), ("st", StParams * 2)]


class TrainOut(C.Structure):
    _fields_ = [("pm", PmParams), ("st", StParams * 2), ("fit", C.c_float), ("done", C.c_int32)]


class TrainOpts(C.Structure):
    _fields_ = [("train_scaling", C.c_int), ("train_transitions", C.c_int), ("train_drift", C.c_int)]


# every symbol include/nanocall_b200.h declares: (restype, argtypes)
_vp, _u32, _i32, _f = C.c_void_p, C.c_uint32, C.c_int32, C.c_float
SYMBOLS = {
    "nc_ctx_create": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(_vp)]),
    "nc_ctx_destroy": (None, [_vp]),
    "nc_last_error": (C.c_char_p, [_vp]),
    "nc_ctx_stream": (_vp, [_vp]),
    "nc_ctx_sync": (C.c_int, [_vp]),
    "nc_ctx_last_kernel_ms": (C.c_float, [_vp]),
    "nc_ctx_last_launches": (C.c_int, [_vp]),
    "nc_ctx_viterbi_stats": (C.c_int, [_vp, _vp, C.c_int]),
    "nc_ctx_set_viterbi_mode": (C.c_int, [_vp, C.c_int]),
    "nc_ctx_device_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.c_char_p, C.c_int]),
    "nc_model_register": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_int)]),
    "nc_model_stats": (C.c_int, [_vp, C.c_int, C.POINTER(_f), C.POINTER(_f)]),
    "nc_viterbi_packed": (C.c_int, [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "nc_viterbi_batch": (C.c_int, [_vp, _u32, _vp, _vp]),
    "nc_base_seq": (_u32, [_u32, _vp, _vp, _vp, _u32]),
    "nc_fwbw": (C.c_int, [_vp, _i32, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, C.POINTER(_f)]),
    "nc_train_round_batch": (C.c_int, [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nc_ctx_train_stats": (C.c_int, [_vp, _vp, C.c_int]),
    "nc_ctx_set_default_transitions": (C.c_int, [_vp, _f, _f, _u32, _vp, _vp, _vp]),
    "nc_ctx_reserve": (C.c_int, [_vp, C.c_uint64, C.c_uint64]),
    "nc_host_alloc": (_vp, [C.c_size_t]),
    "nc_host_free": (None, [_vp]),
    "nc_mean_stdv": (None, [_u32, _vp, C.POINTER(_f), C.POINTER(_f)]),
    "nc_transition_lut": (None, [_f, _f, _vp]),
    "nc_min_skip": (_u32, [_u32, _u32]),
    "nc_plan_dispatch_order": (C.c_int, [_u32, _vp, C.c_uint64, _u32, _vp]),
    "nc_version": (C.c_char_p, []),
}

_lib = None


def build():
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C nanocall_b200/csrc` "
            "(or __graft_entry__.build()); nanocall_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
