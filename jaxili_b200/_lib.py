"""ctypes binding of the C ABI declared in include/mafb200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` /
``python -m jaxili_b200.build`` (nvcc, sm_100a).  There is no CPU fallback: if the
library is missing, or no CUDA device is present, every compute call raises.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmafb200.so")

MAX_HIDDEN = 5

ACTIVATIONS = {
    "relu": 0, "silu": 1, "swish": 1, "tanh": 2, "gelu": 3, "elu": 4,
    "sigmoid": 5, "softplus": 6, "leaky_relu": 7,
}

PACKED_CURRENT = 1
NORM_CURRENT = 2
DP_SLOT = 4
FLOW_ONLY = 8


class MafDesc(C.Structure):
    _fields_ = [
        ("n_in", C.c_int32), ("n_cond", C.c_int32), ("n_layers", C.c_int32),
        ("n_hidden", C.c_int32), ("hidden", C.c_int32 * MAX_HIDDEN),
        ("activation", C.c_int32), ("use_reverse", C.c_int32),
    ]


class MafOptDesc(C.Structure):
    _fields_ = [
        ("lr", C.c_float), ("warmup", C.c_float), ("decay_steps", C.c_float),
        ("init_factor", C.c_float), ("end_factor", C.c_float),
        ("b1", C.c_float), ("b2", C.c_float), ("eps", C.c_float),
        ("clip", C.c_float), ("weight_decay", C.c_float), ("kind", C.c_int32),
    ]


class MafHmcState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("eps", "inv_mass", "prior_kind", "lo", "hi", "z", "p", "g", "logdens", "theta", "scratch")]


# name -> (restype, argtypes); every symbol include/mafb200.h declares
SYMBOLS = {
    "mafb200_last_error": (C.c_char_p, []),
    "mafb200_version": (C.c_int, []),
    "mafb200_create": (C.c_int, [C.POINTER(MafDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "mafb200_destroy": (C.c_int, [C.c_void_p]),
    "mafb200_set_standardization": (C.c_int, [C.c_void_p] * 5),
    "mafb200_debug_trace": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mafb200_flow_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "mafb200_flow_elapsed": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mafb200_dp_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "mafb200_dp_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mafb200_dp_bind_step": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mafb200_dp_reset": (C.c_int, [C.c_void_p]),
    "mafb200_dp_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mafb200_dp_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "mafb200_debug_gemm": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mafb200_set_wide_core": (C.c_int, [C.c_void_p, C.c_int32]),
    "mafb200_set_flow_kernel": (C.c_int, [C.c_void_p, C.c_int32]),
    "mafb200_eval_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_int32, C.c_void_p]),
    "mafb200_param_count": (C.c_int64, [C.c_void_p]),
    "mafb200_get_mask": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mafb200_query": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int32)] * 4),
    "mafb200_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mafb200_log_prob": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_int32, C.c_void_p]),
    "mafb200_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mafb200_inverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mafb200_loss_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_float, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_void_p]),
    "mafb200_log_prob_grad_y": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mafb200_hmc_trajectory": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(MafHmcState),
                                         C.c_int32, C.c_int32, C.c_void_p]),
    "mafb200_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.POINTER(MafOptDesc), C.c_int32, C.c_void_p]),
    "mafb200_pipeline_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.POINTER(MafOptDesc), C.c_float, C.c_int32, C.c_void_p,
                                          C.POINTER(C.c_void_p)]),
    "mafb200_pipeline_capture": (C.c_int, [C.c_void_p]),
    "mafb200_pipeline_fence": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mafb200_pipeline_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "mafb200_pipeline_drain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]),
    "mafb200_pipeline_destroy": (C.c_int, [C.c_void_p]),
    "mafb200_clip_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(MafOptDesc), C.c_int32, C.c_void_p]),
    "mafb200_sample_from_noise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                            C.c_int32, C.c_void_p]),
    "mafb200_sample_philox": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p,
                                        C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int32,
                                        C.c_void_p]),
    "mafb200_ffma_peak": (C.c_int, [C.c_int, C.c_int32, C.POINTER(C.c_float)]),
    "mafb200_fp32_peak": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "mafb200_tf32_peak": (C.c_int, [C.c_int, C.c_int32, C.POINTER(C.c_float)]),
}

_lib = None


class MafB200Error(RuntimeError):
    """Raised when the native library reports an error (no CPU fallback exists)."""


def load():
    """dlopen the in-tree library and bind every declared symbol; fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise MafB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). jaxili_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise MafB200Error(load().mafb200_last_error().decode())
