"""ctypes binding of libwann_b200.so (C ABI in include/wann_b200.h).

There is no CPU fallback: if the library is missing or no sm_100 GPU is present the
import / wann_create fails loudly."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwann_b200.so")

MODE_BF16, MODE_FP32, MODE_FP16, MODE_FP32_TC, MODE_FP16C = 0, 1, 2, 3, 4
MODES = {"bf16": MODE_BF16, "fp32": MODE_FP32, "fp16": MODE_FP16, "fp32_tc": MODE_FP32_TC, "fp16c": MODE_FP16C}
MEM_HOST, MEM_DEVICE = 0, 1
PLANES_I8, PLANES_F32 = 0, 1
N_MOVES, MAX_LEGAL = 155, 48

EXPORTS = [
    "wann_create", "wann_set_weights", "wann_destroy", "wann_last_error", "wann_encode", "wann_legal_mask",
    "wann_eval_probs", "wann_eval_planes", "wann_eval_topk", "wann_eval_expand", "wann_eval_expand_seeded", "wann_expand_ranked", "wann_tree_step", "wann_debug_activations",
    "wann_debug_profile", "wann_debug_centering", "wann_debug_effective_weights", "wann_gptq_round", "wann_kernel_times", "wann_crc32c",
    "wann_host_alloc", "wann_host_free", "wann_get_stats", "wann_set_option", "wann_version",
    "wann_forest_create", "wann_forest_create_device", "wann_forest_run", "wann_forest_run_steps", "wann_forest_debug_cycles", "wann_forest_destroy", "wann_forest_next", "wann_forest_apply", "wann_forest_info",
    "wann_forest_dump", "wann_forest_advance", "wann_forest_play", "wann_host_threads",
]


class WannWeights(ctypes.Structure):
    _fields_ = [("conv_w", ctypes.c_void_p * 5), ("conv_b", ctypes.c_void_p * 5),
                ("fc_w", ctypes.c_void_p), ("fc_b", ctypes.c_void_p),
                ("final_kernel", ctypes.c_int)]


class WannError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (built by `python __graft_entry__.py` / wann_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WannError("%s not found: build it with `python -m wann_b200.build` "
                        "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.wann_last_error.restype = ctypes.c_char_p
    lib.wann_create.argtypes = [ctypes.POINTER(WannWeights), ci, ci, ctypes.POINTER(vp)]
    lib.wann_destroy.argtypes = [vp]
    lib.wann_encode.argtypes = [vp, vp, vp, i64, vp, ci, vp]
    lib.wann_legal_mask.argtypes = [vp, vp, vp, i64, vp, ci, vp]
    lib.wann_eval_probs.argtypes = [vp, vp, vp, i64, vp, vp, ci, vp]
    lib.wann_eval_planes.argtypes = [vp, vp, ci, i64, vp, vp, ci, vp]
    lib.wann_eval_topk.argtypes = [vp, vp, vp, i64, ci, vp, vp, vp, ci, vp]
    lib.wann_eval_expand.argtypes = [vp, vp, vp, i64, ci, vp, vp, vp, vp, vp, ci, vp]
    lib.wann_eval_expand_seeded.argtypes = [vp, vp, vp, i64, ci, vp, vp, vp, vp, vp, vp, vp, ci, vp]
    lib.wann_expand_ranked.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.wann_host_threads.argtypes = [ci]
    lib.wann_forest_create.argtypes = [i64, vp, vp, vp, ctypes.POINTER(vp)]
    lib.wann_forest_create_device.argtypes = [vp, i64, vp, vp, vp, ctypes.c_int32, ctypes.POINTER(vp)]
    lib.wann_forest_run.argtypes = [vp, vp, i64, vp]
    lib.wann_forest_run_steps.argtypes = [vp, vp, i64, i64, vp]
    lib.wann_forest_debug_cycles.argtypes = [vp, vp]
    lib.wann_forest_destroy.argtypes = [vp]
    lib.wann_forest_next.argtypes = [vp, i64, vp, vp, vp, ctypes.POINTER(i64)]
    lib.wann_forest_apply.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.wann_forest_info.argtypes = [vp, i64, vp]
    lib.wann_forest_advance.argtypes = [vp, i64, ci]
    lib.wann_forest_play.argtypes = [vp, vp, ci, vp, vp]
    lib.wann_forest_dump.argtypes = [vp, i64, vp, i64]
    lib.wann_forest_dump.restype = i64
    lib.wann_set_weights.argtypes = [vp, ctypes.POINTER(WannWeights)]
    lib.wann_tree_step.argtypes = [vp, vp, vp, i64, ci, ctypes.c_uint64, ctypes.c_uint64, vp, vp, vp]
    lib.wann_debug_activations.argtypes = [vp, vp, vp, i64, ci, vp, ci]
    lib.wann_debug_profile.argtypes = [vp, vp, vp, i64, vp]
    lib.wann_debug_centering.argtypes = [vp, vp, vp, vp]
    lib.wann_debug_effective_weights.argtypes = [vp, ci, vp]
    lib.wann_gptq_round.argtypes = [vp, ci, vp, ci, vp, ctypes.c_double]
    lib.wann_kernel_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double * 2)]
    lib.wann_crc32c.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint32]
    lib.wann_crc32c.restype = ctypes.c_uint32
    lib.wann_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_uint64]
    lib.wann_host_free.argtypes = [vp]
    lib.wann_get_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64 * 8)]
    lib.wann_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    for name in EXPORTS:
        getattr(lib, name)  # AttributeError if a declared symbol is missing
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise WannError("wann_b200 error %d: %s" % (rc, load().wann_last_error().decode()))
