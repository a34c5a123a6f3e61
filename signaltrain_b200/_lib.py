"""ctypes binding of include/signaltrain_b200.h.  There is no CPU or PyTorch fallback: if the library is
missing or no sm_100 device is present, every entry point raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsignaltrain_b200.so")
NUM_PARAMS = 40
NUM_ACTS = 30

c_float_p = ctypes.c_void_p          # raw device addresses travel as void*
PtrTable = ctypes.c_void_p * NUM_PARAMS
ActTable = ctypes.c_void_p * NUM_ACTS


class StConfig(ctypes.Structure):
    _fields_ = [("chunk", ctypes.c_int), ("ft", ctypes.c_int), ("hop", ctypes.c_int), ("frames_in", ctypes.c_int),
                ("frames_out", ctypes.c_int), ("knobs", ctypes.c_int), ("rank", ctypes.c_int)]


class StAdam(ctypes.Structure):
    _fields_ = [("lr", ctypes.c_float), ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float),
                ("step", ctypes.c_int), ("grad_scale", ctypes.c_float), ("max_norm", ctypes.c_float)]


_SIGNATURES = {
    "st_abi_version": (ctypes.c_int, []),
    "st_create": (ctypes.c_int, [ctypes.POINTER(StConfig), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "st_destroy": (None, [ctypes.c_void_p]),
    "st_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "st_param_name": (ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_int]),
    "st_param_numel": (ctypes.c_long, [ctypes.c_void_p, ctypes.c_int]),
    "st_out_samples": (ctypes.c_int, [ctypes.c_void_p]),
    "st_bins": (ctypes.c_int, [ctypes.c_void_p]),
    "st_init_frontend": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "st_analysis": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, c_float_p, c_float_p, ctypes.c_void_p]),
    "st_synthesis": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, c_float_p,
                                    ctypes.c_void_p]),
    "st_dct_analysis": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "st_dct_synthesis": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, c_float_p, ctypes.c_void_p]),
    "st_compressor_4c": (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_float_p,
                                        ctypes.c_void_p]),
    "st_crop_windows": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_long, ctypes.c_void_p, c_float_p, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, c_float_p, c_float_p, ctypes.c_void_p]),
    "st_forward": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_void_p, c_float_p, c_float_p,
                                  c_float_p, ctypes.c_void_p, ctypes.c_void_p]),
    "st_loss": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_float, ctypes.c_int,
                               c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    "st_loss_shaped": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, c_float_p, ctypes.c_float, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p, c_float_p, c_float_p, ctypes.c_void_p]),
    "st_mae": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, ctypes.c_long, c_float_p, ctypes.c_void_p]),
    "st_backward": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]),
    "st_backward_begin": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]),
    "st_backward_finish": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]),
    "st_clip_grad_norm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, c_float_p, ctypes.c_void_p]),
    "st_adam_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.POINTER(StAdam), ctypes.c_void_p]),
    "st_train_step": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, c_float_p, ctypes.c_float,
                                     ctypes.POINTER(StAdam), c_float_p, ctypes.c_void_p]),
    "st_grad_step": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, c_float_p, ctypes.c_float, c_float_p, ctypes.c_void_p]),
    "st_set_training": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "st_set_precision": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "st_get_precision": (ctypes.c_int, [ctypes.c_void_p]),
    "st_launch_count": (ctypes.c_long, [ctypes.c_void_p]),
    "st_profile_stage_count": (ctypes.c_int, []),
    "st_profile_stage_name": (ctypes.c_char_p, [ctypes.c_int]),
    "st_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "st_profile_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_long)]),
    "st_packed_grad_floats": (ctypes.c_long, [ctypes.c_void_p]),
    "st_pack_grads": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_float_p, ctypes.c_void_p]),
    "st_unpack_grads": (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_void_p, ctypes.c_void_p]),
    "st_grad_step_packed": (ctypes.c_int, [ctypes.c_void_p, c_float_p, c_float_p, c_float_p, ctypes.c_int, ctypes.c_void_p,
                                           c_float_p, c_float_p, ctypes.c_float, c_float_p, ctypes.c_void_p]),
    "st_unpack_clip": (ctypes.c_int, [ctypes.c_void_p, c_float_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float, c_float_p,
                                      ctypes.c_void_p]),
    "st_adam_step_clipped": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.POINTER(StAdam), ctypes.c_void_p]),
    "st_debug_fallbacks": (ctypes.c_long, [ctypes.c_void_p]),
    "st_debug_graph_replays": (ctypes.c_long, [ctypes.c_void_p]),
    "st_debug_ae_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]),
    "st_debug_gemm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p, c_float_p, ctypes.c_long,
                                     c_float_p, c_float_p, ctypes.c_long, c_float_p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "st_debug_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_long]),
    "st_debug_numel": (ctypes.c_long, [ctypes.c_void_p, ctypes.c_char_p]),
    "st_debug_gemm_plan": (ctypes.c_int, [ctypes.POINTER(StConfig), ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """dlopen the in-tree library (built by `make` / __graft_entry__.build()) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"signaltrain_b200: {LIB_PATH} not built (run `make` or __graft_entry__.build()); "
                           "there is no CPU/PyTorch fallback for this path")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.st_abi_version() != 1:
        raise RuntimeError("signaltrain_b200: ABI version mismatch between _lib.py and the shared library")
    _lib = lib
    return lib
