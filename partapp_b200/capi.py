"""ctypes binding of libpsinfer.so -- the C ABI declared in include/psinfer.h.

This is the only way Python reaches the kernels; there is no Python/CPU implementation of the path behind it.
Importing works without a GPU (so the symbol table can be checked on a CPU box); creating a context does not.
"""
import ctypes as C
import os

PS_MAX_PARTS = 64
PS_HYP_VEC = 7

PS_OK, PS_ERR_INVALID, PS_ERR_CUDA, PS_ERR_STATE, PS_ERR_UNSUPPORTED = range(5)
PS_MEM_HOST, PS_MEM_DEVICE = 0, 1
PS_JOINT_POS_GAUSSIAN, PS_JOINT_ROT_GAUSSIAN = 1, 2
PS_INFER_SPARSE, PS_INFER_LOCAL_MAX, PS_INFER_ROOT_HYPS, PS_INFER_KEEP_UNARIES, PS_INFER_NO_BORDER_STRIP = 1, 2, 4, 8, 16

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpsinfer.so")


class ps_config(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("num_parts", C.c_int),
        ("num_rotation_steps", C.c_int),
        ("min_part_rotation", C.c_float),
        ("max_part_rotation", C.c_float),
        ("num_scale_steps", C.c_int),
        ("min_object_scale", C.c_float),
        ("max_object_scale", C.c_float),
        ("height", C.c_int),
        ("width", C.c_int),
        ("root_idx", C.c_int),
        ("is_detect", C.c_ubyte * PS_MAX_PARTS),
        ("is_upright", C.c_ubyte * PS_MAX_PARTS),
        ("is_root", C.c_ubyte * PS_MAX_PARTS),
        ("strip_border_detections", C.c_float),
        ("roi_save_num_samples", C.c_int),
        ("keep_all_scales", C.c_int),
        ("interpolate", C.c_int),
        ("fast_math", C.c_int),
    ]


class ps_joint(C.Structure):
    _fields_ = [
        ("type", C.c_int),
        ("child_idx", C.c_int),
        ("parent_idx", C.c_int),
        ("offset_c", C.c_double * 2),
        ("offset_p", C.c_double * 2),
        ("C", C.c_double * 4),
        ("rot_mean", C.c_double),
        ("rot_sigma", C.c_double),
    ]


_ctx_p = C.c_void_p
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol include/psinfer.h declares
PROTOTYPES = {
    "ps_create": (C.c_int, [C.POINTER(ps_config), C.POINTER(_ctx_p)]),
    "ps_destroy": (None, [_ctx_p]),
    "ps_last_error": (C.c_char_p, [_ctx_p]),
    "ps_set_stream": (C.c_int, [_ctx_p, C.c_void_p]),
    "ps_synchronize": (C.c_int, [_ctx_p]),
    "ps_set_joints": (C.c_int, [_ctx_p, C.POINTER(ps_joint), C.c_int]),
    "ps_flip_joint": (None, [C.POINTER(ps_joint)]),
    "ps_rot_from_index": (C.c_double, [C.POINTER(ps_config), C.c_int]),
    "ps_scale_from_index": (C.c_double, [C.POINTER(ps_config), C.c_int]),
    "ps_index_from_rot": (C.c_int, [C.POINTER(ps_config), C.c_double]),
    "ps_set_unary": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "ps_set_unary_compact": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, _dp, C.c_int]),
    "ps_set_unary_compact_raw": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, _dp, C.c_int]),
    "ps_log_unary": (C.c_int, [_ctx_p, C.c_int, C.c_int]),
    "ps_unary_local_max": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.c_int, _fp, _ip]),
    "ps_set_unaries_compact": (C.c_int, [_ctx_p, C.c_int, _ip, _ip, C.POINTER(C.c_void_p), C.c_int, C.c_int, _dp, C.c_int]),
    "ps_get_unary": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "ps_add_unary_table": (C.c_int, [_ctx_p, C.c_int, _fp, C.c_int, C.c_float]),
    "ps_add_unary_tables": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), _ip, _fp, C.c_int]),
    "ps_add_unary_grid": (C.c_int, [_ctx_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int]),
    "ps_rot_score_table": (None, [C.POINTER(ps_config), C.c_double, C.c_double, _fp]),
    "ps_pos_score_table": (None, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.c_double, _fp]),
    "ps_torso_prior_table": (None, [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_float, _fp]),
    "ps_infer": (C.c_int, [_ctx_p, C.c_int]),
    "ps_max_states": (C.c_int, [_ctx_p, C.c_int]),
    "ps_get_best_conf": (C.c_int, [_ctx_p, _fp]),
    "ps_get_part_hyps": (C.c_int, [_ctx_p, C.c_int, _fp, C.c_int, _ip]),
    "ps_get_marginal": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "ps_get_root_posterior": (C.c_int, [_ctx_p, C.c_void_p, C.c_int]),
    "ps_get_root_hyps": (C.c_int, [_ctx_p, _fp, C.c_int, _ip]),
    "ps_message": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double,
                             C.c_double, C.c_int]),
    "ps_pos_message": (C.c_int, [_ctx_p, C.c_void_p, C.c_void_p, C.c_int, _dp, _dp, C.c_double, C.c_int]),
    "ps_find_local_max": (C.c_int, [_ctx_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _ip]),
    "ps_get_plan_info": (C.c_int, [_ctx_p, C.c_int, C.c_int, C.c_int, _ip]),
    "ps_plan_work_lists": (C.c_int, [C.POINTER(ps_config), _dp, C.c_double, _ip, _dp, _ip, _ip, C.c_int]),
    "ps_plan_walks": (C.c_int, [C.POINTER(ps_config), _dp, C.c_double, _ip, _ip, C.c_int, C.POINTER(C.c_ubyte), C.c_int, _ip]),
    "ps_selftest_math": (C.c_int, [_ctx_p, C.c_uint, C.c_ulonglong, C.POINTER(C.c_ulonglong)]),
    "ps_eval_math": (C.c_int, [_ctx_p, C.c_int, C.c_uint, C.c_uint, _fp]),
    "ps_launch_count": (C.c_longlong, [_ctx_p]),
    "ps_profile_enable": (C.c_int, [_ctx_p, C.c_int]),
    "ps_profile_read": (C.c_int, [_ctx_p, C.c_int, C.POINTER(C.c_char_p), _dp, C.POINTER(C.c_longlong), _ip]),
    "ps_version": (C.c_char_p, []),
}

_lib = None


class PsInferError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("psinfer status %d: %s" % (status, message))
        self.status = status


def load_library(path=None):
    """Loads libpsinfer.so and types its entry points.  Fails loudly if the CUDA extension was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            "%s is missing: build the CUDA extension first (python -m partapp_b200.build). "
            "partapp_b200 has no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
