"""ctypes wrapper of the CPU oracle (oracle/ps_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, never by
partapp_b200.  Parity status: unpinned (the reference ships no golden vectors for this path; see ps_oracle.cpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libps_oracle.so")
SRC = os.path.join(HERE, "ps_oracle.cpp")


class orc_exp_param(C.Structure):
    _fields_ = [("num_rotation_steps", C.c_int), ("min_part_rotation", C.c_float), ("max_part_rotation", C.c_float),
                ("num_scale_steps", C.c_int), ("min_object_scale", C.c_float), ("max_object_scale", C.c_float),
                ("strip_border_detections", C.c_float), ("roi_save_num_samples", C.c_int)]


class orc_joint(C.Structure):
    _fields_ = [("type", C.c_int), ("child_idx", C.c_int), ("parent_idx", C.c_int), ("offset_c", C.c_double * 2),
                ("offset_p", C.c_double * 2), ("C", C.c_double * 4), ("rot_mean", C.c_double), ("rot_sigma", C.c_double)]


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return LIB


_lib = None
_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_rot_from_index.restype = C.c_double
        L.orc_scale_from_index.restype = C.c_double
        L.orc_rot_from_index.argtypes = [C.POINTER(orc_exp_param), C.c_int]
        L.orc_scale_from_index.argtypes = [C.POINTER(orc_exp_param), C.c_int]
        L.orc_index_from_rot.argtypes = [C.POINTER(orc_exp_param), C.c_double]
        L.orc_gaussian_filter.argtypes = [C.c_double, _dp, C.c_int]
        L.orc_eig2d.argtypes = [_dp, _dp, _dp]
        L.orc_gauss_filter_2d.argtypes = [_fp, _fp, C.c_int, C.c_int, _dp, C.c_int]
        L.orc_enlarged_size.argtypes = [C.c_int, C.c_int, _dp, _ip, _ip]
        L.orc_transform_fixed.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _dp, C.c_float, C.c_int]
        L.orc_compare_math.restype = C.c_ulonglong
        L.orc_compare_math.argtypes = [C.c_int, C.c_uint, C.c_uint, _fp]
        L.orc_load_score_grid.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _fp]
        L.orc_message.argtypes = [C.POINTER(orc_exp_param), _fp, _fp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp,
                                  C.c_double, C.c_double, C.c_double, C.c_int, _fp, _fp, _fp]
        L.orc_prepare_unary.argtypes = [_fp, C.c_size_t]
        L.orc_gauss_filter_2d_offset.argtypes = [_fp, _fp, C.c_int, C.c_int, _dp, _dp, C.c_int]
        L.orc_filter_1d_wraparound.argtypes = [_fp, _fp, C.c_int, _fp, C.c_int]
        L.orc_pointwise.argtypes = [C.c_int, _fp, C.c_size_t]
        L.orc_hc_inverse.argtypes = [_dp, _dp]
        L.orc_transformed_bbox.argtypes = [_dp, C.c_int, C.c_int, _dp]
        L.orc_pos_message.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_int]
        L.orc_flip_joint.argtypes = [C.POINTER(orc_joint)]
        L.orc_rot_score_table.argtypes = [C.POINTER(orc_exp_param), C.c_double, C.c_double, _fp]
        L.orc_pos_score_table.argtypes = [C.c_int, C.c_int] + [C.c_double] * 6 + [_fp]
        L.orc_torso_prior_table.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4 + [C.c_float, _fp]
        L.orc_add_rot_table.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp, C.c_float]
        L.orc_add_pos_table.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp, C.c_float]
        L.orc_add_pos_table_unweighted.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp]
        L.orc_add_dpm_score.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_float]
        L.orc_add_load_dpm_score.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_float]
        L.orc_find_local_max.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]
        L.orc_argmax.argtypes = [_fp, C.c_int, _fp]
        L.orc_infer.argtypes = [C.POINTER(orc_exp_param), C.c_int, _ip, _ip, C.c_int, C.POINTER(orc_joint), C.c_int,
                                C.c_int, C.c_int, _fp, C.c_int, _fp, _fp, _fp, _fp, C.c_int, _ip]
        L.orc_get_max_states.argtypes = [C.POINTER(orc_exp_param), C.c_int, C.c_int, C.c_int, _fp, _fp]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(_fp)


def _d(seq):
    arr = np.ascontiguousarray(np.asarray(seq, np.float64).reshape(-1))
    return arr, arr.ctypes.data_as(_dp)


def exp_param(ep):
    """ep: any object with the ExpParam field names (partapp_b200.ExpParam or a test double)."""
    return orc_exp_param(int(ep.num_rotation_steps), float(ep.min_part_rotation), float(ep.max_part_rotation),
                         int(ep.num_scale_steps), float(ep.min_object_scale), float(ep.max_object_scale),
                         float(ep.strip_border_detections), int(ep.roi_save_num_samples))


def joint(j):
    o = orc_joint()
    o.type = int(j.type)
    o.child_idx, o.parent_idx = int(j.child_idx), int(j.parent_idx)
    o.offset_c[0], o.offset_c[1] = float(j.offset_c[0]), float(j.offset_c[1])
    o.offset_p[0], o.offset_p[1] = float(j.offset_p[0]), float(j.offset_p[1])
    cm = np.asarray(j.C, np.float64).reshape(4)
    for i in range(4):
        o.C[i] = float(cm[i])
    o.rot_mean, o.rot_sigma = float(j.rot_mean), float(j.rot_sigma)
    return o


def message(ep, child, off_in, off_out, Cm, rot_mean, rot_sigma, scale, sparse, debug=False):
    """computeRotJointMarginal (reference objectdetect_findrot.cpp:292-456)."""
    child = np.ascontiguousarray(child, np.float32)
    R, H, W = child.shape
    out = np.empty_like(child)
    e = exp_param(ep)
    _a, pa = _d(off_in)
    _b, pb = _d(off_out)
    _c, pc = _d(Cm)
    dbg = [np.empty_like(child) for _ in range(3)] if debug else [None] * 3
    lib().orc_message(C.byref(e), _f(child), _f(out), R, H, W, pa, pb, pc, float(rot_mean), float(rot_sigma),
                      float(scale), int(bool(sparse)), *[(_f(d) if d is not None else None) for d in dbg])
    return (out, dbg) if debug else out


def pos_message(child, offset, Cm, scale, sparse):
    """computePosJointMarginal (reference objectdetect_findpos.cpp:64-89) on each [H][W] slice of child [D][H][W].
    Returns (log_prob_parent, log_prob_child as the reference leaves it: log(exp(child)))."""
    child = np.ascontiguousarray(child, np.float32).copy()
    D, H, W = child.shape
    out = np.empty_like(child)
    _a, pa = _d(offset)
    _c, pc = _d(Cm)
    lib().orc_pos_message(_f(child), _f(out), D, H, W, pa, pc, float(scale), int(bool(sparse)))
    return out, child


def load_score_grid(cells, Tig, H, W, interpolate=False):
    """PartApp::loadScoreGrid mapping (reference libPartApp/partapp.cpp:874-896): cells [R][gh][gw], Tig [R][3][3]."""
    cells = np.ascontiguousarray(cells, np.float32)
    R, gh, gw = cells.shape
    t = np.ascontiguousarray(Tig, np.float64).reshape(R * 9)
    out = np.empty((R, H, W), np.float32)
    lib().orc_load_score_grid(_f(cells), R, gh, gw, t.ctypes.data_as(_dp), H, W, int(bool(interpolate)), _f(out))
    return out


def prepare_unary(raw):
    g = np.ascontiguousarray(raw, np.float32).copy()
    lib().orc_prepare_unary(_f(g), g.size)
    return g


def find_local_max(grid, max_n):
    g = np.ascontiguousarray(grid, np.float32)
    d0, h, w = g.shape
    out = np.empty((max(max_n, 1), 4), np.float32)
    n = lib().orc_find_local_max(_f(g), d0, h, w, int(max_n), _f(out))
    return out[:n].copy()


def argmax(grid):
    g = np.ascontiguousarray(grid, np.float32)
    val = C.c_float()
    idx = lib().orc_argmax(_f(g), g.size, C.byref(val))
    return idx, val.value


def gauss_filter_2d(grid, Cm, sparse):
    g = np.ascontiguousarray(grid, np.float32)
    out = np.empty_like(g)
    _c, pc = _d(Cm)
    lib().orc_gauss_filter_2d(_f(g), _f(out), g.shape[0], g.shape[1], pc, int(bool(sparse)))
    return out


def infer(ep, part_conf, joints, unaries, sparse=True, want_marginals=True, want_hyps=False):
    """computeRootPosteriorRot (reference objectdetect_findrot.cpp:470-727).
    unaries: [P][S][R][H][W] float32, masked in place like the reference does.  Returns a dict."""
    assert unaries.dtype == np.float32 and unaries.flags.c_contiguous
    P, S, R, H, W = unaries.shape
    e = exp_param(ep)
    det = (C.c_int * P)(*[int(bool(v)) for v in part_conf.is_detect])
    upr = (C.c_int * P)(*[int(bool(v)) for v in part_conf.is_upright])
    roots = [p for p in range(P) if part_conf.is_detect[p] and part_conf.is_root[p]]
    assert len(roots) == 1
    js = (orc_joint * len(joints))(*[joint(j) for j in joints])
    marg = np.empty((S, P, R, H, W), np.float32) if want_marginals else None
    rootp = np.empty((S, H, W), np.float32)
    best = np.empty((P, 7), np.float32)
    cap = int(ep.roi_save_num_samples) + 1
    hyps = np.empty((P, cap, 7), np.float32) if want_hyps else None
    nh = (C.c_int * P)()
    rc = lib().orc_infer(C.byref(e), P, det, upr, roots[0], js, len(joints), H, W, _f(unaries), int(bool(sparse)),
                         _f(marg) if marg is not None else None, _f(rootp), _f(best),
                         _f(hyps) if hyps is not None else None, cap, nh)
    if rc != 0:
        raise RuntimeError("orc_infer failed: %d" % rc)
    out = {"root_post": rootp, "best_conf": best, "marginals": marg}
    if want_hyps:
        out["part_hyps"] = [hyps[p, :nh[p]].copy() for p in range(P)]
    return out
