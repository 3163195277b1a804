"""ctypes access to oracle/_ref/libps_ref_core.so: the reference's OWN libMultiArray / libBoostMath code, compiled here
from /root/reference (oracle/ref_core.cpp, `make -C oracle ref`).  TEST INFRASTRUCTURE: used by
tests/test_oracle_vs_ref.py and tests/golden/make_ref_golden.py to pin the oracle's restatement; never by the product."""
import ctypes as C
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libps_ref_core.so")
_fp, _dp, _ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        L.ref_gaussian_filter.argtypes = [C.c_double, _dp, C.c_int]
        L.ref_eig2d.argtypes = [_dp, _dp, _dp]
        L.ref_hc_inverse.argtypes = [_dp, _dp]
        L.ref_hc_compose.argtypes = [C.c_int, C.c_double, C.c_double, _dp]
        L.ref_prod3.argtypes = [_dp, _dp, _dp]
        L.ref_map_point.argtypes = [_dp, C.c_double, C.c_double, _dp, _dp]
        L.ref_transformed_bbox.argtypes = [_dp, C.c_int, C.c_int, _dp]
        L.ref_transform_fixed.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _dp, C.c_float, C.c_int]
        L.ref_transform_resize.argtypes = [_fp, C.c_int, C.c_int, _dp, C.c_float, C.c_int, _fp, C.c_long, _ip, _ip, _dp]
        L.ref_transform_resize.restype = C.c_long
        L.ref_gauss_filter_diag2d.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_double, C.c_double]
        L.ref_gauss_filter_2d.argtypes = [_fp, _fp, C.c_int, C.c_int, _dp, C.c_int]
        L.ref_gauss_filter_2d_offset.argtypes = [_fp, _fp, C.c_int, C.c_int, _dp, _dp, C.c_int]
        L.ref_filter_1d_wraparound.argtypes = [_fp, _fp, C.c_int, _fp, C.c_int]
        L.ref_pointwise.argtypes = [C.c_int, _fp, _fp, C.c_float, C.c_long]
        L.ref_min_max.argtypes = [_fp, C.c_long, _fp, _fp]
        L.ref_bins.argtypes = [C.c_int, C.c_float, C.c_float, C.c_uint, C.c_int, C.c_double]
        L.ref_bins.restype = C.c_double
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(_fp)


def _d(a):
    a = np.ascontiguousarray(a, np.float64)
    return a, a.ctypes.data_as(_dp)


def gaussian_filter(sigma):
    out = np.empty(4096, np.float64)
    n = lib().ref_gaussian_filter(float(sigma), out.ctypes.data_as(_dp), out.size)
    assert n > 0
    return out[:n].copy()


def eig2d(Cm):
    _c, pc = _d(Cm)
    V, E = np.empty(4), np.empty(4)
    lib().ref_eig2d(pc, V.ctypes.data_as(_dp), E.ctypes.data_as(_dp))
    return V.reshape(2, 2), E.reshape(2, 2)


def hc_inverse(T):
    _t, pt = _d(T)
    out = np.empty(9)
    lib().ref_hc_inverse(pt, out.ctypes.data_as(_dp))
    return out.reshape(3, 3)


def transformed_bbox(T, w, h):
    _t, pt = _d(T)
    out = np.empty(4)
    lib().ref_transformed_bbox(pt, int(w), int(h), out.ctypes.data_as(_dp))
    return out


def transform_fixed(grid, out_shape, T, default_value, method):
    g = np.ascontiguousarray(grid, np.float32)
    out = np.empty(out_shape, np.float32)
    _t, pt = _d(T)
    lib().ref_transform_fixed(_f(g), g.shape[0], g.shape[1], _f(out), out.shape[0], out.shape[1], pt, float(default_value),
                              int(method))
    return out


def gauss_filter_2d(grid, Cm, sparse):
    g = np.ascontiguousarray(grid, np.float32)
    out = np.empty_like(g)
    _c, pc = _d(Cm)
    lib().ref_gauss_filter_2d(_f(g), _f(out), g.shape[0], g.shape[1], pc, int(bool(sparse)))
    return out


def gauss_filter_2d_offset(grid, Cm, offset, sparse):
    g = np.ascontiguousarray(grid, np.float32)
    out = np.empty_like(g)
    _c, pc = _d(Cm)
    _o, po = _d(offset)
    lib().ref_gauss_filter_2d_offset(_f(g), _f(out), g.shape[0], g.shape[1], pc, po, int(bool(sparse)))
    return out


def filter_1d_wraparound(col, taps):
    c = np.ascontiguousarray(col, np.float32)
    f = np.ascontiguousarray(taps, np.float32)
    out = np.empty_like(c)
    lib().ref_filter_1d_wraparound(_f(c), _f(out), c.size, _f(f), f.size)
    return out


def pointwise(op, a):
    a = np.ascontiguousarray(a, np.float32).copy()
    lib().ref_pointwise(int(op), _f(a), None, 0.0, a.size)
    return a


def bins(mn, mx, n):
    """(rot_from_index(i) for all i, index_from_rot of each of those centres)."""
    L = lib()
    centres = np.array([L.ref_bins(0, mn, mx, n, i, 0.0) for i in range(n)])
    back = np.array([L.ref_bins(2, mn, mx, n, 0, float(c)) for c in centres])
    return np.concatenate([centres, back])


# ---- the reference's own inference drivers (objectdetect_findrot.cpp compiled unmodified, oracle/ref_drivers.cpp) ------
DRIVERS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libps_ref_drivers.so")
_dlib = None


def drivers_available():
    return os.path.exists(DRIVERS_PATH)


def dlib():
    global _dlib
    if _dlib is None:
        L = C.CDLL(DRIVERS_PATH)
        L.refd_message.argtypes = [_dp, _fp, _fp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double,
                                   C.c_double, C.c_int]
        L.refd_infer.argtypes = [_dp, C.c_int, _ip, _ip, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, _fp, _fp,
                                 _fp, C.c_int, _ip]
        L.refd_find_local_max.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]
        L.refd_load_joints.argtypes = [C.c_int, _dp, C.c_int, C.c_int, _dp, _dp]
        L.refd_pos_message.argtypes = [_fp, _fp, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_int]
        L.refd_root_posterior_pos.argtypes = [_dp, C.c_int, _ip, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp, _fp]
        L.refd_condition.argtypes = [_dp, C.c_int, _ip, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _dp, C.c_float, C.c_int, _fp,
                                     C.c_int]
        _dlib = L
    return _dlib


def _epv(ep):
    return np.array([ep.min_part_rotation, ep.max_part_rotation, ep.num_rotation_steps, ep.min_object_scale,
                     ep.max_object_scale, ep.num_scale_steps, ep.strip_border_detections, ep.roi_save_num_samples], np.float64)


def _quiet(fn):
    """The reference prints progress to stdout from C++; keep test logs readable."""
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        return fn()
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def message(ep, child, off_in, off_out, Cm, rot_mean, rot_sigma, scale, sparse, quiet=True):
    """object_detect::computeRotJointMarginal as the reference compiled it.  quiet=False leaves file descriptor 1
    alone (callers that run several threads redirect it once themselves)."""
    child = np.ascontiguousarray(child, np.float32)
    R, H, W = child.shape
    out = np.empty_like(child)
    e, (_a, pa), (_b, pb), (_c, pc) = _epv(ep), _d(off_in), _d(off_out), _d(Cm)
    call = lambda: dlib().refd_message(e.ctypes.data_as(_dp), _f(child), _f(out), R, H, W, pa, pb, pc, float(rot_mean),
                                       float(rot_sigma), float(scale), int(bool(sparse)))
    _quiet(call) if quiet else call()
    return out


def _joint_rows(joints, one_based=False):
    k = 1 if one_based else 0
    return np.array([[j.type, j.child_idx + k, j.parent_idx + k, j.offset_c[0], j.offset_c[1], j.offset_p[0], j.offset_p[1],
                      j.C[0][0], j.C[0][1], j.C[1][0], j.C[1][1], j.rot_mean, j.rot_sigma] for j in joints], np.float64)


def infer(ep, part_conf, joints, unaries, sparse=True):
    """object_detect::computeRootPosteriorRot + computePartMarginals (+ findLocalMax) as the reference compiled them.
    `unaries` [P][S][R][H][W] is masked in place."""
    assert unaries.dtype == np.float32 and unaries.flags.c_contiguous
    P, S, R, H, W = unaries.shape
    det = np.array([int(bool(v)) for v in part_conf.is_detect], np.int32)
    upr = np.array([int(bool(v)) for v in part_conf.is_upright], np.int32)
    roots = [p for p in range(P) if part_conf.is_detect[p] and part_conf.is_root[p]]
    js = _joint_rows(joints)
    root_post = np.empty((S, H, W), np.float32)
    best = np.empty((P, 7), np.float32)
    marg = np.empty((S, P, R, H, W), np.float32)
    cap = int(ep.roi_save_num_samples) + 1
    hyps = np.empty((P, cap, 7), np.float32)
    nh = np.zeros(P, np.int32)
    e = _epv(ep)
    _quiet(lambda: dlib().refd_infer(e.ctypes.data_as(_dp), P, det.ctypes.data_as(_ip), upr.ctypes.data_as(_ip), roots[0],
                                     js.ctypes.data_as(_dp), len(joints), H, W, _f(unaries), int(bool(sparse)), _f(root_post),
                                     _f(best), _f(marg), _f(hyps), cap, nh.ctypes.data_as(_ip)))
    return {"root_post": root_post, "best_conf": best, "marginals": marg,
            "part_hyps": [hyps[p, :nh[p]].copy() for p in range(P)]}


def root_posterior_pos(ep, part_conf, joints, unaries, sparse=True):
    """mergeRotations + computeRootPosterior (objectdetect_findpos.cpp:118-334) as the reference compiled them.
    unaries [P][S][R][H][W] (log domain).  Returns (merged [P][S][H][W], root posterior [S][H][W])."""
    assert unaries.dtype == np.float32 and unaries.flags.c_contiguous
    P, S, R, H, W = unaries.shape
    det = np.array([int(bool(v)) for v in part_conf.is_detect], np.int32)
    roots = [p for p in range(P) if part_conf.is_detect[p] and part_conf.is_root[p]]
    js = _joint_rows(joints)
    merged = np.zeros((P, S, H, W), np.float32)
    rp = np.zeros((S, H, W), np.float32)
    e = _epv(ep)
    _quiet(lambda: dlib().refd_root_posterior_pos(e.ctypes.data_as(_dp), P, det.ctypes.data_as(_ip), roots[0], js.ctypes.data_as(_dp),
                                                  len(joints), H, W, _f(unaries), int(bool(sparse)), _f(merged), _f(rp)))
    return merged, rp


def find_local_max(grid, max_n):
    """object_detect::findLocalMax (objectdetect_aux.cpp:193-261): rows (dim0, x, y, score)."""
    g = np.ascontiguousarray(grid, np.float32)
    cap = max(g.size, 1)
    out = np.empty((cap, 4), np.float64)
    n = _quiet(lambda: dlib().refd_find_local_max(_f(g), g.shape[0], g.shape[1], g.shape[2], int(max_n),
                                                   out.ctypes.data_as(_dp), cap))
    assert n >= 0
    return out[:n].astype(np.float32)


def load_joints(num_parts, joints, flip):
    """object_detect::loadJoints (objectdetect_aux.cpp:54-141) over joints as load_joint would deliver them (0-based
    here, shifted to the files' 1-based ids on the way in).  Returns (rows of 13 doubles with 0-based ids, detC+invC rows)."""
    rows = _joint_rows(joints, one_based=True)
    out = np.empty_like(rows)
    di = np.empty((len(joints), 5), np.float64)
    _quiet(lambda: dlib().refd_load_joints(int(num_parts), rows.ctypes.data_as(_dp), len(joints), int(bool(flip)),
                                           out.ctypes.data_as(_dp), di.ctypes.data_as(_dp)))
    return out, di


def condition(ep, part_conf, unaries, kind, params=None, weight=1.0, pidx=0, dpm=None):
    """The conditioning adds of objectdetect_icps.cpp as the reference compiled them, on a copy of `unaries`
    [P][S][R][H][W]: kind 0 getRotScoreGrid + addExtraUnary (params [P][2]), 1 getPosScoreGrid + addExtraUnary
    (params [P][4] followed by the detected root position), 2 setTorsoPosPrior (params [4]), 3 addDPMScore (dpm [n][H][W])."""
    u = np.ascontiguousarray(unaries, np.float32).copy()
    P, S, R, H, W = u.shape
    det = np.array([int(bool(v)) for v in part_conf.is_detect], np.int32)
    roots = [p for p in range(P) if part_conf.is_detect[p] and part_conf.is_root[p]]
    prm = np.ascontiguousarray(params if params is not None else [0.0], np.float64).ravel()
    g = np.ascontiguousarray(dpm, np.float32) if dpm is not None else np.zeros((1, H, W), np.float32)
    e = _epv(ep)
    _quiet(lambda: dlib().refd_condition(e.ctypes.data_as(_dp), P, det.ctypes.data_as(_ip), roots[0], H, W, _f(u), int(kind),
                                         prm.ctypes.data_as(_dp), float(weight), int(pidx), _f(g), g.shape[0]))
    return u


def pos_message(child, offset, Cm, scale, sparse):
    """object_detect::computePosJointMarginal as the reference compiled it, on each [H][W] slice of child [D][H][W].
    Returns (log_prob_parent, log_prob_child as the reference leaves it)."""
    ch = np.ascontiguousarray(child, np.float32).copy()
    out = np.empty_like(ch)
    (_o, po), (_c, pc) = _d(offset), _d(Cm)
    for d in range(ch.shape[0]):
        _quiet(lambda: dlib().refd_pos_message(_f(ch[d]), _f(out[d]), ch.shape[1], ch.shape[2], po, pc, float(scale),
                                               int(bool(sparse))))
    return out, ch


# ---- the reference's evaluator (oracle/_ref/libps_ref_eval.so = libPartEval/parteval.cpp + libPartDetect/partdef.cpp) ----
EVAL_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libps_ref_eval.so")
_elib = None


def eval_available():
    return os.path.exists(EVAL_PATH)


def elib():
    global _elib
    if _elib is None:
        _elib = C.CDLL(EVAL_PATH)
        _elib.refe_bbox_endpoints.argtypes = [_dp, _dp]
        _elib.refe_is_gt_match.argtypes = [_dp, _dp, C.c_int, C.c_float]
        _elib.refe_bbox_merge.argtypes = [C.c_int, _dp, C.c_float, C.c_float, C.c_int, _dp]
        _elib.refe_get_part_bbox.argtypes = [_ip, _ip, _ip, C.c_int, _ip, C.c_int, _ip, C.c_int, _ip, C.c_int, _dp, C.c_double, _dp]
        _elib.refe_vis_eval_helper.argtypes = [C.c_char_p, _fp, C.c_int, _ip, _dp, _dp, C.c_int, C.c_float, C.c_float, C.c_int,
                                               _dp, C.c_int]
    return _elib


def _i(a):
    a = np.ascontiguousarray(a, np.int32)
    return a, a.ctypes.data_as(_ip)


def eval_endpoints(bbox10):
    b, bp = _d(bbox10)
    out, op = _d(np.zeros(5))
    _quiet(lambda: elib().refe_bbox_endpoints(bp, op))
    return out


def eval_is_gt_match(gt10, det10, factor=0.5, match_x_axis=False):
    g, gp = _d(gt10)
    d, dp = _d(det10)
    return bool(_quiet(lambda: elib().refe_is_gt_match(gp, dp, int(match_x_axis), float(factor))))


def eval_bbox_merge(kind, boxes, rot_range=(-180.0, 180.0, 48)):
    b, bp = _d(np.concatenate([np.asarray(x, np.float64) for x in boxes]))
    out, op = _d(np.zeros(10))
    _quiet(lambda: elib().refe_bbox_merge(int(kind), bp, float(rot_range[0]), float(rot_range[1]), int(rot_range[2]), op))
    return out


def eval_get_part_bbox(points, part_pos, axis_from, axis_to, offset_and_ext, scale):
    """points: {id: (x, y)}; offset_and_ext = (part_x_axis_offset, ext_x_pos, ext_x_neg, ext_y_pos, ext_y_neg).
    Returns None (axis invalid), False (a point of the part is missing) or the 10 bbox numbers."""
    ids = sorted(points)
    _, idp = _i(ids)
    xs, xp = _i([points[k][0] for k in ids])
    ys, yp = _i([points[k][1] for k in ids])
    pos, pp = _i(part_pos)
    fr, frp = _i(axis_from if len(axis_from) else [0])
    to, top = _i(axis_to if len(axis_to) else [0])
    f5, f5p = _d(offset_and_ext)
    out, op = _d(np.zeros(10))
    rc = _quiet(lambda: elib().refe_get_part_bbox(idp, xp, yp, len(ids), pp, len(part_pos), frp, len(axis_from), top, len(axis_to),
                                                  f5p, float(scale), op))
    return False if rc < 0 else (None if rc == 0 else out)


def eval_vis_eval_helper(part_conf_type, best_conf, window, ext, ext_eval, rot_range=(-180.0, 180.0, 48)):
    """best_conf [P][7]; window [P][4] ints; ext [P][4], ext_eval [Pe][4] = (ext_x_pos, ext_x_neg, ext_y_pos, ext_y_neg)."""
    bc = np.ascontiguousarray(best_conf, np.float32)
    w, wp = _i(window)
    e, ep_ = _d(ext)
    ee, eep = _d(ext_eval)
    out, op = _d(np.zeros(10 * 64))
    n = _quiet(lambda: elib().refe_vis_eval_helper(part_conf_type.encode(), _f(bc), bc.shape[0], wp, ep_, eep, len(np.asarray(ext_eval)),
                                                   float(rot_range[0]), float(rot_range[1]), int(rot_range[2]), op, 64))
    return out[:10 * n].reshape(n, 10)
