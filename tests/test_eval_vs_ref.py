"""partapp_b200/parteval.py against the reference's OWN evaluator code: libPartEval/parteval.cpp and
libPartDetect/partdef.cpp compiled unmodified into oracle/_ref/libps_ref_eval.so (oracle/Makefile target `ref`).

Live comparisons run where that library exists (this container); the same cases are also held as fixtures in
tests/golden/ref_eval.npz (written by tests/golden/make_ref_eval_golden.py) so that they run on the GPU box, where
/root/reference does not exist."""
import os

import numpy as np
import pytest

from oracle import refcore
from partapp_b200 import parteval as pe

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ref_eval.npz")
live = pytest.mark.skipif(not refcore.eval_available(), reason="oracle/_ref/libps_ref_eval.so not built (needs /root/reference)")


def box10(b):
    return np.array([b.part_pos[0], b.part_pos[1], b.part_x_axis[0], b.part_x_axis[1], b.part_y_axis[0], b.part_y_axis[1],
                     b.min_proj_x, b.max_proj_x, b.min_proj_y, b.max_proj_y], np.float64)


def from10(v):
    return pe.PartBBox(np.array(v[0:2], np.float64), np.array(v[2:4], np.float64), np.array(v[4:6], np.float64), float(v[6]),
                       float(v[7]), float(v[8]), float(v[9]))


def random_box(rng):
    th = rng.uniform(-np.pi, np.pi)
    ax = np.array([np.cos(th), np.sin(th)])
    lo_x, lo_y = -rng.uniform(2, 30), -rng.uniform(2, 60)
    return pe.PartBBox(rng.uniform(0, 400, 2), ax, np.array([-ax[1], ax[0]]), lo_x, lo_x + rng.uniform(5, 60), lo_y,
                       lo_y + rng.uniform(5, 120))


# ---- case generators shared by the live tests and the golden writer ---------------------------------------------------
def match_cases(n=200, seed=1):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        gt = random_box(rng)
        det = pe._copy_bbox(gt)
        det.part_pos = gt.part_pos + rng.normal(0, 0.25 * (gt.max_proj_y - gt.min_proj_y) * rng.uniform(0, 1.5), 2)
        th = np.arctan2(gt.part_x_axis[1], gt.part_x_axis[0]) + rng.normal(0, 0.3)
        det.part_x_axis = np.array([np.cos(th), np.sin(th)])
        det.part_y_axis = np.array([-det.part_x_axis[1], det.part_x_axis[0]])
        out.append((box10(gt), box10(det), [0.5, 0.3, 1.0][k % 3]))
    return out


def merge_cases(n=60, seed=2):
    rng = np.random.default_rng(seed)
    return [[box10(random_box(rng)) for _ in range(4)] for _ in range(n)]


def part_bbox_cases(n=80, seed=3):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        pts = {i: (int(rng.integers(0, 500)), int(rng.integers(0, 500))) for i in range(8)}
        if k % 4 == 3:
            pts[1] = pts[0]                              # degenerate axis: from == to
        npos = [1, 2, 3, 4][k % 4]
        pos = [int(v) for v in rng.choice(8, npos, replace=False)]
        if k % 4 == 3:
            fr, to = [0], [1]
        elif k % 5 == 0:
            fr, to = [], []
        elif npos >= 3 and k % 2 == 0:
            fr, to = [int(v) for v in rng.choice(8, 2, replace=False)], [int(v) for v in rng.choice(8, 2, replace=False)]
        else:
            a, b = rng.choice(8, 2, replace=False)
            fr, to = [int(a)], [int(b)]
        f5 = [float(rng.choice([0, 90, -90, 180])), rng.uniform(0, 30), rng.uniform(0, 30), rng.uniform(0, 30), rng.uniform(0, 30)]
        f5 = [float(np.float32(v)) for v in f5]          # proto floats
        out.append((pts, pos, fr, to, f5, float(rng.uniform(0.5, 1.5))))
    return out


PART_COUNTS = {"human_full": 10, "human_full_joints": 18, "human_full_torso4": 22, "human_full_14_parts": 14,
               "human_full_22_parts": 22, "human_full_12_parts": 12, "other": 6}


def helper_cases(seed=4):
    rng = np.random.default_rng(seed)
    out = []
    for t, P in PART_COUNTS.items():
        for rep in range(3):
            best = np.zeros((P, 7), np.float32)
            best[:, 1] = np.float32(rng.choice([0.8, 1.0, 1.25]))
            best[:, 2] = rng.integers(0, 48, P)
            best[:, 3] = -180 + 360.0 / 48 * (0.5 + best[:, 2])
            best[:, 4] = rng.integers(20, 380, P)
            best[:, 5] = rng.integers(20, 580, P)
            best[:, 6] = rng.uniform(-20, 0, P)
            window = np.column_stack([rng.integers(20, 80, P), rng.integers(40, 140, P), rng.integers(5, 40, P),
                                      rng.integers(5, 70, P)]).astype(np.int32)
            ext = np.float32(rng.uniform(5, 40, (P, 4))).astype(np.float64)
            Pe = P if t in ("human_full", "other") else (6 if t == "human_full_12_parts" else 10)
            ext_eval = np.float32(rng.uniform(2, 30, (Pe, 4))).astype(np.float64)
            out.append((t, best, window, ext, ext_eval))
    return out


def py_endpoints(b10):
    t, m, ln = pe.get_bbox_endpoints(from10(b10))
    return np.array([t[0], t[1], m[0], m[1], ln])


def py_merge(kind, boxes, rot_range=(-180.0, 180.0, 48)):
    b = [from10(x) for x in boxes]
    r = pe.bbox_merge4(*b[:4]) if kind == 4 else (pe.bbox_merge2(b[0], b[1]) if kind == 2 else pe.bbox_merge_rot(b[0], b[1], rot_range))
    return box10(r)


def py_part_bbox(case):
    pts, pos, fr, to, f5, scale = case
    rect = pe.AnnoRect(points=dict(pts))
    pd = pe.PartDef(0, pos, fr, to, f5[0], f5[1], f5[2], f5[3], f5[4])
    if not pe.annorect_has_part(rect, pd):
        return False
    b = pe.get_part_bbox(rect, pd, scale)
    return None if b is None else box10(b)


def py_helper(case):
    t, best, window, ext, ext_eval = case
    pps = [pe.PartParam(*[int(v) for v in w]) for w in window]
    boxes = [pe.bbox_from_hyp(best[p], pps[p]) for p in range(len(best))]
    mk = lambda e: [pe.PartDef(0, [], [], [], 0.0, float(r[0]), float(r[1]), float(r[2]), float(r[3])) for r in e]
    out = pe.convert_eval_bboxes(t, boxes, [float(v) for v in best[:, 1]], mk(ext_eval), mk(ext))
    return np.stack([box10(b) for b in out])


def ref_helper(case):
    t, best, window, ext, ext_eval = case
    return refcore.eval_vis_eval_helper(t, best, window, ext, ext_eval)


def _close(a, b, what):
    # same formulas in double: differences only from numpy's vs uBLAS's summation inside dot products (1 ulp scale)
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-9, err_msg=what)


# ---- live: the reference's compiled code ------------------------------------------------------------------------------
@live
def test_endpoints_and_matching_rule_equal_reference_code():
    n_match = 0
    for gt, det, f in match_cases():
        _close(py_endpoints(gt), refcore.eval_endpoints(gt), "get_bbox_endpoints")
        want = refcore.eval_is_gt_match(gt, det, f)
        assert pe.is_gt_match_bbox(from10(gt), from10(det), f) == want
        n_match += want
    assert 20 < n_match < 180          # both outcomes occur


@live
def test_bbox_merge_variants_equal_reference_code():
    for boxes in merge_cases():
        _close(py_merge(2, boxes), refcore.eval_bbox_merge(2, boxes[:2]), "bbox_merge(2)")
        _close(py_merge(4, boxes), refcore.eval_bbox_merge(4, boxes), "bbox_merge(4)")
        for rr in ((-180.0, 180.0, 48), (-180.0, 180.0, 24), (-90.0, 90.0, 12)):
            _close(py_merge(3, boxes, rr), refcore.eval_bbox_merge(3, boxes[:2], rr), "bbox_merge(exp_param)")


@live
def test_get_part_bbox_equals_reference_code():
    kinds = set()
    for case in part_bbox_cases():
        pts, pos, fr, to, f5, scale = case
        got = py_part_bbox(case)
        want = refcore.eval_get_part_bbox(pts, pos, fr, to, f5, scale)
        if want is None or want is False:
            assert got is want
            kinds.add(str(want))
        else:
            _close(got, want, "get_part_bbox")
            kinds.add("box%d" % min(len(pos), 3))
    assert {"None", "box1", "box2", "box3"} <= kinds


@live
@pytest.mark.parametrize("case", helper_cases(), ids=lambda c: c[0])
def test_model_part_to_evaluation_part_conversion_equals_reference_code(case):
    got, want = py_helper(case), ref_helper(case)
    assert got.shape == want.shape, case[0]
    _close(got, want, "vis_eval_helper (%s)" % case[0])


# ---- golden: the same cases, recorded outputs of the reference's code -------------------------------------------------
def test_evaluator_matches_recorded_reference_outputs():
    g = np.load(GOLDEN, allow_pickle=False)
    for k, (gt, det, f) in enumerate(match_cases()):
        _close(py_endpoints(gt), g["endpoints"][k], "endpoints %d" % k)
        assert pe.is_gt_match_bbox(from10(gt), from10(det), f) == bool(g["match"][k])
    for k, boxes in enumerate(merge_cases()):
        _close(py_merge(2, boxes), g["merge2"][k], "merge2")
        _close(py_merge(4, boxes), g["merge4"][k], "merge4")
        _close(py_merge(3, boxes), g["merge3"][k], "merge3")
    for k, case in enumerate(part_bbox_cases()):
        got = py_part_bbox(case)
        code = int(g["part_bbox_code"][k])
        if code == 1:
            _close(got, g["part_bbox"][k], "get_part_bbox %d" % k)
        else:
            assert got is (None if code == 0 else False)
    for k, case in enumerate(helper_cases()):
        want = g["helper_%d" % k]
        _close(py_helper(case), want, "vis_eval_helper %s" % case[0])
