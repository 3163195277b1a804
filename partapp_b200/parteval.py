"""PCP matching rule of the reference's evaluator, for the "PCP-identical part estimates" check of BASELINE.json.

Restates (host side, numpy -- this is evaluation, not the hot path):
  bbox_from_pos       reference src/libs/libPartApp/partapp.cpp:59-81
  get_bbox_endpoints  src/libs/libPartEval/parteval.cpp:45-68 (the non-endpoint branch)
  is_gt_match         src/libs/libPartEval/parteval.cpp:79-110 (match_x_axis == false): both endpoint distances
                      strictly below factor * ground-truth segment length, factor 0.5 for PCP.
A part estimate is a `best_conf` row (PartHyp::toVect, objectdetect.h:139-160):
[scaleidx, scale, rotidx, rot_deg, x, y, score].
"""
import math
import os
import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class PartParam:
    """PartWindowParam.PartParam (libPartDetect/PartWindowParam.proto:3-19)."""
    window_size_x: int = 0
    window_size_y: int = 0
    pos_offset_x: int = 0
    pos_offset_y: int = 0


def bbox_endpoints(row, pp: PartParam):
    # PartHyp::getPartBBox, objectdetect.h:107-113: m_rot / 180 is a float division, the product with M_PI is double
    scale, rot = float(row[1]), float(np.float32(row[3]) / np.float32(180)) * np.pi
    pos = np.array([float(int(row[4])), float(int(row[5]))])
    x_axis = np.array([np.cos(rot), np.sin(rot)])
    y_axis = np.array([-x_axis[1], x_axis[0]])
    min_proj_y = -scale * pp.pos_offset_y
    max_proj_y = min_proj_y + scale * pp.window_size_y
    return pos + min_proj_y * y_axis, pos + max_proj_y * y_axis, max_proj_y - min_proj_y


def is_gt_match(gt_row, det_row, pp: PartParam, factor=0.5):
    gt_top, gt_bot, gt_len = bbox_endpoints(gt_row, pp)
    d_top, d_bot, _ = bbox_endpoints(det_row, pp)
    return bool(np.linalg.norm(gt_top - d_top) < factor * gt_len and np.linalg.norm(gt_bot - d_bot) < factor * gt_len)


def pcp_identical(best_conf_ref, best_conf_test, part_params: Sequence[PartParam], factor=0.5):
    """Fraction of parts whose estimate in `best_conf_test` PCP-matches the estimate in `best_conf_ref`
    (the reference's estimate plays the role of the ground truth)."""
    ref = np.asarray(best_conf_ref).reshape(-1, 7)
    tst = np.asarray(best_conf_test).reshape(-1, 7)
    assert ref.shape == tst.shape and len(part_params) >= 1
    n = len(ref)
    ok = sum(is_gt_match(ref[i], tst[i], part_params[i % len(part_params)], factor) for i in range(n))
    return ok / float(n)


# ---------------------------------------------------------------------------------------------------------------------
# eval_segments: the reference's PCP evaluation of a --find_obj run.  Restates, for EVAL_TYPE_PS,
#   eval_segments        libPartEval/parteval.cpp:1162-1345
#   vis_eval_helper      parteval.cpp:312-383,520-524 (loadPartHyp :186-201 -> PartHyp::getPartBBox)
#   get_part_bbox & co.  libPartDetect/partdef.cpp:91-340, annorect_has_part :468-482
#   AnnoRect parsing     libAnnotation/annorect.cpp:30-130 (annopoint coordinates are read as ints)
#   model -> evaluation parts   parteval.cpp:520-857 (the tail of vis_eval_helper, per ExpParam.part_conf_type),
#                               bbox_merge :208-285, get_shrink_factor_x :287-297


def parse_prototext(text: str) -> dict:
    """Protobuf text format -> {field: [values...]}; nested messages become dicts.  Enough for ExpParam,
    PartConfig and PartWindowParam files (libProtoBuf/protobuf_aux.hpp:33-60 parses them with TextFormat)."""
    tok = re.findall(r'"(?:[^"\\]|\\.)*"|\'(?:[^\'\\]|\\.)*\'|[{}<>:]|[^\s{}<>:#"\']+|#[^\n]*', text)
    tok = [t for t in tok if not t.startswith('#')]
    pos = 0

    def scalar(t):
        if t[0] in '"\'':
            return bytes(t[1:-1], 'utf-8').decode('unicode_escape')
        if t in ('true', 'True'):
            return True
        if t in ('false', 'False'):
            return False
        try:
            return int(t, 0)
        except ValueError:
            pass
        try:
            return float(t.rstrip('fF'))
        except ValueError:
            return t  # enum identifier

    def message(closer):
        nonlocal pos
        out: Dict[str, list] = {}
        while pos < len(tok) and tok[pos] != closer:
            name = tok[pos]
            pos += 1
            if tok[pos] == ':':
                pos += 1
            if tok[pos] in '{<':
                close = '}' if tok[pos] == '{' else '>'
                pos += 1
                val = message(close)
                pos += 1
            else:
                val = scalar(tok[pos])
                pos += 1
                while pos < len(tok) and tok[pos][0] in '"\'' and isinstance(val, str):  # adjacent string literals concatenate
                    val += scalar(tok[pos])
                    pos += 1
            out.setdefault(name, []).append(val)
        return out

    return message(None)


@dataclass
class PartDef:
    """libPartDetect/PartConfig.proto:1-37, the fields the evaluation reads."""
    part_id: int = 0
    part_pos: List[int] = field(default_factory=list)
    part_x_axis_from: List[int] = field(default_factory=list)
    part_x_axis_to: List[int] = field(default_factory=list)
    part_x_axis_offset: float = 0.0
    ext_x_pos: float = 0.0
    ext_x_neg: float = 0.0
    ext_y_pos: float = 0.0
    ext_y_neg: float = 0.0

    @staticmethod
    def from_message(m: dict) -> "PartDef":
        one = lambda k, d: m[k][-1] if k in m else d
        return PartDef(int(one("part_id", 0)), [int(v) for v in m.get("part_pos", [])],
                       [int(v) for v in m.get("part_x_axis_from", [])], [int(v) for v in m.get("part_x_axis_to", [])],
                       float(one("part_x_axis_offset", 0)), float(one("ext_x_pos", 0)), float(one("ext_x_neg", 0)),
                       float(one("ext_y_pos", 0)), float(one("ext_y_neg", 0)))


def load_part_conf(path: str) -> List[PartDef]:
    with open(path) as f:
        return [PartDef.from_message(m) for m in parse_prototext(f.read()).get("part", [])]


def load_window_param(path: str) -> List[PartParam]:
    with open(path) as f:
        msg = parse_prototext(f.read())
    one = lambda m, k: int(m[k][-1]) if k in m else 0
    return [PartParam(one(m, "window_size_x"), one(m, "window_size_y"), one(m, "pos_offset_x"), one(m, "pos_offset_y"))
            for m in msg.get("part", [])]


@dataclass
class AnnoRect:
    x1: float = 0.0
    y1: float = 0.0
    x2: float = 0.0
    y2: float = 0.0
    points: Dict[int, Tuple[int, int]] = field(default_factory=dict)  # id -> (x, y); first occurrence wins


@dataclass
class Annotation:
    image: str = ""
    rects: List[AnnoRect] = field(default_factory=list)


def load_annolist(path: str) -> List[Annotation]:
    """XML annotation list (.al).  Coordinates of annopoints are truncated to int like getElementDataInt (atoi)."""
    def to_int(s):
        m = re.match(r'\s*[-+]?\d+', s or "")
        return int(m.group(0)) if m else 0

    def to_float(s):
        try:
            return float(s)
        except (TypeError, ValueError):
            return 0.0

    out = []
    for a in ET.parse(path).getroot().iter("annotation"):
        ann = Annotation(image=(a.findtext("image/name") or "").strip())
        for r in a.findall("annorect"):
            rect = AnnoRect(to_float(r.findtext("x1")), to_float(r.findtext("y1")), to_float(r.findtext("x2")),
                            to_float(r.findtext("y2")))
            for p in r.findall("annopoints/point"):
                pid = to_int(p.findtext("id"))
                rect.points.setdefault(pid, (to_int(p.findtext("x")), to_int(p.findtext("y"))))
            ann.rects.append(rect)
        out.append(ann)
    return out


@dataclass
class PartBBox:
    """libPartDetect/partdef.h PartBBox (the fields the matching rule reads)."""
    part_pos: np.ndarray
    part_x_axis: np.ndarray
    part_y_axis: np.ndarray
    min_proj_x: float = 0.0
    max_proj_x: float = 0.0
    min_proj_y: float = 0.0
    max_proj_y: float = 0.0


def _rotation(rad):  # boost_math.hpp:39-50
    c, s = math.cos(rad), math.sin(rad)
    return np.array([[c, -s], [s, c]])


def annorect_has_part(rect: AnnoRect, pd: PartDef) -> bool:  # partdef.cpp:468-482
    return all(i in rect.points for i in pd.part_pos)


def _axis(rect: AnnoRect, pd: PartDef) -> Optional[np.ndarray]:
    """get_part_x_axis / get_part_x_axis_complex (partdef.cpp:155-243); None = invalid axis."""
    nf, nt = len(pd.part_x_axis_from), len(pd.part_x_axis_to)
    if nf <= 1 and nt <= 1:
        if nf == 0 and nt == 0:
            return np.array([1.0, 0.0])
        f, t = rect.points.get(pd.part_x_axis_from[0]), rect.points.get(pd.part_x_axis_to[0])
        if f is None or t is None:
            return None
        v = np.array([float(t[0] - f[0]), float(t[1] - f[1])])
    else:
        # the complex variant accumulates in float32 (partdef.cpp:162-190)
        fr = np.zeros(2, np.float32)
        to = np.zeros(2, np.float32)
        for i in pd.part_x_axis_from:
            fr += np.array(rect.points[i], np.float32)
        fr /= np.float32(nf)
        for i in pd.part_x_axis_to:
            to += np.array(rect.points[i], np.float32)
        to /= np.float32(nt)
        v = (to - fr).astype(np.float64)
    if not (abs(v[0]) > 1e-6 or abs(v[1]) > 1e-6):
        return None
    return _rotation(pd.part_x_axis_offset * math.pi / 180.0) @ v


def get_part_bbox(rect: AnnoRect, pd: PartDef, scale: float) -> Optional[PartBBox]:
    """get_part_bbox (partdef.cpp:352-360): atomic for < 3 position points, else the corner-based variant."""
    assert annorect_has_part(rect, pd)
    pts = [np.array(rect.points[i], np.float64) for i in pd.part_pos]
    if len(pd.part_pos) < 3:
        pos = np.zeros(2)
        for p in pts:
            pos = pos + p
        pos = pos * (1.0 / len(pts))                                  # get_part_position, partdef.cpp:140-158
        ref_pts = pts
        if len(pd.part_x_axis_from) > 1 or len(pd.part_x_axis_to) > 1:
            raise AssertionError("get_part_x_axis: more than one axis point on an atomic part (partdef.cpp:200)")
    else:
        xs, ys = [p[0] for p in pts], [p[1] for p in pts]
        x0, x1, y0, y1 = min(xs), max(xs), min(ys), max(ys)
        pos = np.array([0.5 * (x0 + x1), 0.5 * (y0 + y1)])            # get_part_position_complex, :128-138
        ref_pts = [np.array(c) for c in ((x0, y0), (x1, y1), (x0, y1), (x1, y0))]  # get_bbox_corners, :91-126
    ax = _axis(rect, pd)
    if ax is None:
        return None
    ax = ax / math.sqrt(ax[0] * ax[0] + ax[1] * ax[1])
    ay = _rotation(math.pi / 2) @ ax
    px = [float(ax @ (p - pos)) for p in ref_pts]
    py = [float(ay @ (p - pos)) for p in ref_pts]
    return PartBBox(pos, ax, ay, min(px) - scale * pd.ext_x_neg, max(px) + scale * pd.ext_x_pos,
                    min(py) - scale * pd.ext_y_neg, max(py) + scale * pd.ext_y_pos)


def bbox_from_hyp(row, pp: PartParam) -> PartBBox:
    """PartHyp::getPartBBox -> bbox_from_pos (objectdetect.h:107-113, partapp.cpp:59-81) for a best_conf row."""
    # m_rot is a float and `m_rot / 180` a FLOAT division (objectdetect.h:111); only the product with M_PI is double.
    # Pinned against the reference's compiled code by tests/test_eval_vs_ref.py.
    scale, rot = float(row[1]), float(np.float32(row[3]) / np.float32(180)) * math.pi
    ax = np.array([math.cos(rot), math.sin(rot)])
    ay = np.array([-ax[1], ax[0]])
    mx, my = -scale * pp.pos_offset_x, -scale * pp.pos_offset_y
    return PartBBox(np.array([float(int(row[4])), float(int(row[5]))]), ax, ay, mx, mx + scale * pp.window_size_x, my,
                    my + scale * pp.window_size_y)


def get_bbox_endpoints(b: PartBBox):  # parteval.cpp:45-68, use_endpoints == false
    return b.part_pos + b.min_proj_y * b.part_y_axis, b.part_pos + b.max_proj_y * b.part_y_axis, b.max_proj_y - b.min_proj_y


def is_gt_match_bbox(gt: PartBBox, det: PartBBox, factor: float = 0.5) -> bool:  # parteval.cpp:79-110
    assert 0 <= factor <= 1.0
    gt_top, gt_bot, gt_len = get_bbox_endpoints(gt)
    d_top, d_bot, _ = get_bbox_endpoints(det)
    return bool(np.linalg.norm(gt_top - d_top) < factor * gt_len and np.linalg.norm(gt_bot - d_bot) < factor * gt_len)


def _copy_bbox(b: PartBBox) -> PartBBox:
    return PartBBox(b.part_pos.copy(), b.part_x_axis.copy(), b.part_y_axis.copy(), b.min_proj_x, b.max_proj_x,
                    b.min_proj_y, b.max_proj_y)


def bbox_merge2(b1: PartBBox, b2: PartBBox) -> PartBBox:
    """bbox_merge(bbox1, bbox2) (parteval.cpp:225-235): two joint parts of one limb; axes and x extent of the first."""
    r = _copy_bbox(b1)
    r.part_pos = 0.5 * (b1.part_pos + b2.part_pos)
    r.min_proj_y = float(r.part_y_axis @ (b2.part_pos - r.part_pos))
    r.max_proj_y = float(r.part_y_axis @ (b1.part_pos - r.part_pos))
    return r


def bbox_merge4(b1: PartBBox, b2: PartBBox, b3: PartBBox, b4: PartBBox) -> PartBBox:
    """bbox_merge of four corner parts (parteval.cpp:208-223): the torso of "human_full_torso4"."""
    r = _copy_bbox(b1)
    r.part_pos = 0.25 * (b1.part_pos + b2.part_pos + b3.part_pos + b4.part_pos)
    py = lambda b: float(r.part_y_axis @ (b.part_pos - r.part_pos))
    px = lambda b: float(r.part_x_axis @ (b.part_pos - r.part_pos))
    r.min_proj_y, r.max_proj_y = min(py(b3), py(b4)), max(py(b1), py(b2))
    r.min_proj_x, r.max_proj_x = min(px(b1), px(b4)), max(px(b2), px(b3))
    return r


def bbox_merge_rot(b1: PartBBox, b2: PartBBox, rot_range: Tuple[float, float, int]) -> PartBBox:
    """bbox_merge(bbox1, bbox2, exp_param) (parteval.cpp:237-285): the stick between two joint positions, its axis
    snapped to the nearest rotation bin centre (first minimum of |atan2 - bin|, bins not wrapped)."""
    pos = 0.5 * (b1.part_pos + b2.part_pos)
    min_x = min(b1.part_pos[0], b2.part_pos[0]) - pos[0]
    max_x = max(b1.part_pos[0], b2.part_pos[0]) - pos[0]
    min_y = min(b1.part_pos[1], b2.part_pos[1]) - pos[1]
    max_y = max(b1.part_pos[1], b2.part_pos[1]) - pos[1]
    lo, hi = (min_x, max_x) if max_x - min_x > max_y - min_y else (min_y, max_y)
    d = b1.part_pos - b2.part_pos
    part_rot = math.atan2(d[1], d[0])
    rmin, rmax, n = rot_range
    best, best_diff = None, float("inf")
    for ridx in range(n):
        deg = rmin if rmin == rmax else rmin + (rmax - rmin) / n * (0.5 + ridx)      # rot_from_index, partapp_aux.hpp:45-58
        disc = deg / 180.0 * math.pi
        if best_diff > abs(part_rot - disc):
            best_diff, best = abs(part_rot - disc), disc
    ay = np.array([math.cos(best), math.sin(best)])
    return PartBBox(pos, np.array([-ay[1], ay[0]]), ay, b1.min_proj_x, b1.max_proj_x, lo, hi)


def get_shrink_factor_x(ext_x_pos: float, pidx: int, rootidx: int = 4) -> float:
    """parteval.cpp:287-297; float arithmetic like the reference (the result only scales the x extent)."""
    off = 30.0 if pidx == rootidx else (20.0 if pidx == rootidx + 1 else 15.0)
    return float(np.float32(0.7 * off / ext_x_pos)) if ext_x_pos != 0 else float("inf")


def convert_eval_bboxes(part_conf_type: str, boxes: Sequence[PartBBox], scales: Sequence[float],
                        part_conf_eval: Sequence[PartDef], part_conf: Sequence[PartDef],
                        rot_range: Tuple[float, float, int] = (-180.0, 180.0, 48), merge: bool = True) -> List[PartBBox]:
    """Tail of vis_eval_helper (parteval.cpp:520-857): from the model's part boxes (PartHyp::getPartBBox of every
    best_conf row; `scales[i]` = that hypothesis's m_scale) to the boxes of the evaluation parts -- joint parts merged
    into limbs where the model has them, the y extension of the detection window removed (PCP compares stick ends),
    the x extent shrunk for display.  `part_conf` is the model's PartConfig (ext_x_pos of the shrink)."""
    b = [_copy_bbox(x) for x in boxes]
    ev = part_conf_eval

    def strip_y(box, i, scale=1.0):
        box.min_proj_y += scale * ev[i].ext_y_neg
        box.max_proj_y -= scale * ev[i].ext_y_pos

    def shrink(out, ext, rootidx=4, skip=()):
        for i, box in enumerate(out):
            if i in skip:
                continue
            f = get_shrink_factor_x(ext[i], i, rootidx)
            box.max_proj_x *= f
            box.min_proj_x *= f
        return out

    t = part_conf_type
    if t == "human_full_joints":                                                      # :527-578
        assert len(b) == 18
        out = [bbox_merge2(b[0], b[1]), bbox_merge2(b[2], b[3]), bbox_merge2(b[6], b[7]), bbox_merge2(b[4], b[5]),
               _copy_bbox(b[8]), _copy_bbox(b[17]),
               bbox_merge2(b[9], b[10]), bbox_merge2(b[11], b[12]), bbox_merge2(b[15], b[16]), bbox_merge2(b[13], b[14])]
        strip_y(out[4], 4, scales[8])
        strip_y(out[5], 5, scales[17])
        return shrink(out, [part_conf[i].ext_x_pos for i in range(len(out))])
    if t == "human_full_torso4":                                                      # :579-630
        assert len(b) == 22
        out = [bbox_merge2(b[0], b[1]), bbox_merge2(b[2], b[3]), bbox_merge2(b[6], b[7]), bbox_merge2(b[4], b[5]),
               bbox_merge4(b[16], b[17], b[18], b[19]), _copy_bbox(b[20]),
               bbox_merge2(b[8], b[9]), bbox_merge2(b[10], b[11]), bbox_merge2(b[14], b[15]), bbox_merge2(b[12], b[13])]
        strip_y(out[5], 5, scales[20])
        return shrink(out, [part_conf[i].ext_x_pos for i in range(len(out))], skip=(4,))
    if t == "human_full_14_parts" and merge:                                          # :631-689
        assert len(b) == 14
        m = lambda i, j: bbox_merge_rot(b[i], b[j], rot_range)
        out = [m(0, 1), m(1, 2), m(4, 3), m(5, 4), _copy_bbox(b[6]), _copy_bbox(b[7]), m(8, 9), m(9, 10), m(12, 11),
               m(13, 12)]
        strip_y(out[4], 4, scales[6])
        strip_y(out[5], 5, scales[7])
        ext = [part_conf[i].ext_x_pos for i in (0, 1, 2, 3, 6, 7, 9, 10, 11, 12)]
        return shrink(out, ext)
    if t == "human_full_22_parts" and merge:                                          # :690-769
        assert len(b) == 22
        out = [_copy_bbox(b[i]) for i in (1, 3, 6, 8, 10, 11, 13, 15, 18, 20)]
        for i in (0, 1, 2, 3, 6, 7, 8, 9):
            strip_y(out[i], i)
        strip_y(out[4], 4, scales[10])
        strip_y(out[5], 5, scales[11])
        ext = [part_conf[i].ext_x_pos for i in (1, 3, 6, 8, 10, 11, 12, 14, 16, 18)]
        return shrink(out, ext)
    if t == "human_full_12_parts" and merge:                                          # :770-824
        assert len(b) == 12
        out = [_copy_bbox(b[i]) for i in (0, 1, 3, 5, 8, 10)]
        strip_y(out[0], 0, scales[0])
        strip_y(out[1], 1, scales[1])
        for i in (2, 3, 4, 5):
            strip_y(out[i], i)
        ext = [part_conf[i].ext_x_pos for i in (0, 1, 3, 5, 8, 10)]
        return shrink(out, ext, rootidx=0)
    # every other type (:825-849): one evaluation part per model part
    for i, box in enumerate(b):
        assert scales[i] > 0
        strip_y(box, i, scales[i])
        f = get_shrink_factor_x(part_conf[i].ext_x_pos, i) if t == "human_full" else float(np.float32(0.7))
        box.max_proj_x *= f
        box.min_proj_x *= f
    return b


@dataclass
class SegmentEval:
    ratio: float
    seg_correct: int
    seg_total: int
    per_part_correct: List[int]
    per_part_total: List[int]
    endpoints: Dict[int, np.ndarray]  # imgidx -> [P + 1][6] matrix the reference saves as seg_endpoints/endpoints_%04d.mat


def eval_segments(annolist: Sequence[Annotation], part_conf_eval: Sequence[PartDef], window_param: Sequence[PartParam],
                  load_best_conf, firstidx: int, lastidx: int, scale: float = 1.0, eval_didx: int = -1,
                  save_dir: Optional[str] = None, part_conf: Optional[Sequence[PartDef]] = None,
                  part_conf_type: str = "human_full", rot_range: Tuple[float, float, int] = (-180.0, 180.0, 48)) -> SegmentEval:
    """EVAL_TYPE_PS branch of eval_segments (parteval.cpp:1162-1345).  `load_best_conf(imgidx)` returns the [P][7]
    `best_conf` of pose_est_imgidx%04d.mat; `scale` is scale_from_index(exp_param, 0) (:1207-1208); `part_conf` is the
    model's part configuration (default: the evaluation one), `part_conf_type` ExpParam.part_conf_type and
    `rot_range` (min_part_rotation, max_part_rotation, num_rotation_steps) -- what vis_eval_helper needs to turn model
    parts into evaluation parts.  With `save_dir`, the per-image endpoint matrices are written like the reference's
    seg_endpoints directory."""
    P = len(part_conf_eval)
    model_conf = part_conf_eval if part_conf is None else part_conf
    correct, total = [0] * P, [0] * P
    seg_correct = seg_total = 0
    endpoints = {}
    for imgidx in range(firstidx, lastidx + 1):
        best_conf = np.asarray(load_best_conf(imgidx), np.float32).reshape(-1, 7)
        assert best_conf.shape[0] == len(model_conf), "nParts == best_conf.shape()[0] (parteval.cpp:194)"
        dets = convert_eval_bboxes(part_conf_type, [bbox_from_hyp(best_conf[i], window_param[i]) for i in range(len(best_conf))],
                                   [float(r[1]) for r in best_conf], part_conf_eval, model_conf, rot_range)
        assert len(dets) == P, "eval_bbox.size() == part_conf_eval.part_size() (parteval.cpp:1250)"
        ep = np.zeros((P + 1, 6))
        n_ok = n_seg = 0
        rects = annolist[imgidx].rects
        assert len(rects) > 0
        rect = rects[eval_didx if eval_didx >= 0 else 0]
        for pidx in range(P):
            det = dets[pidx]
            top, bot, _ = get_bbox_endpoints(det)
            ep[pidx, :4] = [bot[0], bot[1], top[0], top[1]]
            if not rect.points or not annorect_has_part(rect, part_conf_eval[pidx]):
                continue
            gt = get_part_bbox(rect, part_conf_eval[pidx], scale)
            if gt is None:
                # the reference ignores get_part_bbox's return value and would match against an unset box
                raise ValueError("image %d part %d: annotated axis points coincide" % (imgidx, pidx))
            gt.min_proj_y += scale * part_conf_eval[pidx].ext_y_neg   # undo the extension: PCP compares stick ends
            gt.max_proj_y -= scale * part_conf_eval[pidx].ext_y_pos
            match = is_gt_match_bbox(gt, det)
            seg_correct += int(match)
            correct[pidx] += int(match)
            ep[pidx, 4] = 1.0 if match else 0.0
            seg_total += 1
            n_ok += int(match)
            n_seg += 1
            total[pidx] += 1
            ep[pidx, 5] = 1.0
        ep[P, 0] = (1.0 * n_ok / n_seg) if n_seg else float("nan")
        endpoints[imgidx] = ep
        if save_dir is not None:
            import scipy.io
            os.makedirs(save_dir, exist_ok=True)
            scipy.io.savemat(os.path.join(save_dir, "endpoints_%04d.mat" % imgidx), {"endpoints": ep})
    ratio = seg_correct / float(seg_total) if seg_total else 0.0
    return SegmentEval(ratio, seg_correct, seg_total, correct, total, endpoints)


def _image_size(path: str) -> Tuple[int, int]:
    """(width, height) from a PNG / JPEG / PNM header (the reference loads the image for it, findrot.cpp:752-760)."""
    import struct
    with open(path, "rb") as f:
        h = f.read(32)
        if h[:4] == b"\x89PNG":
            return struct.unpack(">II", h[16:24])
        if h[:2] == b"\xff\xd8":
            f.seek(2)
            while True:
                m = f.read(4)
                if len(m) < 4 or m[0] != 0xFF:
                    break
                ln = (m[2] << 8) | m[3]
                if 0xC0 <= m[1] <= 0xCF and m[1] not in (0xC4, 0xC8, 0xCC):
                    s5 = f.read(5)
                    return ((s5[3] << 8) | s5[4], (s5[1] << 8) | s5[2])
                f.seek(ln - 2, 1)
        if h[:1] == b"P" and h[1:2] in (b"2", b"3", b"5", b"6"):
            f.seek(2)
            w, hh = f.read(64).split()[:2]
            return int(w), int(hh)
    raise ValueError("cannot read the size of image " + path)


# ---- the other hypothesis sources of vis_eval_helper (parteval.cpp:326-383, :495-514) --------------------------------
def load_score_grid_direct(cells: np.ndarray, Tig: np.ndarray, height: int, width: int) -> np.ndarray:
    """PartApp::loadScoreGrid with TM_DIRECT (partapp.cpp:830-903, multi_array_transform.hpp:167-192) in numpy: every
    non-zero grid cell (x1 outer, y1 inner: later writers win) lands on round(Tig * (x1, y1)); unevaluated cells stay 0."""
    R, gh, gw = cells.shape
    out = np.zeros((R, height, width), np.float32)
    y1, x1 = np.meshgrid(np.arange(gh), np.arange(gw), indexing="ij")
    order = np.lexsort((y1.ravel(), x1.ravel()))               # scatter order: x1 outer, y1 inner
    xs, ys = x1.ravel()[order].astype(np.float64), y1.ravel()[order].astype(np.float64)
    for r in range(R):
        T = np.asarray(Tig[r], np.float64)
        x3 = (T[0, 0] * xs + T[0, 1] * ys) + T[0, 2]
        y3 = (T[1, 0] * xs + T[1, 1] * ys) + T[1, 2]
        ix, iy = np.floor(x3 + 0.5).astype(np.int64), np.floor(y3 + 0.5).astype(np.int64)
        v = cells[r].ravel()[order]
        ok = (v != 0) & (ix >= 0) & (ix < width) & (iy >= 0) & (iy < height)
        out[r, iy[ok], ix[ok]] = v[ok]                          # duplicates: numpy keeps the last one = the last writer
    return out


def unary_best_hyp(scores: np.ndarray, scaleidx: int, scale: float, rot_range: Tuple[float, float, int]) -> np.ndarray:
    """EVAL_TYPE_UNARIES (parteval.cpp:355-379): the first strict maximum of one part's score grids in the reference's
    scan order -- rotation, then x, then y -- as a best_conf row."""
    R, H, W = scores.shape
    flat = np.transpose(scores, (0, 2, 1)).reshape(-1)          # [r][x][y]
    k = int(np.argmax(flat))                                     # first occurrence of the maximum = strict '>' scan
    r, rem = divmod(k, W * H)
    x, y = divmod(rem, H)
    rmin, rmax, n = rot_range
    rot = rmin if rmin == rmax else rmin + (rmax - rmin) / n * (0.5 + r)
    return np.array([scaleidx, np.float32(scale), r, np.float32(rot), x, y, flat[k]], np.float32)


def disc_ps_best_hyp(vect_scale_idx, vect_rot_idx, vect_iy, vect_ix, posterior, scale_range: Tuple[float, float, int],
                     rot_range: Tuple[float, float, int], vect_didx=None, didx: int = -1) -> np.ndarray:
    """EVAL_TYPE_DISC_PS (parteval.cpp:384-493): a part's estimate is the sample with the first strictly largest
    posterior among the samples libDiscPS drew for it (`samples_pidx<p>.mat`: vect_scale_idx / vect_rot_idx / vect_iy /
    vect_ix [/ vect_didx], `samples_imgidx%04d_post.mat`: samples_post_part<p>); with `didx` >= 0 only the samples of that
    subject compete (:441-482).  Returns the PartHyp of the winner as a best_conf row (objectdetect.h:91-97, :139-160)."""
    post = np.asarray(posterior, np.float64).reshape(-1)
    cols = [np.asarray(v).reshape(-1) for v in (vect_scale_idx, vect_rot_idx, vect_iy, vect_ix)]
    assert post.size > 0 and all(c.size == post.size for c in cols), "sample vectors and posterior differ in length (:451-455)"
    if didx == -1 or vect_didx is None:
        k = int(np.argmax(post))                     # boost_math::get_max: first index of the maximum
    else:
        d = np.asarray(vect_didx).reshape(-1)
        assert d.size == post.size
        idx = np.flatnonzero(d == didx)
        assert idx.size > 0, "maxidx >= 0 (:485)"
        k = int(idx[np.argmax(post[idx])])
    si, ri, iy, ix = (int(c[k]) for c in cols)
    smin, smax, ns = scale_range
    rmin, rmax, nr = rot_range
    scale = smin if smin == smax else smin + (smax - smin) / ns * (0.5 + si)
    rot = rmin if rmin == rmax else rmin + (rmax - rmin) / nr * (0.5 + ri)
    return np.array([si, np.float32(scale), ri, np.float32(rot), ix, iy, np.float32(post[k])], np.float32)


def eval_segments_roi(annolist: Sequence[Annotation], roi_counts: Sequence[int], part_conf_eval: Sequence[PartDef],
                      window_param: Sequence[PartParam], load_best_conf_roi, firstidx: int, lastidx: int, scale: float = 1.0,
                      part_conf: Optional[Sequence[PartDef]] = None, part_conf_type: str = "human_full",
                      rot_range: Tuple[float, float, int] = (-180.0, 180.0, 48)) -> SegmentEval:
    """eval_segments_roi (parteval.cpp:1779-1889): every region of interest of an image has its own pose_est file
    (`load_best_conf_roi(imgidx, roi_idx)` = best_conf of pose_est_imgidx%04d_roi%04d.mat, :495-514); a part counts as
    correct if it matches ANY annotated person of the image ("check all gt rectangles", first match wins).  Unlike
    eval_segments the ground-truth box keeps its y extension here (:1845-1847)."""
    P = len(part_conf_eval)
    model_conf = part_conf_eval if part_conf is None else part_conf
    correct = [0] * P
    seg_correct = seg_total = 0
    for imgidx in range(firstidx, lastidx + 1):
        for roi_idx in range(roi_counts[imgidx]):
            best_conf = np.asarray(load_best_conf_roi(imgidx, roi_idx), np.float32).reshape(-1, 7)
            assert best_conf.shape[0] == len(model_conf)
            dets = convert_eval_bboxes(part_conf_type, [bbox_from_hyp(best_conf[i], window_param[i]) for i in range(len(best_conf))],
                                       [float(r[1]) for r in best_conf], part_conf_eval, model_conf, rot_range)
            assert len(dets) == P
            for pidx in range(P):
                match = False
                for rect in annolist[imgidx].rects:
                    if not rect.points:
                        continue
                    gt = get_part_bbox(rect, part_conf_eval[pidx], scale)
                    if gt is None:
                        raise ValueError("image %d part %d: annotated axis points coincide" % (imgidx, pidx))
                    if is_gt_match_bbox(gt, dets[pidx]):
                        match = True
                        break
                seg_correct += int(match)
                correct[pidx] += int(match)
                seg_total += 1
    ratio = seg_correct / float(seg_total) if seg_total else 0.0
    per_total = seg_total // P if P else 0
    return SegmentEval(ratio, seg_correct, seg_total, correct, [per_total] * P, {})


def eval_segments_experiment(expopt: str, first: Optional[int] = None, numimgs: Optional[int] = None,
                             save_endpoints: bool = True, eval_type: str = "ps", eval_didx: int = -1) -> SegmentEval:
    """`partapp --expopt X --eval_segments` for a finished `--find_obj` run (main.cpp:834-845): reads the expopt, the
    part configuration (part_conf_eval if given), window_param.txt, the test annotation list and the
    pose_est_imgidx%04d.mat files under <log_dir>/<log_subdir>/part_marginals.
    eval_type "ps" (EVAL_TYPE_PS) reads those files; "unaries" (EVAL_TYPE_UNARIES, parteval.cpp:326-383) takes every
    part's estimate from the maximum of its detector score grids (<scoregrid_dir>/imgidx%d-pidx%d-o0-scoregrid.mat);
    "roi" is eval_segments_roi over pose_est_imgidx%04d_roi%04d.mat and the ROI annotation list; "disc_ps"
    (EVAL_TYPE_DISC_PS, :384-493) takes the best of the samples libDiscPS drew and scored (`eval_didx` >= 0: the samples
    and the annotated rectangle of that subject)."""
    import scipy.io
    base = os.path.dirname(os.path.abspath(expopt))
    rel = lambda p: p if os.path.isabs(p) else os.path.normpath(os.path.join(base, p))  # complete_relative_path
    with open(expopt) as f:
        ep = parse_prototext(f.read())
    one = lambda k, d=None: ep[k][-1] if k in ep else d
    log_dir = rel(one("log_dir", "."))
    log_subdir = one("log_subdir") or os.path.splitext(os.path.basename(expopt))[0]   # partapp.cpp:300-306
    class_dir = rel(one("class_dir")) if one("class_dir") else os.path.join(log_dir, log_subdir, "class")
    model_conf = load_part_conf(rel(one("part_conf")))
    conf = load_part_conf(rel(one("part_conf_eval"))) if one("part_conf_eval") else model_conf
    win = load_window_param(os.path.join(class_dir, "window_param.txt"))
    annos: List[Annotation] = []
    anno_dir: List[str] = []                                            # convertFullPath (partapp.cpp:87-107)
    for ds in ep.get("test_dataset", []):
        part = load_annolist(rel(ds))
        annos += part
        anno_dir += [os.path.dirname(rel(ds))] * len(part)
    n = len(annos)
    firstidx = 0 if first is None else first
    lastidx = n - 1 if numimgs is None else min(n - 1, firstidx + numimgs - 1)
    smin, smax, ns = float(one("min_object_scale", 1.0)), float(one("max_object_scale", 1.0)), int(one("num_scale_steps", 1))
    scale = smin if smin == smax else smin + (smax - smin) / ns * 0.5                  # scale_from_index(exp_param, 0)
    hyp_dir = os.path.join(log_dir, log_subdir, "part_marginals")
    load = lambda i: scipy.io.loadmat(os.path.join(hyp_dir, "pose_est_imgidx%04d.mat" % i))["best_conf"]
    rot_range = (float(one("min_part_rotation", -180.0)), float(one("max_part_rotation", 180.0)),
                 int(one("num_rotation_steps", 48)))                                    # ExpParam.proto defaults
    ptype = str(one("part_conf_type", "human_full"))
    if eval_type == "roi":
        roi_annos = load_annolist(rel(one("roi_annolist")))
        assert len(roi_annos) == n, "roi_annolist.size() == m_test_annolist.size() (parteval.cpp:1803)"
        roi_dir = os.path.join(log_dir, log_subdir, "part_marginals_roi")
        load_roi = lambda i, k: scipy.io.loadmat(os.path.join(roi_dir, "pose_est_imgidx%04d_roi%04d.mat" % (i, k)))["best_conf"]
        return eval_segments_roi(annos, [len(a.rects) for a in roi_annos], conf, win, load_roi, firstidx, lastidx, scale,
                                 part_conf=model_conf, part_conf_type=ptype, rot_range=rot_range)
    if eval_type == "disc_ps":
        # parteval.cpp:401-426 and :1210-1238: samples under dai_samples_dir, their posteriors under
        # <log_dir>/<log_subdir>/<part_marginals_samples|test_scoregrid_samples>_post, endpoints saved next to them
        sdir = rel(one("dai_samples_dir"))
        kind = os.path.basename(sdir.rstrip("/"))
        assert kind in ("part_marginals_samples", "test_scoregrid_samples"), "unknown part samples type (:421)"
        post_dir = os.path.join(log_dir, log_subdir, kind + "_post")
        srange = (smin, smax, ns)

        def load(i):                                                     # noqa: F811
            post = scipy.io.loadmat(os.path.join(post_dir, "samples_imgidx%04d_post.mat" % i))
            rows = []
            for p in range(len(model_conf)):
                m = scipy.io.loadmat(os.path.join(sdir, "samples_imgidx%04d" % i, "samples_pidx%d.mat" % p))
                rows.append(disc_ps_best_hyp(m["vect_scale_idx"], m["vect_rot_idx"], m["vect_iy"], m["vect_ix"],
                                             post["samples_post_part%d" % p], srange, rot_range,
                                             m.get("vect_didx") if eval_didx >= 0 else None, eval_didx))
            return np.stack(rows)
        hyp_dir = post_dir
    if eval_type == "unaries":
        assert len(model_conf) in (6, 10, 21), "single-scale settings only (parteval.cpp:344-346)"
        sg_dir = rel(one("scoregrid_dir")) if one("scoregrid_dir") else os.path.join(log_dir, log_subdir, "test_scoregrid")
        sizes = {}

        def load(i):                                                     # noqa: F811 -- the unary source replaces the file reader
            if i not in sizes:
                sizes[i] = _image_size(annos[i].image if os.path.exists(annos[i].image) else os.path.join(anno_dir[i], annos[i].image))
            W, H = sizes[i]
            rows = []
            for p in range(len(win)):
                m = scipy.io.loadmat(os.path.join(sg_dir, "imgidx%d-pidx%d-o0-scoregrid.mat" % (i, p)))
                cg, Ti2, T2g = m["cell_scoregrid"], m["transform_Ti2"].astype(np.float64), m["transform_T2g"].astype(np.float64)
                R = cg.shape[1]
                cells = np.stack([np.asarray(cg[0, r], np.float32) for r in range(R)])
                Tig = np.stack([Ti2[0, r] @ T2g[0, r] for r in range(R)])
                rows.append(unary_best_hyp(load_score_grid_direct(cells, Tig, H, W), 0, scale, rot_range))
            return np.stack(rows)
        hyp_dir = sg_dir
    return eval_segments(annos, conf, win, load, firstidx, lastidx, scale, eval_didx=eval_didx,
                         save_dir=os.path.join(hyp_dir, "seg_endpoints") if save_endpoints else None,
                         part_conf=model_conf, part_conf_type=ptype, rot_range=rot_range)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="PCP evaluation of a --find_obj run (the reference's --eval_segments)")
    ap.add_argument("--expopt", required=True)
    ap.add_argument("--first", type=int)
    ap.add_argument("--numimgs", type=int)
    ap.add_argument("--eval-type", default="ps", choices=["ps", "unaries", "roi", "disc_ps"])
    ap.add_argument("--eval-didx", type=int, default=-1, help="subject (annotated rectangle) to evaluate, disc_ps only")
    a = ap.parse_args()
    r = eval_segments_experiment(a.expopt, a.first, a.numimgs, eval_type=a.eval_type, eval_didx=a.eval_didx)
    print("seg_correct: %d\nseg_total: %d\nratio: %g" % (r.seg_correct, r.seg_total, r.ratio))
    for i, (c, t) in enumerate(zip(r.per_part_correct, r.per_part_total)):
        print("part: %d, correct: %d, total: %d, ratio: %g" % (i, c, t, c / float(t if t else 1)))
