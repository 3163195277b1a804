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
    scale, rot = float(row[1]), float(row[3]) / 180.0 * np.pi      # PartHyp::getPartBBox, objectdetect.h:107-113
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
# The "human_full_joints" part merging (parteval.cpp:527-570) belongs to another part_conf_type and is not restated.


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
    scale, rot = float(row[1]), float(row[3]) / 180 * math.pi
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
                  save_dir: Optional[str] = None) -> SegmentEval:
    """EVAL_TYPE_PS branch of eval_segments (parteval.cpp:1162-1345).  `load_best_conf(imgidx)` returns the [P][7]
    `best_conf` of pose_est_imgidx%04d.mat; `scale` is scale_from_index(exp_param, 0) (:1207-1208).
    With `save_dir`, the per-image endpoint matrices are written like the reference's seg_endpoints directory."""
    P = len(part_conf_eval)
    correct, total = [0] * P, [0] * P
    seg_correct = seg_total = 0
    endpoints = {}
    for imgidx in range(firstidx, lastidx + 1):
        best_conf = np.asarray(load_best_conf(imgidx), np.float32).reshape(-1, 7)
        assert best_conf.shape[0] == P, "eval_bbox.size() == part_conf_eval.part_size() (parteval.cpp:1250)"
        ep = np.zeros((P + 1, 6))
        n_ok = n_seg = 0
        rects = annolist[imgidx].rects
        assert len(rects) > 0
        rect = rects[eval_didx if eval_didx >= 0 else 0]
        for pidx in range(P):
            det = bbox_from_hyp(best_conf[pidx], window_param[pidx])
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


def eval_segments_experiment(expopt: str, first: Optional[int] = None, numimgs: Optional[int] = None,
                             save_endpoints: bool = True) -> SegmentEval:
    """`partapp --expopt X --eval_segments` for a finished `--find_obj` run (main.cpp:834-845): reads the expopt, the
    part configuration (part_conf_eval if given), window_param.txt, the test annotation list and the
    pose_est_imgidx%04d.mat files under <log_dir>/<log_subdir>/part_marginals."""
    import scipy.io
    base = os.path.dirname(os.path.abspath(expopt))
    rel = lambda p: p if os.path.isabs(p) else os.path.normpath(os.path.join(base, p))  # complete_relative_path
    with open(expopt) as f:
        ep = parse_prototext(f.read())
    one = lambda k, d=None: ep[k][-1] if k in ep else d
    log_dir = rel(one("log_dir", "."))
    log_subdir = one("log_subdir") or os.path.splitext(os.path.basename(expopt))[0]   # partapp.cpp:300-306
    class_dir = rel(one("class_dir")) if one("class_dir") else os.path.join(log_dir, log_subdir, "class")
    conf = load_part_conf(rel(one("part_conf_eval") or one("part_conf")))
    win = load_window_param(os.path.join(class_dir, "window_param.txt"))
    annos: List[Annotation] = []
    for ds in ep.get("test_dataset", []):
        annos += load_annolist(rel(ds))
    n = len(annos)
    firstidx = 0 if first is None else first
    lastidx = n - 1 if numimgs is None else min(n - 1, firstidx + numimgs - 1)
    smin, smax, ns = float(one("min_object_scale", 1.0)), float(one("max_object_scale", 1.0)), int(one("num_scale_steps", 1))
    scale = smin if smin == smax else smin + (smax - smin) / ns * 0.5                  # scale_from_index(exp_param, 0)
    hyp_dir = os.path.join(log_dir, log_subdir, "part_marginals")
    load = lambda i: scipy.io.loadmat(os.path.join(hyp_dir, "pose_est_imgidx%04d.mat" % i))["best_conf"]
    return eval_segments(annos, conf, win, load, firstidx, lastidx, scale,
                         save_dir=os.path.join(hyp_dir, "seg_endpoints") if save_endpoints else None)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="PCP evaluation of a --find_obj run (the reference's --eval_segments)")
    ap.add_argument("--expopt", required=True)
    ap.add_argument("--first", type=int)
    ap.add_argument("--numimgs", type=int)
    a = ap.parse_args()
    r = eval_segments_experiment(a.expopt, a.first, a.numimgs)
    print("seg_correct: %d\nseg_total: %d\nratio: %g" % (r.seg_correct, r.seg_total, r.ratio))
    for i, (c, t) in enumerate(zip(r.per_part_correct, r.per_part_total)):
        print("part: %d, correct: %d, total: %d, ratio: %g" % (i, c, t, c / float(t if t else 1)))
