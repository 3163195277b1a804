"""PCP matching rule of the reference's evaluator, for the "PCP-identical part estimates" check of BASELINE.json.

Restates (host side, numpy -- this is evaluation, not the hot path):
  bbox_from_pos       reference src/libs/libPartApp/partapp.cpp:59-81
  get_bbox_endpoints  src/libs/libPartEval/parteval.cpp:45-68 (the non-endpoint branch)
  is_gt_match         src/libs/libPartEval/parteval.cpp:79-110 (match_x_axis == false): both endpoint distances
                      strictly below factor * ground-truth segment length, factor 0.5 for PCP.
A part estimate is a `best_conf` row (PartHyp::toVect, objectdetect.h:139-160):
[scaleidx, scale, rotidx, rot_deg, x, y, score].
"""
from dataclasses import dataclass
from typing import Sequence

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
