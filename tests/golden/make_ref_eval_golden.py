"""Records what the reference's own evaluator code (oracle/_ref/libps_ref_eval.so) returns for the cases of
tests/test_eval_vs_ref.py into tests/golden/ref_eval.npz.  Run in the container that has /root/reference:

    make -C oracle ref && python tests/golden/make_ref_eval_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refcore  # noqa: E402
from tests import test_eval_vs_ref as T  # noqa: E402

assert refcore.eval_available(), "build oracle/_ref first (make -C oracle ref)"
out = {}
mc = T.match_cases()
out["endpoints"] = np.stack([refcore.eval_endpoints(gt) for gt, _, _ in mc])
out["match"] = np.array([refcore.eval_is_gt_match(gt, det, f) for gt, det, f in mc])
gc = T.merge_cases()
out["merge2"] = np.stack([refcore.eval_bbox_merge(2, b[:2]) for b in gc])
out["merge4"] = np.stack([refcore.eval_bbox_merge(4, b) for b in gc])
out["merge3"] = np.stack([refcore.eval_bbox_merge(3, b[:2]) for b in gc])
codes, boxes = [], []
for pts, pos, fr, to, f5, scale in T.part_bbox_cases():
    r = refcore.eval_get_part_bbox(pts, pos, fr, to, f5, scale)
    codes.append(-1 if r is False else (0 if r is None else 1))
    boxes.append(np.zeros(10) if r is None or r is False else r)
out["part_bbox_code"] = np.array(codes)
out["part_bbox"] = np.stack(boxes)
for k, case in enumerate(T.helper_cases()):
    out["helper_%d" % k] = T.ref_helper(case)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_eval.npz"), **out)
print("wrote tests/golden/ref_eval.npz:", {k: v.shape for k, v in out.items() if not k.startswith("helper")})
