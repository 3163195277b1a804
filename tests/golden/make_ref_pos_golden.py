"""Records the reference's own mergeRotations + computeRootPosterior (oracle/_ref/libps_ref_drivers.so =
objectdetect_findpos.cpp compiled unmodified) on the cases of ref_pos_driver_cases.py into tests/golden/ref_pos_driver.npz.

    make -C oracle ref && python tests/golden/make_ref_pos_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refcore  # noqa: E402
from tests.golden import ref_pos_driver_cases as C  # noqa: E402

out = {}
for name, (ep, pc, joints, un, sparse) in C.cases().items():
    merged, rp = refcore.root_posterior_pos(ep, pc, joints, np.ascontiguousarray(un), sparse)
    out[name + "/merged"] = merged
    out[name + "/root_post"] = rp
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_pos_driver.npz"), **out)
print({k: v.shape for k, v in out.items()})
