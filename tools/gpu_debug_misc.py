"""Diagnostics for the two failures of the first round-2 GPU run: the POS_GAUSSIAN driver on the branching case and the
multi-worker CLI."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from partapp_b200 import PsContext  # noqa: E402
from tests.golden import ref_pos_driver_cases as C  # noqa: E402
from tests.make_experiment import make  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "ref_pos_driver.npz"))
for name, (ep, pc, joints, un, sparse) in C.cases().items():
    want = g[name + "/root_post"]
    P, S, R, H, W = un.shape
    with PsContext(ep, pc, H, W) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            for s in range(S):
                ctx.set_unary(p, s, un[p, s])
        ctx.infer(sparse=sparse, root_hyps=True)
        got = ctx.root_posterior()
    fin = np.isfinite(want) & np.isfinite(got)
    d = np.abs(got[fin].astype(np.float64) - want[fin])
    print(name, "cells differing", int((got != want).sum()), "of", got.size, "max abs", d.max() if d.size else 0,
          "max rel", (d / np.maximum(np.abs(want[fin]), 1e-30)).max() if d.size else 0, "inf mismatch",
          int((np.isinf(got) != np.isinf(want)).sum()))
    print("  want[0,:2,:6]", want[0, :2, :6], "\n  got       ", got[0, :2, :6])

cli = os.path.join(ROOT, "partapp_b200", "psinfer_partapp")
subprocess.run(["make", "-C", os.path.join(ROOT, "partapp_b200", "csrc", "host")], check=True, capture_output=True)
tmp = tempfile.mkdtemp()
info = make(os.path.join(tmp, "b"), num_images=8, P=4, R=8, H=48, W=40)
for extra in (["--gpus", "1", "--contexts", "1"], ["--gpus", "1", "--contexts", "3"], ["--distribute", "--ncpu", "4", "--batch_num", "1"]):
    r = subprocess.run([cli, "--expopt", info["expopt"], "--find_obj"] + extra, capture_output=True, text=True)
    print(extra, "rc", r.returncode, "|", r.stdout.strip()[-200:], "|", r.stderr.strip()[-400:])
    print("  files:", sorted(os.listdir(os.path.join(info["base"], "part_marginals")))[:10])
