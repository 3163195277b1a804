"""One message through the level-batched route (fused x+y Gaussian) and through the two-pass route of round 1, with a
map of where they differ -- run on the GPU box when a parity test of the fused kernel fails:

    python tools/gpu_debug_fused.py [H W R]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from partapp_b200 import ExpParam, PsContext, synth  # noqa: E402

H, W, R = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (200, 152, 8)
ep = ExpParam(num_rotation_steps=R)
rng = np.random.default_rng(0)
child = (rng.standard_normal((R, H, W)) * 3 - 5).astype(np.float32)
j = synth.make_joints(10, seed=7, max_offset=min(H, W) / 10.0, sigma_range=(2.0, min(H, W) / 20.0))[1]
args = (j.offset_p, j.offset_c, j.C, -j.rot_mean, j.rot_sigma, 1.0, False)


def run(env):
    for k in ("PSINFER_NO_BATCH", "PSINFER_NO_GRAPH", "PSINFER_ALL_TILES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    with PsContext(ep, synth.part_conf(2), H, W) as ctx:
        out = ctx.message(child, *args)
        n = ctx.launch_count()
    return out, n


old, n_old = run({"PSINFER_NO_BATCH": "1"})
new, n_new = run({})
alltiles, _ = run({"PSINFER_ALL_TILES": "1"})
print("launches: two-pass %d, fused %d" % (n_old, n_new))
for name, got in (("fused", new), ("fused, no work lists", alltiles)):
    neq = got != old
    print("%s: %d of %d cells differ" % (name, neq.sum(), neq.size))
    if neq.any():
        r, y, x = np.argwhere(neq)[0]
        print("  first at r=%d y=%d x=%d: %r vs %r" % (r, y, x, got[r, y, x], old[r, y, x]))
        print("  per slice:", neq.reshape(R, -1).sum(1).tolist())
        rows = neq.any(axis=(0, 2))
        cols = neq.any(axis=(0, 1))
        print("  rows %d..%d, cols %d..%d" % (np.flatnonzero(rows)[0], np.flatnonzero(rows)[-1], np.flatnonzero(cols)[0],
                                              np.flatnonzero(cols)[-1]))
        d = np.abs(got.astype(np.float64) - old)[neq]
        print("  |diff| max %g median %g; NaN in fused: %d" % (d.max(), np.median(d), int(np.isnan(got).sum())))
