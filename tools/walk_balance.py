"""Host-only: how evenly do the walks of the fused Gaussian kernel load the four warp schedulers of an SM?
For every message of the bench model (cfg-2 synthetic joints, both directions share the covariance) walks the plan
(`ps_plan_walks`) and sums, per step, the taps of each filter warp in the x phase (mask bit) and the y phase (group
count), then compares the busiest scheduler (warps w and w + 4 share one) with the mean."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
from partapp_b200 import ExpParam, capi, synth  # noqa: E402
from partapp_b200.objectdetect import PartConf, make_config  # noqa: E402


def walks_of(lib, cfg, Cm, scale=1.0):
    dims = (ctypes.c_int * 7)()
    cap, mcap = 4096, 1 << 16
    wl = (ctypes.c_int * (4 * cap))()
    ml = (ctypes.c_ubyte * mcap)()
    nm = ctypes.c_int(0)
    Cc = (ctypes.c_double * 4)(*np.asarray(Cm, np.float64).ravel())
    rc = lib.ps_plan_walks(ctypes.byref(cfg), Cc, scale, dims, wl, cap, ml, mcap, ctypes.byref(nm))
    assert rc == 0, rc
    EH, EW, nx, ny, halo, lag, nw = [int(v) for v in dims]
    return EH, EW, nx, ny, halo, lag, np.array(wl[:4 * nw], np.int64).reshape(-1, 4), np.array(ml[:nm.value], np.uint8)


def main():
    H, W, R, P = 600, 400, 24, 10
    lib = capi.load_library()
    cfg = make_config(ExpParam(num_rotation_steps=R), PartConf([True] * P, [False] * P, [True] * P), H, W)
    joints = synth.make_joints(P)
    tot_x = tot_y = 0
    mx_x = mx_y = 0.0   # sum over steps of (max scheduler load * 4)
    blk = 0.0           # sum over steps of per-block critical path: max warp * 8
    steps = 0
    for j in joints:
        EH, EW, nx, ny, halo, lag, walks, masks = walks_of(lib, cfg, np.asarray(j.C))
        lx, ly = 2 * nx + 1, 2 * ny + 1
        for strip, row0, ng, moff in walks:
            nxb = (ng + 7) // 8 + lag
            for i in range(nxb):
                m = int(masks[moff + i])
                xw = np.array([(m >> w) & 1 for w in range(8)]) * lx
                ob = i - lag
                yw = np.zeros(8)
                if ob >= 0:
                    act = min(8, ng - 8 * ob)
                    yw[:max(0, act)] = ly
                    for w in range(8):
                        if row0 + 64 * ob + 8 * w >= EH:
                            yw[w] = 0
                sx = xw[:4] + xw[4:]
                sy = yw[:4] + yw[4:]
                tot_x += xw.sum(); tot_y += yw.sum()
                mx_x += 4 * sx.max(); mx_y += 4 * sy.max()
                steps += 1
        print("joint %d->%d  EH %d EW %d  taps %d/%d lag %d walks %d" % (j.child_idx, j.parent_idx, EH, EW, lx, ly, lag, len(walks)))
    print("x phase: useful warp-taps %.3g, 4*max-scheduler %.3g -> balance %.3f" % (tot_x, mx_x, tot_x / mx_x))
    print("y phase: useful warp-taps %.3g, 4*max-scheduler %.3g -> balance %.3f" % (tot_y, mx_y, tot_y / mx_y))
    print("both: %.3f   steps %d" % ((tot_x + tot_y) / (mx_x + mx_y), steps))


if __name__ == "__main__":
    main()
