"""Counts, over every fp32 input of the path's domain, where the device exp/log differ from the host glibc the
oracle uses ((float)exp((double)x), (float)log((double)x)).  Run under gpurun; writes profiles/r01_libm_exhaustive.txt.
"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from partapp_b200 import ExpParam, PsContext, synth  # noqa: E402

CH = 1 << 24


def f2b(x):
    return int(np.float32(x).view(np.uint32))


def sweep(ctx, op, lo_bits, hi_bits, pool):
    bad = tested = 0
    futs = []
    L = oracle.lib()
    for first in range(lo_bits, hi_bits, CH):
        n = min(CH, hi_bits - first)
        dev = ctx.eval_math(op, first, n)
        futs.append(pool.submit(L.orc_compare_math, op, first, n, dev.ctypes.data_as(oracle._fp)))
        futs[-1].keep = dev
        tested += n
        if len(futs) > 16:
            bad += futs.pop(0).result()
    for f in futs:
        bad += f.result()
    return bad, tested


def main():
    out = []
    t0 = time.time()
    with PsContext(ExpParam(num_rotation_steps=8), synth.part_conf(2), 8, 8) as ctx, ThreadPoolExecutor(os.cpu_count()) as pool:
        # exp: x in [-104, 0] (the message path: x = value - max <= 0) and (0, 88] (root marginal)
        b1, n1 = sweep(ctx, 0, f2b(-0.0), f2b(-104.0) + 1, pool)     # negative floats: bits increase with |x|
        b2, n2 = sweep(ctx, 0, 0, f2b(88.0) + 1, pool)
        out.append("exp  x in [-104, -0]: %d inputs, %d differ from glibc (%.2e)" % (n1, b1, b1 / n1))
        out.append("exp  x in [+0, 88]  : %d inputs, %d differ from glibc (%.2e)" % (n2, b2, b2 / n2))
        b3, n3 = sweep(ctx, 1, 0, f2b(np.inf), pool)
        out.append("log  x in [+0, max] : %d inputs, %d differ from glibc (%.2e)" % (n3, b3, b3 / n3))
    out.append("elapsed %.0f s on %d host threads" % (time.time() - t0, os.cpu_count()))
    text = ("Device exp/log of the path vs the host libm of the oracle, every fp32 input (tools/libm_exhaustive.py).\n"
            "Both sides evaluate in fp64 (< 1 ulp) and narrow; they can only differ when the fp64 values straddle an fp32\n"
            "rounding boundary.\n" + "\n".join(out) + "\n")
    print(text)
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/r01_libm_exhaustive.txt", "w").write(text)


if __name__ == "__main__":
    main()
