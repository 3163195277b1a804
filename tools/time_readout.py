"""One-off timing of the readout options at the benchmark size (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from partapp_b200 import ExpParam, PsContext, synth

ep = ExpParam(num_rotation_steps=24)
P, H, W = 10, 600, 400
cells, Tig = synth.compact_scores(ep, H, W, P, 0)
with PsContext(ep, synth.part_conf(P), H, W) as ctx:
    ctx.set_joints(synth.make_joints(P, seed=7))
    for lm, rh in ((False, False), (False, True), (True, True)):
        for rep in range(2):
            for p in range(P):
                ctx.set_unary_compact(p, 0, cells[p, 0], Tig)
            ctx.synchronize()
            t0 = time.perf_counter()
            ctx.infer(sparse=True, local_max=lm, root_hyps=rh)
            ctx.best_conf()
            t1 = time.perf_counter()
        n = [len(ctx.part_hyps(p)) for p in range(P)] if lm else None
        nr = len(ctx.root_hyps()) if rh else None
        print("local_max=%s root_hyps=%s: %.1f ms  part hyps %s root hyps %s" % (lm, rh, (t1 - t0) * 1e3, n, nr))
