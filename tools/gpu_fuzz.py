"""Runs ON THE GPU BOX: randomised whole inferences through the C ABI against the CPU oracle -- grid sizes (multiples of
four and not: the latter take the one-message route), rotation counts, tree sizes, one or two scales, full / diagonal /
mixed covariances, filter lengths from a few taps to beyond a 64-row block, sparse and dense unaries, upright roots,
border strips, compact ingest repeated on one context.  Every marginal cell, the root posterior and the argmax records
must be identical.   usage: python tools/gpu_fuzz.py [cases] [seed] [max grid side]   -> gpurun_out/fuzz_<seed>.txt"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (test infrastructure: the checker)
from partapp_b200 import ExpParam, PsContext, synth  # noqa: E402


def main(cases=None, seed=None, maxdim=None):
    cases = cases if cases is not None else int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = seed if seed is not None else int(sys.argv[2]) if len(sys.argv) > 2 else 1
    maxdim = maxdim if maxdim is not None else int(sys.argv[3]) if len(sys.argv) > 3 else 150
    rng = np.random.default_rng(seed)
    bad = 0
    lines = []
    t_all = time.time()
    for k in range(cases):
        R = int(rng.choice([8, 12, 24, 48]))
        S = int(rng.choice([1, 1, 2]))
        P = int(rng.integers(2, 7))
        H = int(rng.integers(24, maxdim))
        W = int(rng.integers(24, maxdim))
        if rng.random() < 0.6:
            W = (W + 3) // 4 * 4
        smax = float(rng.choice([2.5, 5.0, 9.0, 14.0]))
        diag = rng.random() < 0.15
        ep = ExpParam(num_rotation_steps=R, num_scale_steps=S, min_object_scale=1.0 if S == 1 else 0.85,
                      max_object_scale=1.0 if S == 1 else 1.15, roi_save_num_samples=8,
                      strip_border_detections=float(rng.choice([0.0, 0.0, 0.05])))
        pc = synth.part_conf(P, upright_root=bool(rng.random() < 0.25))
        joints = synth.make_joints(P, seed=int(rng.integers(1, 10 ** 6)), diagonal=diag, max_offset=float(rng.uniform(2, 12)),
                                   sigma_range=(1.0, smax))
        if not diag and rng.random() < 0.3 and len(joints) > 1:   # mixed: one diagonal joint among full ones
            dj = synth.make_joints(P, seed=int(rng.integers(1, 10 ** 6)), diagonal=True, max_offset=5.0, sigma_range=(1.0, smax))
            joints[0].C = dj[0].C
        sparse = bool(rng.random() < 0.7)
        compact = sparse and rng.random() < 0.5
        tag = "case %d R%d S%d P%d %dx%d sig<=%.1f diag%d sparse%d compact%d strip%.2f upright%d" % (
            k, R, S, P, H, W, smax, diag, sparse, compact, ep.strip_border_detections, int(pc.is_upright[0] or any(pc.is_upright)))
        try:
            with PsContext(ep, pc, H, W, keep_all_scales=True) as ctx:
                ctx.set_joints(joints)
                for rep in range(2 if compact else 1):          # the second image on one lattice skips the fill
                    if compact:
                        cells, Tig = synth.compact_scores(ep, H, W, P, 10 * k + rep)
                        cells = cells.copy()
                        cells[rng.random(cells.shape) < 0.3] = 0.0
                        ps_ = [p for p in range(P) for s in range(S)]
                        ss = [s for p in range(P) for s in range(S)]
                        ctx.set_unaries_compact(ps_, ss, [cells[p, s] for p, s in zip(ps_, ss)], Tig)
                        un = np.stack([np.stack([oracle.prepare_unary(oracle.load_score_grid(cells[p, s], Tig, H, W))
                                                 for s in range(S)]) for p in range(P)])
                    else:
                        raw = synth.raw_scores(ep, H, W, P, k, stride=4 if sparse else 1)
                        un = oracle.prepare_unary(raw)
                        for p in range(P):
                            for s in range(S):
                                ctx.set_unary(p, s, un[p, s])
                    want = oracle.infer(ep, pc, joints, np.ascontiguousarray(un.copy()), sparse=sparse)
                    ctx.infer(sparse=sparse, keep_unaries=bool(compact and rep == 0 and rng.random() < 0.5))
                    ok = np.array_equal(ctx.best_conf(), want["best_conf"])
                    diff = 0
                    for s in range(S):
                        for p in range(P):
                            g = ctx.marginal(p, s)
                            w = want["marginals"][s, p]
                            diff += int(((g != w) & ~(np.isnan(g) & np.isnan(w))).sum())
                    rp = ctx.root_posterior()
                    rdiff = int(((rp != want["root_post"]) & ~(np.isnan(rp) & np.isnan(want["root_post"]))).sum())
                    status = "ok" if ok and diff == 0 and rdiff == 0 else "MISMATCH argmax_equal=%s marginal_cells=%d root_cells=%d" % (ok, diff, rdiff)
                    if status != "ok":
                        bad += 1
                    lines.append("%s rep%d: %s" % (tag, rep, status))
        except Exception as e:  # noqa: BLE001
            bad += 1
            lines.append("%s: EXCEPTION %r" % (tag, e))
        if not lines[-1].endswith(': ok'):
            print(lines[-1], flush=True)
    lines.append("%d cases, %d bad, %.0f s" % (cases, bad, time.time() - t_all))
    print(lines[-1])
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/fuzz_%d.txt" % seed, "w") as f:
        f.write("\n".join(lines) + "\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
