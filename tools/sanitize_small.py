"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): one inference with full
covariances, local maxima and root hypotheses, plus a compact ingest and a standalone message."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from partapp_b200 import ExpParam, PsContext, synth

ep = ExpParam(num_rotation_steps=8, roi_save_num_samples=10)
P, H, W = 4, 40, 36
cells, Tig = synth.compact_scores(ep, H, W, P, 1)
joints = synth.make_joints(P, seed=5, max_offset=6, sigma_range=(1.5, 3))
with PsContext(ep, synth.part_conf(P), H, W) as ctx:
    ctx.set_joints(joints)
    for p in range(P):
        ctx.set_unary_compact(p, 0, cells[p, 0], Tig)
    ctx.infer(sparse=True, local_max=True, root_hyps=True)
    print(ctx.best_conf()[:, 2:6])
    j = joints[0]
    out = ctx.message(ctx.get_unary(0, 0), j.offset_c, j.offset_p, [[4.0, 0], [0, 3.0]], 0.2, 0.5, 1.0, True)
    print(float(out.max()))
    # the other entry points: raw ingest + detection maxima + in-place log, the POS_GAUSSIAN message, DPM adds
    ctx.set_unary_compact_raw(1, 0, cells[1, 0], Tig)
    print(len(ctx.unary_local_max(1, 0, 10)))
    ctx.log_unary(1, 0)
    par, ch = ctx.pos_message(ctx.get_unary(1, 0), (3.5, -2.0), [[5.0, 1.0], [1.0, 3.0]], 1.0, False)
    print(float(par.max()))
    ctx.add_unary_grid(0, np.random.default_rng(0).random((H, W), dtype=np.float32), 1, 0.5)
# rotation counts with a specialised rotation filter (R = 24) in both arithmetic modes, bilinear ingest
ep24 = ExpParam(num_rotation_steps=24, roi_save_num_samples=10, interpolate=True)
cells, Tig = synth.compact_scores(ep24, H, W, P, 2)
for fast in (False, True):
    with PsContext(ep24, synth.part_conf(P), H, W, fast_math=fast) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            ctx.set_unary_compact(p, 0, cells[p, 0], Tig)
        ctx.infer(sparse=True)
        print(fast, ctx.best_conf()[:, 2:6].tolist())
# a grid of several 64-cell strips with an oblique covariance: the Gaussian work lists hold partial tiles and leave
# most of the eigen-frame bounding box out
ep4 = ExpParam(num_rotation_steps=4)
H2, W2 = 150, 100
rng = np.random.default_rng(3)
child = (rng.standard_normal((4, H2, W2)) * 2 - 3).astype(np.float32)
with PsContext(ep4, synth.part_conf(2), H2, W2) as ctx:
    for sparse in (True, False):
        out = ctx.message(child, (5.0, -3.0), (-4.0, 6.0), [[20.0, 9.0], [9.0, 14.0]], 0.2, 0.5, 1.0, sparse)
        print(sparse, float(out.max()))
# round 2: one-call ingest, fused table add, the level-batched route replayed from its CUDA graph (second and third
# inference with the same joints), a 10-part tree (levels of 5 and 4 messages), the POS_GAUSSIAN model
ep8 = ExpParam(num_rotation_steps=8, roi_save_num_samples=10)
P10, H3, W3 = 10, 72, 64
cells, Tig = synth.compact_scores(ep8, H3, W3, P10, 4)
joints10 = synth.make_joints(P10, seed=7, max_offset=8, sigma_range=(2, 6))
with PsContext(ep8, synth.part_conf(P10), H3, W3) as ctx:
    ctx.set_joints(joints10)
    for rep in range(3):
        ctx.set_unaries_compact(list(range(P10)), [0] * P10, [cells[p, 0] for p in range(P10)], Tig)
        rot = ctx.rot_score_table(0.3, 0.2)
        pos = ctx.pos_score_table(2.0, -3.0, 90.0, 60.0, W3 / 2, H3 / 2)
        ctx.add_unary_tables(3, [rot, pos], [0, 1], [0.8, 0.6])
        ctx.infer(sparse=True)
        print(rep, ctx.best_conf()[:3, 2:6].tolist(), ctx.launch_count())
for j in joints:
    j.type = 1
with PsContext(ep, synth.part_conf(P), H, W) as ctx:
    ctx.set_joints(joints)
    cells, Tig = synth.compact_scores(ep, H, W, P, 1)
    for p in range(P):
        ctx.set_unary_compact(p, 0, cells[p, 0], Tig)
    ctx.infer(sparse=True, root_hyps=True)
    print("pos model", float(np.nanmax(ctx.root_posterior())), len(ctx.root_hyps()))
