"""GPU parity at the sizes bench.py measures (BASELINE.json configs[1], configs[3], configs[4]).

Every comparison demands the same bits as the CPU oracle (which tests/test_oracle_vs_ref.py pins to the reference's own
code): whole inferences at 600x400x24 -- every marginal cell, the root posterior, best_conf -- a 1000x1000x48 message
and a two-scale inference at that size, a 22-part tree, and the fast_math mode's argmax identity counted over 50
benchmark-size images instead of assumed.  Results that are counts are also written to gpurun_out/ as JSON.
"""
import json
import os

import numpy as np
import pytest

import oracle
from partapp_b200 import ExpParam, PsContext, synth
from partapp_b200 import objectdetect as od

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LZ = np.float32(-1e6)


def _same(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, what
    neq = (got != want) & ~(np.isnan(got) & np.isnan(want))
    assert not neq.any(), "%s: %d of %d cells differ (first at %s)" % (what, neq.sum(), neq.size,
                                                                     np.argwhere(neq)[0] if neq.any() else None)


def _report(name, obj):
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "w") as f:
            json.dump(obj, f, indent=1)
    except OSError:
        pass


def test_cfg2_whole_inference_matches_oracle():
    """The bench workload itself: 10-part tree, R = 24, 600 x 400, joints of seed 7, unaries through the compact
    ingest -- all ten marginals, the root posterior and the argmax records, bit for bit."""
    ep = ExpParam(num_rotation_steps=24, roi_save_num_samples=20)
    P, H, W = 10, 600, 400
    pc = synth.part_conf(P)
    joints = synth.make_joints(P, seed=7)
    cells, Tig = synth.compact_scores(ep, H, W, P, 0)
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 0))
    want = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
    with PsContext(ep, pc, H, W) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            ctx.set_unary_compact(p, 0, cells[p, 0], Tig)
        ctx.infer(sparse=True)
        best = ctx.best_conf()
        assert np.array_equal(best, want["best_conf"]), "best_conf differs"
        for p in range(P):
            _same(ctx.marginal(p), want["marginals"][0, p], "cfg-2 marginal of part %d" % p)
        _same(ctx.root_posterior(), want["root_post"], "cfg-2 root posterior")
        # the level-batched schedule: a second image (plans and scatter maps cached) costs at most 60 launches
        n0 = ctx.launch_count()
        for p in range(P):
            ctx.set_unary_compact(p, 0, cells[p, 0], Tig)
        ctx.infer(sparse=True)
        assert np.array_equal(ctx.best_conf(), want["best_conf"])
        assert ctx.launch_count() - n0 <= 60, "%d launches for one image" % (ctx.launch_count() - n0)


def test_cfg5_size_message_matches_oracle():
    """One message on the stress grid (R = 48, 1000 x 1000) at scale 1.2: upward (sparse) and downward."""
    ep = ExpParam(num_rotation_steps=48, num_scale_steps=5, min_object_scale=0.8, max_object_scale=1.2)
    H, W = 1000, 1000
    child = oracle.prepare_unary(synth.raw_scores(ExpParam(num_rotation_steps=48), H, W, 1, 5)[0, 0])
    j = synth.make_joints(10, seed=7)[2]
    with PsContext(ep, synth.part_conf(2), H, W) as ctx:
        up = ctx.message(child, j.offset_c, j.offset_p, j.C, j.rot_mean, j.rot_sigma, 1.2, True)
        want_up = oracle.message(ep, child, j.offset_c, j.offset_p, j.C, j.rot_mean, j.rot_sigma, 1.2, True)
        _same(up, want_up, "1000x1000x48 upward message")
        down = ctx.message(want_up, j.offset_p, j.offset_c, j.C, -j.rot_mean, j.rot_sigma, 1.2, False)
        want_down = oracle.message(ep, want_up, j.offset_p, j.offset_c, j.C, -j.rot_mean, j.rot_sigma, 1.2, False)
        _same(down, want_down, "1000x1000x48 downward message")


def test_cfg5_size_two_scale_inference_matches_oracle():
    """A 3-part chain over two scales on the stress grid: marginals of both scales, root posterior, argmax."""
    ep = ExpParam(num_rotation_steps=48, num_scale_steps=2, min_object_scale=0.9, max_object_scale=1.1,
                  roi_save_num_samples=5)
    P, H, W = 3, 1000, 1000
    pc = synth.part_conf(P)
    joints = synth.make_joints(P, seed=11, max_offset=40, sigma_range=(4, 12))
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 3))
    want = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
    with PsContext(ep, pc, H, W, keep_all_scales=True) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            for s in range(2):
                ctx.set_unary(p, s, un[p, s])
        ctx.infer(sparse=True)
        assert np.array_equal(ctx.best_conf(), want["best_conf"])
        for s in range(2):
            for p in range(P):
                _same(ctx.marginal(p, s), want["marginals"][s, p], "1000x1000x48 marginal s%d p%d" % (s, p))
        _same(ctx.root_posterior(), want["root_post"], "1000x1000x48 root posterior")


def test_22_part_tree_medium_grid_matches_oracle():
    """configs[3] tree (22 parts, chains of five) on a 208 x 152 grid with R = 24: level batches of different widths."""
    ep = ExpParam(num_rotation_steps=24, roi_save_num_samples=5)
    P, H, W = 22, 208, 152
    pc = synth.part_conf(P)
    joints = synth.make_joints(P, seed=9, max_offset=20, sigma_range=(2, 8))
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 4))
    want = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
    with PsContext(ep, pc, H, W) as ctx:
        res = od.computeRootPosteriorRot(ctx, [[un[p, 0]] for p in range(P)], joints, True, write_back_masked=False)
        assert np.array_equal(res.best_conf, want["best_conf"])
        for p in range(P):
            _same(ctx.marginal(p), want["marginals"][0, p], "22-part marginal of part %d" % p)
        _same(res.root_part_posterior, want["root_post"], "22-part root posterior")


def test_conditioned_model_tables_in_one_pass_match_oracle():
    """configs[3] step at a medium size: joints swapped per image plus rotation, position and torso-prior tables
    (findrot.cpp:913-949) applied by the fused multi-table add -- same bits as one add per table and as the oracle."""
    ep = ExpParam(num_rotation_steps=24, roi_save_num_samples=5)
    P, H, W = 10, 120, 96
    pc = synth.part_conf(P)
    _, root = synth.tree(P)
    joints = synth.make_joints(P, seed=21, max_offset=12, sigma_range=(2, 6), type_id=3)
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 6))
    with PsContext(ep, pc, H, W) as ctx:
        rng = np.random.default_rng(5)
        tables = {}
        for p in range(P):
            rot = ctx.rot_score_table(rng.uniform(-1, 1), rng.uniform(0.05, 0.5))
            if p == root:
                pos = ctx.torso_prior_table(rng.uniform(-5, 5), rng.uniform(-5, 5), 900.0, 1600.0, 0.7)
                tables[p] = ([rot, pos], [0, 2], [0.8, 1.0])
            else:
                pos = ctx.pos_score_table(rng.uniform(-20, 20), rng.uniform(-20, 20), 400.0, 900.0, W / 2, H / 2)
                tables[p] = ([rot, pos], [0, 1], [0.8, 0.6])
        # one add per table (the pinned path) ...
        ctx.set_joints(joints)
        for p in range(P):
            ctx.set_unary(p, 0, un[p, 0])
            for t, k, w in zip(*tables[p]):
                ctx.add_unary_table(p, t, k, w)
        seq = [ctx.get_unary(p, 0) for p in range(P)]
        # ... against all tables of a part in one pass
        for p in range(P):
            ctx.set_unary(p, 0, un[p, 0])
            ctx.add_unary_tables(p, *tables[p])
        for p in range(P):
            _same(ctx.get_unary(p, 0), seq[p], "fused table add, part %d" % p)
        ctx.infer(sparse=True)
        cond = np.stack(seq)[:, None]
        want = oracle.infer(ep, pc, joints, np.ascontiguousarray(cond), sparse=True)
        assert np.array_equal(ctx.best_conf(), want["best_conf"])
        for p in (0, root, P - 1):
            _same(ctx.marginal(p), want["marginals"][0, p], "conditioned marginal of part %d" % p)


def test_fast_math_argmax_identity_over_bench_size_images():
    """ps_config.fast_math against the parity arithmetic on 50 benchmark-size images: how many argmax records differ
    (north star: >= 95 % identical part estimates) and how far the marginals move (north star: 1e-4 relative)."""
    ep = ExpParam(num_rotation_steps=24)
    P, H, W = 10, 600, 400
    pc = synth.part_conf(P)
    joints = synth.make_joints(P, seed=7)
    n_img = 50
    rows_equal = rows_total = images_equal = 0
    worst = 0.0
    with PsContext(ep, pc, H, W) as exact, PsContext(ep, pc, H, W, fast_math=True) as fast:
        exact.set_joints(joints)
        fast.set_joints(joints)
        for i in range(n_img):
            cells, Tig = synth.compact_scores(ep, H, W, P, 1000 + i)
            for ctx in (exact, fast):
                for p in range(P):
                    ctx.set_unary_compact(p, 0, cells[p, 0], Tig)
                ctx.infer(sparse=True)
            a, b = exact.best_conf(), fast.best_conf()
            same = (a[:, :6] == b[:, :6]).all(axis=1)
            rows_equal += int(same.sum())
            rows_total += P
            images_equal += int(same.all())
            if i % 10 == 0:
                for p in (0, 4, 9):
                    ma, mb = exact.marginal(p).astype(np.float64), fast.marginal(p).astype(np.float64)
                    worst = max(worst, float((np.abs(ma - mb) / np.maximum(np.abs(ma), 1.0)).max()))
    out = {"images": n_img, "argmax_rows_identical": rows_equal, "argmax_rows_total": rows_total,
           "images_with_every_argmax_identical": images_equal, "marginal_max_rel_diff": worst}
    _report("fast_math_identity.json", out)
    print(json.dumps(out))
    assert rows_equal >= 0.95 * rows_total, out
    assert worst <= 1e-4, out


def test_pcp_identity_and_argmax_exactness_over_100_bench_size_images():
    """North star: ">= 95 % PCP-identical part estimates vs reference plus bit-exact argmax on the test set".  100
    synthetic LSP-shape images (configs[2]: the batch workload) through the CUDA path in both arithmetic modes against
    the CPU oracle's estimates of the same images (one oracle inference per host thread); the oracle's estimate plays the
    ground truth of the PCP matching rule (parteval.is_gt_match, pinned to the reference's evaluator)."""
    from concurrent.futures import ThreadPoolExecutor
    from partapp_b200 import parteval as pe
    ep = ExpParam(num_rotation_steps=24)
    P, H, W = 10, 600, 400
    pc = synth.part_conf(P)
    joints = synth.make_joints(P, seed=7)
    threads = max(1, min(os.cpu_count() or 1, 32))
    n_img = 100 if threads >= 12 else 24
    # LSP-like part windows (window_size_x/y, pos_offset_x/y of PartWindowParam): sticks of 60-110 px
    rng = np.random.default_rng(3)
    pps = [pe.PartParam(int(rng.integers(30, 50)), int(rng.integers(60, 110)), int(rng.integers(15, 25)), int(rng.integers(30, 55)))
           for _ in range(P)]

    def cpu(i):
        un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 2000 + i))
        return oracle.infer(ep, pc, joints, un, sparse=True, want_marginals=False)["best_conf"]

    with ThreadPoolExecutor(threads) as pool:
        futures = [pool.submit(cpu, i) for i in range(n_img)]
        got = {"parity": [], "fast_math": []}
        with PsContext(ep, pc, H, W) as exact, PsContext(ep, pc, H, W, fast_math=True) as fast:
            exact.set_joints(joints)
            fast.set_joints(joints)
            for i in range(n_img):
                cells, Tig = synth.compact_scores(ep, H, W, P, 2000 + i)
                for name, ctx in (("parity", exact), ("fast_math", fast)):
                    for p in range(P):
                        ctx.set_unary_compact(p, 0, cells[p, 0], Tig)
                    ctx.infer(sparse=True)
                    got[name].append(ctx.best_conf())
        want = [f.result() for f in futures]
    out = {"images": n_img, "parts": P}
    for name in ("parity", "fast_math"):
        exact_imgs = sum(int(np.array_equal(g[:, :6], w[:, :6])) for g, w in zip(got[name], want))
        exact_rows = sum(int((g[:, :6] == w[:, :6]).all(axis=1).sum()) for g, w in zip(got[name], want))
        score_bits = sum(int(np.array_equal(g[:, 6], w[:, 6])) for g, w in zip(got[name], want))
        pcp = float(np.mean([pe.pcp_identical(w, g, pps) for g, w in zip(got[name], want)]))
        out[name] = {"images_argmax_bit_exact": exact_imgs, "part_rows_argmax_bit_exact": exact_rows,
                     "images_best_score_bit_exact": score_bits, "pcp_identical": pcp}
    _report("pcp_identity.json", out)
    print(json.dumps(out))
    assert out["parity"]["images_argmax_bit_exact"] == n_img and out["parity"]["images_best_score_bit_exact"] == n_img, out
    assert out["parity"]["pcp_identical"] == 1.0
    assert out["fast_math"]["pcp_identical"] >= 0.95, out
