"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): argmax part positions and rotations bit-exact; marginals within 1e-4 relative.
The kernels are written to be bit-identical to the oracle (same fp32 summation order, fp64 exp/log whose narrowed
results equal glibc's for every fp32 input, profiles/r01_libm_exhaustive.txt), so the checks here demand EXACT
equality of every cell (_cmp with max_ulp_frac = 0); the 1e-4 tolerance of the north star is never used.
"""
import os

import numpy as np
import pytest

import oracle
from partapp_b200 import ExpParam, Joint, PartConf, PsContext, synth
from partapp_b200 import objectdetect as od

pytestmark = pytest.mark.gpu

LZ = np.float32(-1e6)


def _cmp(got, want, what, max_ulp_frac=0.0, rtol=1e-4):
    """Exact support of LOG_ZERO; the fraction of non-identical cells must not exceed max_ulp_frac (default 0: every
    cell identical); rtol only matters if a caller relaxes max_ulp_frac."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, what
    assert np.array_equal(got == LZ, want == LZ), what + ": LOG_ZERO support differs"
    neq = (got != want) & ~(np.isnan(got) & np.isnan(want))   # the same NaN in the same cell is agreement
    frac = neq.mean()
    if neq.any():
        rel = np.abs(got[neq].astype(np.float64) - want[neq]) / np.maximum(np.abs(want[neq].astype(np.float64)), 1e-30)
        assert rel.max() <= rtol, "%s: max rel err %g" % (what, rel.max())
    assert frac <= max_ulp_frac, "%s: %.3g of cells differ" % (what, frac)
    return frac


def _ctx(ep, P, H, W, **kw):
    return PsContext(ep, synth.part_conf(P, **kw), H, W)


def _child_grid(ep, H, W, seed, dense=False):
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, 1, seed)[0, 0])
    if dense:
        rng = np.random.default_rng(seed)
        un = (rng.standard_normal(un.shape) * 3 - 5).astype(np.float32)
    return un


MSG_CASES = [
    # name, R, H, W, off_in, off_out, C, rot_mean, rot_sigma, scale, sparse, dense
    ("diag_small", 8, 40, 48, (3.2, -5.7), (-4.4, 2.1), [[9.0, 0], [0, 4.0]], 0.3, 0.5, 1.0, True, False),
    ("diag_dense", 8, 40, 48, (3.2, -5.7), (-4.4, 2.1), [[9.0, 0], [0, 4.0]], -0.4, 0.9, 1.0, False, True),
    ("general_sparse", 8, 40, 48, (6.0, 1.5), (-2.5, -7.0), [[12.0, 5.0], [5.0, 7.0]], 0.2, 0.6, 1.0, True, False),
    ("general_dense", 8, 40, 48, (6.0, 1.5), (-2.5, -7.0), [[12.0, 5.0], [5.0, 7.0]], 0.2, 0.6, 1.0, False, True),
    ("general_negcov", 12, 37, 53, (-3.0, 4.5), (5.5, 1.0), [[6.0, -4.0], [-4.0, 10.0]], -0.7, 0.3, 1.0, False, False),
    ("rot_sigma_zero", 8, 32, 32, (2.0, 2.0), (-2.0, 1.0), [[4.0, 0], [0, 4.0]], 0.0, 0.0, 1.0, True, False),
    ("rot_kernel_clipped", 6, 32, 36, (2.0, 2.0), (-2.0, 1.0), [[4.0, 1.0], [1.0, 4.0]], 0.1, 2.5, 1.0, True, False),
    ("odd_R_clipped", 7, 30, 34, (1.0, -2.0), (3.0, 1.0), [[5.0, 0], [0, 3.0]], 0.5, 3.0, 1.0, True, False),
    ("big_shift_no_wrap", 8, 32, 32, (1.0, 1.0), (1.0, 1.0), [[4.0, 0], [0, 4.0]], 2.6, 0.4, 1.0, True, False),
    ("scale_1p2", 8, 40, 48, (3.2, -5.7), (-4.4, 2.1), [[9.0, 2.0], [2.0, 4.0]], 0.3, 0.5, 1.2, True, False),
    ("offsets_off_grid", 8, 24, 28, (60.0, -50.0), (4.0, 2.0), [[4.0, 0], [0, 4.0]], 0.0, 0.5, 1.0, True, False),
    ("wide_sigma", 8, 48, 56, (2.0, 1.0), (-1.0, 3.0), [[150.0, 30.0], [30.0, 90.0]], 0.0, 0.5, 1.0, False, True),
    ("ragged_W", 8, 33, 45, (2.5, 1.5), (-1.5, 3.5), [[8.0, 3.0], [3.0, 6.0]], 0.0, 0.5, 1.0, True, False),
]


@pytest.mark.parametrize("case", MSG_CASES, ids=[c[0] for c in MSG_CASES])
def test_message_matches_oracle(case):
    name, R, H, W, oi, oo, Cm, rm, rs, sc, sparse, dense = case
    ep = ExpParam(num_rotation_steps=R)
    child = _child_grid(ep, H, W, 11, dense)
    want = oracle.message(ep, child, oi, oo, Cm, rm, rs, sc, sparse)
    with _ctx(ep, 2, H, W) as ctx:
        got = od.computeRotJointMarginal(ctx, child, oi, oo, Cm, rm, rs, sc, sparse)
    _cmp(got, want, name)


def test_message_all_log_zero_input():
    ep = ExpParam(num_rotation_steps=8)
    child = np.full((8, 24, 24), LZ, np.float32)
    args = ((2.0, 1.0), (1.0, -2.0), [[4.0, 1.0], [1.0, 5.0]], 0.2, 0.5, 1.0, True)
    want = oracle.message(ep, child, *args)
    with _ctx(ep, 2, 24, 24) as ctx:
        got = ctx.message(child, *args)
    _cmp(got, want, "all LOG_ZERO")


@pytest.mark.parametrize("diagonal", [True, False], ids=["diagC", "fullC"])
@pytest.mark.parametrize("P", [4, 10])
def test_infer_matches_oracle(P, diagonal):
    ep = ExpParam(num_rotation_steps=8, roi_save_num_samples=50)
    H, W = 48, 40
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 3))
    joints = synth.make_joints(P, seed=5, diagonal=diagonal, max_offset=8, sigma_range=(1.5, 4))
    pc = synth.part_conf(P)
    ref_un = un.copy()
    want = oracle.infer(ep, pc, joints, ref_un, sparse=True)
    with PsContext(ep, pc, H, W) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            ctx.set_unary(p, 0, un[p, 0])
        ctx.infer(sparse=True)
        best = ctx.best_conf()
        # argmax: scaleidx, rotidx, x, y bit-exact
        assert np.array_equal(best[:, [0, 2, 4, 5]], want["best_conf"][:, [0, 2, 4, 5]])
        assert np.array_equal(best[:, [1, 3]], want["best_conf"][:, [1, 3]])
        np.testing.assert_allclose(best[:, 6], want["best_conf"][:, 6], rtol=1e-6)
        for p in range(P):
            _cmp(ctx.marginal(p), want["marginals"][0, p], "marginal part %d" % p)
        _cmp(ctx.root_posterior(), want["root_post"], "root posterior")


def test_infer_multiscale_upright_strip():
    ep = ExpParam(num_rotation_steps=8, num_scale_steps=3, min_object_scale=0.8, max_object_scale=1.2,
                  strip_border_detections=0.1, roi_save_num_samples=20)
    P, H, W = 4, 40, 44
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 9))
    joints = synth.make_joints(P, seed=2, max_offset=8, sigma_range=(1.5, 4))
    pc = synth.part_conf(P, upright_root=True)
    ref_un = un.copy()
    want = oracle.infer(ep, pc, joints, ref_un, sparse=True)
    with PsContext(ep, pc, H, W, keep_all_scales=True) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            for s in range(3):
                ctx.set_unary(p, s, un[p, s])
        ctx.infer(sparse=True)
        assert np.array_equal(ctx.best_conf()[:, :6], want["best_conf"][:, :6])
        for s in range(3):
            for p in range(P):
                _cmp(ctx.marginal(p, s), want["marginals"][s, p], "marginal s%d p%d" % (s, p))
        _cmp(ctx.root_posterior(), want["root_post"], "root posterior")
        # the unaries are masked in place exactly like the reference mutates its argument
        for p in range(P):
            for s in range(3):
                assert np.array_equal(ctx.get_unary(p, s), ref_un[p, s])


def test_prepare_unary_on_device_matches_oracle():
    ep = ExpParam(num_rotation_steps=8)
    raw = synth.raw_scores(ep, 40, 36, 2, 1)
    want = oracle.prepare_unary(raw)
    with _ctx(ep, 2, 40, 36) as ctx:
        for p in range(2):
            ctx.set_unary(p, 0, raw[p, 0], raw_scores=True)
            _cmp(ctx.get_unary(p, 0), want[p, 0], "unary prep")


def test_find_local_max_matches_oracle():
    rng = np.random.default_rng(0)
    g = rng.standard_normal((5, 30, 34)).astype(np.float32)
    g[1, 5:9, 5:9] = 7.0  # plateau: ties are allowed in the 8-neighbourhood
    ep = ExpParam(num_rotation_steps=8)
    with _ctx(ep, 2, 8, 8) as ctx:
        for K in (1000, 20):
            got = ctx.find_local_max(g, K)
            want = oracle.find_local_max(g, K)
            assert len(got) == len(want)
            if K == 1000:
                assert np.array_equal(got, want)  # scan order
            else:
                assert np.array_equal(np.sort(got[:, 3])[::-1], np.sort(want[:, 3])[::-1])


# ---- golden fixtures (oracle outputs committed under tests/golden) -------------------------------------------------

def test_messages_match_golden_fixtures():
    import os
    from tests.golden.make_golden import CASES, case_input
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "messages.npz"))
    for name, spec in CASES.items():
        ep, g = case_input(name)
        R, H, W = g.shape
        oi, oo, Cm, rm, rs, sc, sparse = spec[5:]
        with _ctx(ep, 2, H, W) as ctx:
            got = ctx.message(g, oi, oo, Cm, rm, rs, sc, sparse)
        _cmp(got, z[name], "golden " + name)


def test_infer_matches_golden_fixture():
    import os
    from tests.golden.make_golden import infer_case
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "infer.npz"))
    ep, pc, joints, un = infer_case()
    P, S, R, H, W = un.shape
    with PsContext(ep, pc, H, W) as ctx:
        res = od.computeRootPosteriorRot(ctx, [[un[p, s] for s in range(S)] for p in range(P)], joints, True)
        assert np.array_equal(res.best_conf[:, :6], z["best_conf"][:, :6])
        _cmp(res.root_part_posterior, z["root_post"], "golden root posterior")
        for p in range(P):
            _cmp(ctx.marginal(p), z["marginals"][0, p], "golden marginal %d" % p)


# ---- the rest of the seam -------------------------------------------------------------------------------------------

def test_conditioning_adds_match_oracle():
    """a5: addExtraUnary with rotation / position tables and the torso prior (icps.cpp:137-191,228-281,366-423,526-548)."""
    import ctypes
    ep = ExpParam(num_rotation_steps=8)
    P, H, W = 3, 28, 24
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 2))
    L = oracle.lib()
    fp = ctypes.POINTER(ctypes.c_float)
    rot_t = np.zeros(8, np.float32)
    L.orc_rot_score_table(ctypes.byref(oracle.exp_param(ep)), 0.4, 0.3, rot_t.ctypes.data_as(fp))
    pos_t = np.zeros((H, W), np.float32)
    L.orc_pos_score_table(H, W, 3.0, -2.0, 30.0, 45.0, 10.0, 12.0, pos_t.ctypes.data_as(fp))
    tor_t = np.zeros((H, W), np.float32)
    L.orc_torso_prior_table(H, W, 1.0, 2.0, 60.0, 90.0, 0.8, tor_t.ctypes.data_as(fp))
    want = un.copy()
    L.orc_add_rot_table(want[1, 0].ctypes.data_as(fp), 8, H, W, rot_t.ctypes.data_as(fp), 0.35)
    L.orc_add_pos_table(want[1, 0].ctypes.data_as(fp), 8, H, W, pos_t.ctypes.data_as(fp), 0.6)
    L.orc_add_pos_table_unweighted(want[0, 0].ctypes.data_as(fp), 8, H, W, tor_t.ctypes.data_as(fp))
    with _ctx(ep, P, H, W) as ctx:
        for p in range(P):
            ctx.set_unary(p, 0, un[p, 0])
        ctx.add_unary_table(1, rot_t, 0, 0.35)
        ctx.add_unary_table(1, pos_t, 1, 0.6)
        ctx.add_unary_table(0, tor_t, 2)
        for p in range(P):
            assert np.array_equal(ctx.get_unary(p, 0), want[p, 0]), "part %d" % p


def test_flipped_joints_match_oracle():
    ep = ExpParam(num_rotation_steps=8, roi_save_num_samples=10)
    P, H, W = 4, 36, 32
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 4))
    joints = [j.flipped() for j in synth.make_joints(P, seed=9, max_offset=6, sigma_range=(1.5, 3))]
    ref = synth.make_joints(P, seed=9, max_offset=6, sigma_range=(1.5, 3))
    for a, b in zip(joints, ref):   # ps_flip_joint == orc_flip_joint
        ob = oracle.joint(b)
        oracle.lib().orc_flip_joint(ob)
        assert list(a.offset_c) == [ob.offset_c[0], ob.offset_c[1]] and a.rot_mean == ob.rot_mean
        assert np.array_equal(np.asarray(a.C).reshape(4), [ob.C[i] for i in range(4)])
    pc = synth.part_conf(P)
    want = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
    with PsContext(ep, pc, H, W) as ctx:
        res = od.computeRootPosteriorRot(ctx, [[un[p, 0]] for p in range(P)], joints, True)
        assert np.array_equal(res.best_conf[:, :6], want["best_conf"][:, :6])


def test_get_max_states_matches_oracle():
    import ctypes
    ep = ExpParam(num_rotation_steps=8, roi_save_num_samples=25)
    P, H, W = 3, 30, 26
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 6))
    want = np.zeros((P, 7), np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    oracle.lib().orc_get_max_states(ctypes.byref(oracle.exp_param(ep)), P, H, W, un.ctypes.data_as(fp),
                                    want.ctypes.data_as(fp))
    with _ctx(ep, P, H, W) as ctx:
        best, hyps = od.getMaxStates(ctx, [[un[p, 0]] for p in range(P)], local_max=True)
        assert np.array_equal(best, want)
        for p in range(P):
            lm = oracle.find_local_max(un[p, 0], 25)
            assert len(hyps[p]) == 1 + len(lm)
            assert np.array_equal(np.sort(hyps[p][1:, 6])[::-1], np.sort(lm[:, 3])[::-1])


def test_infer_local_maxima_and_root_hypotheses():
    ep = ExpParam(num_rotation_steps=8, roi_save_num_samples=30)
    P, H, W = 4, 40, 36
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 8))
    joints = synth.make_joints(P, seed=1, max_offset=6, sigma_range=(1.5, 3))
    pc = synth.part_conf(P)
    want = oracle.infer(ep, pc, joints, un.copy(), sparse=True, want_hyps=True)
    with PsContext(ep, pc, H, W) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            ctx.set_unary(p, 0, un[p, 0])
        ctx.infer(sparse=True, local_max=True, root_hyps=True)
        for p in range(P):
            got, ref = ctx.part_hyps(p), want["part_hyps"][p]
            assert len(got) == len(ref)
            assert np.array_equal(got[0, :6], ref[0, :6])                       # entry 0 = the argmax record
            np.testing.assert_allclose(np.sort(got[1:, 6])[::-1], np.sort(ref[1:, 6])[::-1], rtol=1e-6)
        rh = ctx.root_hyps()
        ref = oracle.find_local_max(want["root_post"], 1000)
        sel = ref[:, 3] > -5e5
        got_sel = rh[rh[:, 3] > -5e5]
        assert len(got_sel) == sel.sum()
        np.testing.assert_allclose(np.sort(got_sel[:, 3]), np.sort(ref[sel, 3]), rtol=1e-6)


def test_keep_unaries_restores_inputs():
    ep = ExpParam(num_rotation_steps=8, strip_border_detections=0.2)
    P, H, W = 3, 24, 30
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 12))
    joints = synth.make_joints(P, seed=1, max_offset=4, sigma_range=(1.5, 2.5))
    with PsContext(ep, synth.part_conf(P, upright_root=True), H, W) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            ctx.set_unary(p, 0, un[p, 0])
        ctx.infer(sparse=True, keep_unaries=True)
        ctx.best_conf()
        for p in range(P):
            assert np.array_equal(ctx.get_unary(p, 0), un[p, 0])


def test_22_part_tree_matches_oracle():
    ep = ExpParam(num_rotation_steps=8, roi_save_num_samples=5)
    P, H, W = 22, 32, 28
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 21))
    joints = synth.make_joints(P, seed=11, max_offset=5, sigma_range=(1.2, 2.5))
    pc = synth.part_conf(P)
    want = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
    with PsContext(ep, pc, H, W) as ctx:
        res = od.computeRootPosteriorRot(ctx, [[un[p, 0]] for p in range(P)], joints, True, write_back_masked=False)
        assert np.array_equal(res.best_conf[:, :6], want["best_conf"][:, :6])
        for p in (0, 4, 10, 21):
            _cmp(ctx.marginal(p), want["marginals"][0, p], "22-part marginal %d" % p)


def test_error_statuses_replace_asserts():
    from partapp_b200 import PsInferError, capi
    ep = ExpParam(num_rotation_steps=8)
    with _ctx(ep, 3, 16, 16) as ctx:
        good = synth.make_joints(3, max_offset=3, sigma_range=(1.5, 2))
        with pytest.raises(PsInferError) as e:
            ctx.infer()
        assert e.value.status == capi.PS_ERR_STATE
        with pytest.raises(PsInferError) as e:
            ctx.set_joints(good[:1])                       # aux.cpp:129: need P-1 joints
        assert e.value.status == capi.PS_ERR_INVALID
        bad = [Joint(1, 0, [1, 1], [1, 1], [[2, 0], [0, 2]], 0, 0.5, type=capi.PS_JOINT_POS_GAUSSIAN), good[1]]
        with pytest.raises(PsInferError) as e:
            ctx.set_joints(bad)                            # findrot.cpp:766
        assert e.value.status == capi.PS_ERR_UNSUPPORTED
        notpd = [Joint(1, 0, [1, 1], [1, 1], [[2, 0], [0, -1]], 0, 0.5), good[1]]
        with pytest.raises(PsInferError):
            ctx.set_joints(notpd)                          # filter.hpp:219
    # a non-root part with two children (findrot.cpp:210)
    with _ctx(ep, 4, 16, 16) as ctx:
        js = [Joint(1, 0, [1, 1], [1, 1], [[2, 0], [0, 2]], 0, 0.5), Joint(2, 1, [1, 1], [1, 1], [[2, 0], [0, 2]], 0, 0.5),
              Joint(3, 1, [1, 1], [1, 1], [[2, 0], [0, 2]], 0, 0.5)]
        with pytest.raises(PsInferError) as e:
            ctx.set_joints(js)
        assert e.value.status == capi.PS_ERR_INVALID


# ---- full size (BASELINE.json configs[1]: R=24, 600x400) -----------------------------------------------------------

def test_full_size_message_matches_oracle():
    ep = ExpParam(num_rotation_steps=24)
    H, W = 600, 400
    child = oracle.prepare_unary(synth.raw_scores(ep, H, W, 1, 5)[0, 0])
    j = synth.make_joints(10, seed=7)[1]
    with _ctx(ep, 2, H, W) as ctx:
        up = ctx.message(child, j.offset_c, j.offset_p, j.C, j.rot_mean, j.rot_sigma, 1.0, True)
        want_up = oracle.message(ep, child, j.offset_c, j.offset_p, j.C, j.rot_mean, j.rot_sigma, 1.0, True)
        _cmp(up, want_up, "full-size upward message")
        down = ctx.message(want_up, j.offset_p, j.offset_c, j.C, -j.rot_mean, j.rot_sigma, 1.0, False)
        want_down = oracle.message(ep, want_up, j.offset_p, j.offset_c, j.C, -j.rot_mean, j.rot_sigma, 1.0, False)
        _cmp(down, want_down, "full-size downward message")


def test_full_size_properties():
    """Size-independent properties at the benchmark size: repeatability (bitwise), shift of the unaries by a constant
    shifts every marginal by P times that constant up to rounding and leaves the argmax in place, and the root
    posterior is the log-sum-exp of the root marginal."""
    ep = ExpParam(num_rotation_steps=24)
    P, H, W = 10, 600, 400
    raw = synth.raw_scores(ep, H, W, P, 0)
    joints = synth.make_joints(P, seed=7)
    with PsContext(ep, synth.part_conf(P), H, W) as ctx:
        ctx.set_joints(joints)

        def run():
            for p in range(P):
                ctx.set_unary(p, 0, raw[p, 0], raw_scores=True)
            ctx.infer(sparse=True)
            return ctx.best_conf(), ctx.marginal(4), ctx.root_posterior()

        b1, m1, r1 = run()
        b2, m2, r2 = run()
        assert np.array_equal(b1, b2) and np.array_equal(m1, m2) and np.array_equal(r1, r2)
        # planted bumps make the argmax meaningful: every part's best score is finite and well above LOG_ZERO
        assert (b1[:, 6] > -1e5).all()
        # root posterior = log sum_r exp(root marginal) (findrot.cpp:714-726), checked in float64
        with np.errstate(under="ignore"):
            lse = np.log(np.exp(m1.astype(np.float64)).sum(0))
        ok = np.isfinite(lse) & (r1[0] > -1e5)
        assert ok.sum() > 1000
        np.testing.assert_allclose(r1[0][ok], lse[ok], rtol=1e-4, atol=1e-3)


# ---- unary ingest from compact detector grids (PartApp::loadScoreGrid, partapp.cpp:830-903) -----------

@pytest.mark.parametrize("rotated", [False, True], ids=["lattice", "rotated_lattice"])
def test_compact_ingest_matches_load_score_grid(rotated):
    ep = ExpParam(num_rotation_steps=8)
    P, H, W = 2, 44, 52
    cells, Tig = synth.compact_scores(ep, H, W, P, 3, rotated=rotated)
    with _ctx(ep, P, H, W) as ctx:
        for p in range(P):
            want = oracle.prepare_unary(oracle.load_score_grid(cells[p, 0], Tig, H, W))
            ctx.set_unary_compact(p, 0, cells[p, 0], Tig)
            got = ctx.get_unary(p, 0)
            _cmp(got, want, "compact ingest part %d" % p)


def test_compact_ingest_collisions_last_writer_wins():
    """A down-scaling transform maps several grid cells onto one image cell: the reference's x-outer / y-inner
    scatter order decides which one survives (transform.hpp:176-190)."""
    ep = ExpParam(num_rotation_steps=4)
    H, W = 20, 24
    rng = np.random.default_rng(1)
    cells = rng.uniform(0.05, 1.0, (4, 30, 36)).astype(np.float32)
    cells[rng.random(cells.shape) < 0.3] = 0.0           # unevaluated cells are skipped, not written
    Tig = np.zeros((4, 3, 3))
    for r in range(4):
        th = 0.3 * r
        Tig[r] = [[0.6 * np.cos(th), -0.6 * np.sin(th), 4.0 + r], [0.6 * np.sin(th), 0.6 * np.cos(th), 1.5], [0, 0, 1]]
    want = oracle.prepare_unary(oracle.load_score_grid(cells, Tig, H, W))
    with _ctx(ep, 2, H, W) as ctx:
        ctx.set_unary_compact(0, 0, cells, Tig)
        _cmp(ctx.get_unary(0, 0), want, "colliding scatter")


@pytest.mark.parametrize("rotated", [False, True], ids=["lattice", "rotated_lattice"])
def test_compact_ingest_interpolate_matches_load_score_grid(rotated):
    """ExpParam.interpolate: TM_BILINEAR gather through inverse(Tig) (partapp.cpp:889-891, transform.hpp:196-238)."""
    ep = ExpParam(num_rotation_steps=8, interpolate=True)
    P, H, W = 2, 44, 52
    cells, Tig = synth.compact_scores(ep, H, W, P, 3, rotated=rotated)
    dense = np.random.default_rng(3).uniform(-0.2, 1.0, cells[0, 0].shape).astype(np.float32)  # fully evaluated grid
    with _ctx(ep, P, H, W) as ctx:
        for p, c in ((0, cells[0, 0]), (1, dense)):
            want = oracle.prepare_unary(oracle.load_score_grid(c, Tig, H, W, interpolate=True))
            ctx.set_unary_compact(p, 0, c, Tig)
            _cmp(ctx.get_unary(p, 0), want, "interpolated ingest part %d" % p)
        assert (ctx.get_unary(1, 0) > -1e5).sum() > 100


@pytest.mark.parametrize("rotated", [False, True], ids=["lattice", "rotated_lattice"])
def test_compact_ingest_all_parts_in_one_call(rotated):
    """ps_set_unaries_compact: every (part, scale) grid of an image on one lattice, one fill + one scatter launch --
    same unaries and same folded maxima (the inference that follows must not change) as one call per grid."""
    ep = ExpParam(num_rotation_steps=8, num_scale_steps=2, min_object_scale=0.9, max_object_scale=1.1)
    P, H, W = 4, 44, 52
    cells, Tig = synth.compact_scores(ep, H, W, P, 7, rotated=rotated)
    joints = synth.make_joints(P, seed=5, max_offset=6, sigma_range=(1.5, 3))
    pc = synth.part_conf(P)
    with PsContext(ep, pc, H, W) as one, PsContext(ep, pc, H, W) as many:
        many.set_joints(joints)
        one.set_joints(joints)
        for p in range(P):
            for s in range(2):
                many.set_unary_compact(p, s, cells[p, s], Tig)
        ps_ = [p for p in range(P) for s in range(2)]
        ss = [s for p in range(P) for s in range(2)]
        n0 = one.launch_count()
        one.set_unaries_compact(ps_, ss, [cells[p, s] for p, s in zip(ps_, ss)], Tig)
        assert one.launch_count() - n0 <= 3
        for p in range(P):
            for s in range(2):
                assert np.array_equal(one.get_unary(p, s), many.get_unary(p, s)), (p, s)
        one.infer(sparse=True)
        many.infer(sparse=True)
        assert np.array_equal(one.best_conf(), many.best_conf())
        assert np.array_equal(one.marginal(1), many.marginal(1))
        # a subset of the grids (not the whole buffer) takes the per-grid fills
        one.set_unaries_compact([2], [1], [cells[0, 0]], Tig)
        many.set_unary_compact(2, 1, cells[0, 0], Tig)
        assert np.array_equal(one.get_unary(2, 1), many.get_unary(2, 1))
        assert np.array_equal(one.get_unary(2, 0), many.get_unary(2, 0))


@pytest.mark.parametrize("rotated", [False, True], ids=["lattice", "rotated_lattice"])
def test_compact_ingest_image_after_image_without_the_fill(rotated):
    """A second image on the same lattice overwrites every lattice cell instead of refilling the buffer (LOG_ZERO where
    its own score is 0 and the previous image's was not); any other writer of the unaries, a different lattice, or an
    inference that masks them in place brings the fill back.  Every step is compared with loadScoreGrid + the unary
    prep of the oracle for exactly that image."""
    ep = ExpParam(num_rotation_steps=8, num_scale_steps=2, min_object_scale=0.9, max_object_scale=1.1)
    P, H, W = 3, 44, 52
    ps_ = [p for p in range(P) for s in range(2)]
    ss = [s for p in range(P) for s in range(2)]

    def image(k, rot=rotated, stride=4):
        cells, Tig = synth.compact_scores(ep, H, W, P, 20 + k, rotated=rot, stride=stride)
        rng = np.random.default_rng(50 + k)
        cells = cells.copy()
        cells[rng.random(cells.shape) < 0.4] = 0.0       # a different set of unevaluated cells in every image
        return cells, Tig

    def check(ctx, cells, Tig, what):
        for p in range(P):
            for s in range(2):
                want = oracle.prepare_unary(oracle.load_score_grid(cells[p, s], Tig, H, W))
                _cmp(ctx.get_unary(p, s), want, "%s part %d scale %d" % (what, p, s))

    def ingest(ctx, cells, Tig):
        n0 = ctx.launch_count()
        ctx.set_unaries_compact(ps_, ss, [cells[p, s] for p, s in zip(ps_, ss)], Tig)
        return ctx.launch_count() - n0

    upright = synth.part_conf(P)
    upright.is_upright[1] = True
    with PsContext(ep, synth.part_conf(P), H, W) as ctx, PsContext(ep, upright, H, W) as up:
        ctx.set_joints(synth.make_joints(P, seed=5, max_offset=6, sigma_range=(1.5, 3)))
        up.set_joints(synth.make_joints(P, seed=5, max_offset=6, sigma_range=(1.5, 3)))
        c0, T0 = image(0)
        assert ingest(ctx, c0, T0) == 3                   # fill, max reset, scatter
        check(ctx, c0, T0, "first image")
        c1, T1 = image(1)
        assert ingest(ctx, c1, T1) == 2                   # same lattice: no fill
        check(ctx, c1, T1, "second image, fill skipped")
        ctx.infer(sparse=True)                            # reads the unaries, writes none of them
        c2, T2 = image(2)
        assert ingest(ctx, c2, T2) == 2
        check(ctx, c2, T2, "third image after an inference")
        ctx.add_unary_table(1, np.linspace(-1, 0, 8).astype(np.float32), kind=0, weight=0.5)   # another writer
        c3, T3 = image(3)
        assert ingest(ctx, c3, T3) == 3
        check(ctx, c3, T3, "image after a conditioning add")
        ctx.set_unary_compact(0, 1, c2[0, 1], T2)         # single-grid call: also another writer
        assert ingest(ctx, c3, T3) == 3
        c4, T4 = image(4, rot=not rotated)                # a different lattice
        assert ingest(ctx, c4, T4) == 3
        check(ctx, c4, T4, "image on another lattice")
        c5, T5 = image(5, rot=not rotated)
        assert ingest(ctx, c5, T5) == 2
        check(ctx, c5, T5, "second image on the other lattice")
        # upright masking rewrites unary slices in place (findrot.cpp:509-523): without keep_unaries the fill returns
        assert ingest(up, c0, T0) == 3 and ingest(up, c1, T1) == 2
        up.infer(sparse=True, keep_unaries=True)
        assert ingest(up, c2, T2) == 2
        check(up, c2, T2, "after an inference that restored the unaries")
        up.infer(sparse=True)
        assert ingest(up, c3, T3) == 3
        check(up, c3, T3, "after an inference that masked the unaries in place")


def test_compact_ingest_many_rotations_back_to_back():
    """R > 64: the per-rotation transforms go through a device staging buffer shared by every (part, scale) call.
    Calls issued back to back, each with its OWN transforms, must not overwrite the rows a previous call's scatter
    kernel is still reading."""
    ep = ExpParam(num_rotation_steps=72)
    P, H, W = 4, 36, 44
    cells, Tig0 = synth.compact_scores(ep, H, W, P, 5, rotated=True)
    tigs = []
    for p in range(P):
        T = Tig0.copy()
        T[:, 0, 2] += 1.5 * p          # a different lattice per part
        T[:, 1, 2] -= 0.75 * p
        tigs.append(T)
    with _ctx(ep, P, H, W) as ctx:
        for p in range(P):             # no synchronisation between the calls
            ctx.set_unary_compact(p, 0, cells[p, 0], tigs[p])
        for p in range(P):
            want = oracle.prepare_unary(oracle.load_score_grid(cells[p, 0], tigs[p], H, W))
            _cmp(ctx.get_unary(p, 0), want, "R=72 ingest part %d" % p)


def test_find_local_max_between_infer_and_getters():
    """ps_infer only enqueues; a synchronous ps_find_local_max issued before the getters shares the top-K slot and
    the pinned winner buffers with the pending readout and must not clobber it (also when it needs a larger top-K)."""
    ep = ExpParam(num_rotation_steps=8, roi_save_num_samples=30)
    P, H, W = 4, 40, 36
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 2))
    joints = synth.make_joints(P, seed=4, max_offset=6, sigma_range=(1.5, 3))
    pc = synth.part_conf(P)
    with PsContext(ep, pc, H, W) as ctx:
        ctx.set_joints(joints)

        def run(disturb):
            for p in range(P):
                ctx.set_unary(p, 0, un[p, 0])
            ctx.infer(sparse=True, local_max=True, root_hyps=True)
            if disturb:
                ctx.find_local_max(un[1, 0], 5000)   # > the readout's top-K capacity: buffers are reallocated
            return [ctx.part_hyps(p) for p in range(P)], ctx.root_hyps(), ctx.best_conf()

        h0, r0, b0 = run(False)
        h1, r1, b1 = run(True)
        assert np.array_equal(b0, b1)
        assert np.array_equal(r0, r1)
        for p in range(P):
            assert np.array_equal(h0[p], h1[p]), "part %d hypotheses clobbered" % p


# ---- BASELINE.json configs[3] / configs[4] shapes at test size ------------------------------------------------------

def test_conditioned_model_swaps_joints_per_image():
    """configs[3]: per image, every joint is one entry of a finite per-joint type table (aux.cpp:76-99)."""
    ep = ExpParam(num_rotation_steps=8, roi_save_num_samples=5)
    P, H, W = 6, 36, 32
    pc = synth.part_conf(P)
    table = [synth.make_joints(P, seed=13, max_offset=6, sigma_range=(1.5, 3), type_id=t) for t in range(3)]
    with PsContext(ep, pc, H, W) as ctx:
        for img in range(4):
            rng = np.random.default_rng(img)
            joints = [table[int(rng.integers(0, 3))][j] for j in range(P - 1)]
            un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, img))
            want = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
            res = od.computeRootPosteriorRot(ctx, [[un[p, 0]] for p in range(P)], joints, True, write_back_masked=False)
            assert np.array_equal(res.best_conf[:, :6], want["best_conf"][:, :6]), "image %d" % img
            _cmp(ctx.marginal(2), want["marginals"][0, 2], "conditioned image %d" % img)


def test_stress_shape_48_rotations_multiscale():
    """configs[4] shape (R = 48, several scales) at a grid the oracle finishes in seconds."""
    ep = ExpParam(num_rotation_steps=48, num_scale_steps=2, min_object_scale=0.8, max_object_scale=1.2,
                  roi_save_num_samples=5)
    P, H, W = 4, 56, 48
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 17))
    joints = synth.make_joints(P, seed=3, max_offset=8, sigma_range=(1.5, 4))
    pc = synth.part_conf(P)
    want = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
    with PsContext(ep, pc, H, W, keep_all_scales=True) as ctx:
        res = od.computeRootPosteriorRot(ctx, [[un[p, s] for s in range(2)] for p in range(P)], joints, True,
                                         write_back_masked=False)
        assert np.array_equal(res.best_conf[:, :6], want["best_conf"][:, :6])
        for s in range(2):
            _cmp(ctx.marginal(1, s), want["marginals"][s, 1], "R48 marginal scale %d" % s)
        _cmp(res.root_part_posterior, want["root_post"], "R48 root posterior")


# ---- randomized sweep ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("seed", range(12))
def test_random_messages_match_oracle(seed):
    """Random shapes (odd and even widths, 4..25 rotations), offsets, covariances (diagonal with probability 1/4),
    rotation Gaussians (including sigma = 0 and kernels that wrap the whole circle), scales and sparsity flags."""
    rng = np.random.default_rng(1000 + seed)
    R = int(rng.integers(4, 26))
    H, W = int(rng.integers(17, 70)), int(rng.integers(17, 70))
    ep = ExpParam(num_rotation_steps=R, min_part_rotation=float(rng.choice([-180, -90, -120])),
                  max_part_rotation=float(rng.choice([180, 90, 150])))
    dense = bool(rng.integers(0, 2))
    child = _child_grid(ep, H, W, seed, dense)
    s1, s2 = rng.uniform(0.6, 6.0, 2)
    if rng.random() < 0.25:
        Cm = [[s1 * s1, 0.0], [0.0, s2 * s2]]
    else:
        th = rng.uniform(0, np.pi)
        Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        Cm = Rm @ np.diag([s1 * s1, s2 * s2]) @ Rm.T
        Cm[1, 0] = Cm[0, 1]
        Cm = Cm.tolist()
    oi, oo = rng.uniform(-12, 12, 2), rng.uniform(-12, 12, 2)
    rm = float(rng.uniform(-1.5, 1.5))
    rs = float(rng.choice([0.0, rng.uniform(0.05, 3.0)]))
    sc = float(rng.choice([1.0, rng.uniform(0.7, 1.4)]))
    sparse = bool(rng.integers(0, 2))
    want = oracle.message(ep, child, oi, oo, Cm, rm, rs, sc, sparse)
    with _ctx(ep, 2, H, W) as ctx:
        got = ctx.message(child, oi, oo, Cm, rm, rs, sc, sparse)
    _cmp(got, want, "random message %d (R=%d %dx%d dense=%s sparse=%s)" % (seed, R, H, W, dense, sparse))


@pytest.mark.parametrize("case", [c for c in MSG_CASES if c[0] in ("general_sparse", "general_dense", "wide_sigma", "diag_dense")],
                         ids=lambda c: c[0])
def test_fallback_kernels_without_tma(case, monkeypatch):
    """PSINFER_NO_TMA=1 routes the Gaussian passes through the shared-memory staged kernels (the path taken when a
    filter is too long for a TMA box or a pitch is not a multiple of 4); results must not change."""
    monkeypatch.setenv("PSINFER_NO_TMA", "1")
    name, R, H, W, oi, oo, Cm, rm, rs, sc, sparse, dense = case
    ep = ExpParam(num_rotation_steps=R)
    child = _child_grid(ep, H, W, 11, dense)
    want = oracle.message(ep, child, oi, oo, Cm, rm, rs, sc, sparse)
    with _ctx(ep, 2, H, W) as ctx:
        got = ctx.message(child, oi, oo, Cm, rm, rs, sc, sparse)
    _cmp(got, want, name + " (no TMA)")


@pytest.mark.parametrize("k", range(8))
def test_multi_tile_work_lists_match_oracle(k):
    """Grids of several 64-cell strips and covariance axes at eight angles: the Gaussian work lists (one interval of
    8-row groups per strip, partial last tile) leave out most of the eigen-frame bounding box; every cell the
    read-back touches must still carry the reference's bits."""
    rng = np.random.default_rng(4200 + k)
    R = 4
    H, W = [(150, 230), (260, 140), (129, 191), (200, 200)][k % 4]
    ep = ExpParam(num_rotation_steps=R)
    th = (k + 0.37) * np.pi / 8
    s1, s2 = rng.uniform(2.0, 9.0, 2)
    Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    Cm = Rm @ np.diag([s1 * s1, s2 * s2]) @ Rm.T
    Cm[1, 0] = Cm[0, 1]
    Cm = Cm.tolist()
    dense = bool(k & 1)
    child = _child_grid(ep, H, W, 70 + k, dense)
    oi, oo = rng.uniform(-20, 20, 2), rng.uniform(-20, 20, 2)
    want = oracle.message(ep, child, oi, oo, Cm, 0.3, 0.6, 1.0, not dense)
    with _ctx(ep, 2, H, W) as ctx:
        got = ctx.message(child, oi, oo, Cm, 0.3, 0.6, 1.0, not dense)
    _cmp(got, want, "multi-tile message %d (%dx%d, angle %.2f)" % (k, H, W, th))


def test_work_list_longer_than_shared_memory_copy():
    """A 2400 x 2400 grid with an oblique covariance needs more than 1024 tiles per slice: the Gaussian kernel then
    reads its work list from global memory instead of its shared-memory copy."""
    ep = ExpParam(num_rotation_steps=2)
    H = W = 2400
    rng = np.random.default_rng(77)
    child = np.full((2, H, W), LZ, np.float32)
    ys, xs = rng.integers(0, H, 4000), rng.integers(0, W, 4000)
    child[rng.integers(0, 2, 4000), ys, xs] = rng.uniform(-6, 0, 4000).astype(np.float32)
    th = 0.7
    Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    Cm = Rm @ np.diag([9.0, 4.0]) @ Rm.T
    Cm[1, 0] = Cm[0, 1]
    want = oracle.message(ep, child, (3.0, -2.0), (-1.0, 4.0), Cm.tolist(), 0.0, 0.0, 1.0, True)
    with _ctx(ep, 2, H, W) as ctx:
        got = ctx.message(child, (3.0, -2.0), (-1.0, 4.0), Cm.tolist(), 0.0, 0.0, 1.0, True)
    _cmp(got, want, "message on a 2400 x 2400 grid")   # 5.76 M cells of footprint > 1024 tiles of 64 x 64


def test_work_lists_change_no_bit_and_skip_cells(monkeypatch):
    """Same inference with and without the work lists (PSINFER_ALL_TILES=1 filters every eigen-frame tile): identical
    marginals; and the lists really leave cells out (ps_get_plan_info reports fewer cells than rows x cols)."""
    ep = ExpParam(num_rotation_steps=6, roi_save_num_samples=5)
    P, H, W = 4, 170, 140
    raw = synth.raw_scores(ep, H, W, P, 3)
    joints = synth.make_joints(P, seed=9, max_offset=15, sigma_range=(2, 7))

    def run():
        with _ctx(ep, P, H, W) as ctx:
            ctx.set_joints(joints)
            for p in range(P):
                ctx.set_unary(p, 0, raw[p, 0], raw_scores=True)
            ctx.infer(sparse=True)
            infos = [ctx.plan_info(j, d) for j in range(P - 1) for d in (0, 1)]
            return ctx.best_conf(), [ctx.marginal(p) for p in range(P)], infos

    best_a, marg_a, infos_a = run()
    monkeypatch.setenv("PSINFER_ALL_TILES", "1")
    best_b, marg_b, infos_b = run()
    assert np.array_equal(best_a, best_b)
    for p in range(P):
        assert np.array_equal(marg_a[p], marg_b[p]), "marginal %d changes with the work lists" % p
    full = [i for i in infos_a if not i["diag"]]
    assert full, "the synthetic joints should have full covariances"
    for ia, ib in zip(infos_a, infos_b):
        assert ib["x_cells"] == ib["rows"] * ib["cols"] and ib["y_cells"] == ib["rows"] * ib["cols"]
        if not ia["diag"]:
            assert ia["y_cells"] < ib["y_cells"] and ia["x_cells"] <= ib["x_cells"]


# ---- edge cases -------------------------------------------------------------------------------------------------------

def test_tiny_grids_and_single_rotation_bins():
    """Smallest shapes: a 3x3 grid, a filter far wider than the grid, 2 rotation bins (kernel clipped to 1 tap)."""
    for (R, H, W, Cm, rs) in ((2, 3, 3, [[30.0, 4.0], [4.0, 20.0]], 2.0), (3, 5, 4, [[2.0, 0], [0, 50.0]], 0.7),
                              (4, 1, 9, [[3.0, 1.0], [1.0, 2.0]], 0.3)):
        ep = ExpParam(num_rotation_steps=R)
        rng = np.random.default_rng(R * 100 + H)
        child = (rng.standard_normal((R, H, W)) - 2).astype(np.float32)
        args = ((1.2, -0.7), (-0.4, 1.1), Cm, 0.4, rs, 1.0, False)
        want = oracle.message(ep, child, *args)
        with _ctx(ep, 2, H, W) as ctx:
            _cmp(ctx.message(child, *args), want, "tiny %dx%dx%d" % (R, H, W))


def test_all_log_zero_unaries_infer():
    """Every unary cell unevaluated.  In the reference this degenerates: the LOG_ZERO fill of the shifted grid is
    *above* the max of a -2e6 belief, exp overflows to inf and inf * 0 taps give NaN.  The device reproduces the same
    inf / NaN cells."""
    ep = ExpParam(num_rotation_steps=4, roi_save_num_samples=3)
    P, H, W = 3, 12, 10
    un = np.full((P, 1, 4, H, W), LZ, np.float32)
    joints = synth.make_joints(P, seed=2, max_offset=2, sigma_range=(1.0, 1.5))
    pc = synth.part_conf(P)
    want = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
    with PsContext(ep, pc, H, W) as ctx:
        res = od.computeRootPosteriorRot(ctx, [[un[p, 0]] for p in range(P)], joints, True, write_back_masked=False)
        assert np.array_equal(res.best_conf, want["best_conf"], equal_nan=True)
        for p in range(P):
            _cmp(ctx.marginal(p), want["marginals"][0, p], "all-LOG_ZERO marginal %d" % p)
        assert np.isinf(want["best_conf"][:, 6]).any() and np.isnan(want["marginals"]).any()


def test_degenerate_configs_are_rejected():
    from partapp_b200 import PsInferError, capi
    with pytest.raises(PsInferError) as e:     # min == max rotation needs exactly one bin (partapp_aux.hpp:29-31)
        PsContext(ExpParam(num_rotation_steps=4, min_part_rotation=0, max_part_rotation=0), synth.part_conf(2), 8, 8)
    assert e.value.status == capi.PS_ERR_INVALID
    with pytest.raises(PsInferError):          # strip_border_detections must be < 0.5 (findrot.cpp:529)
        PsContext(ExpParam(num_rotation_steps=4, strip_border_detections=0.5), synth.part_conf(2), 8, 8)
    # a single rotation bin with min == max: rot_step_size == 0 trips findrot.cpp:321
    ep = ExpParam(num_rotation_steps=1, min_part_rotation=0, max_part_rotation=0)
    with PsContext(ep, synth.part_conf(2), 8, 8) as ctx:
        with pytest.raises(PsInferError) as e:
            ctx.message(np.zeros((1, 8, 8), np.float32), (1, 1), (1, 1), [[2, 0], [0, 2]], 0.0, 0.5, 1.0, True)
        assert "rot_step_size" in str(e.value)


def test_dpm_score_fusion_matches_oracle():
    """a5: addDPMScore / addLoadDPMScore (objectdetect_icps.cpp:445-524), per-rotation and broadcast grids."""
    import ctypes
    ep = ExpParam(num_rotation_steps=8)
    P, H, W = 2, 22, 26
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 5))
    rng = np.random.default_rng(4)
    logg = (rng.standard_normal((8, H, W)) - 3).astype(np.float32)         # log-domain DPM prior, per rotation
    raw1 = rng.uniform(-0.2, 1.0, (1, H, W)).astype(np.float32)            # raw DPM scores, one grid for all rotations
    raw1[0, :3] = 5e-5                                                      # below the 1e-4 floor
    fp = ctypes.POINTER(ctypes.c_float)
    want = un.copy()
    L = oracle.lib()
    L.orc_add_dpm_score(want[0, 0].ctypes.data_as(fp), 8, H, W, logg.ctypes.data_as(fp), 8, 0.7)
    L.orc_add_load_dpm_score(want[1, 0].ctypes.data_as(fp), 8, H, W, raw1.ctypes.data_as(fp), 1, 1.3)
    with _ctx(ep, P, H, W) as ctx:
        for p in range(P):
            ctx.set_unary(p, 0, un[p, 0])
        ctx.add_unary_grid(0, logg, 0, 0.7)
        ctx.add_unary_grid(1, raw1, 1, 1.3)
        assert np.array_equal(ctx.get_unary(0, 0), want[0, 0])               # mode 0 is exact
        # mode 1 goes through logf: glibc's is < 1 ulp, the device's is correctly rounded -> allow rare last-bit cells
        _cmp(ctx.get_unary(1, 0), want[1, 0], "addLoadDPMScore", max_ulp_frac=1e-3, rtol=1e-6)


# ---- the other caller of the boundary: findObjectRoiHelper (objectdetect_roi.cpp:45-278) ---------------

def test_find_object_roi_helper_matches_oracle_composition():
    """Region-of-interest inference: TM_DIRECT ingest + clip, detection maxima on the SCORES, log, sparse inference at
    one forced scale, ROI offset on every hypothesis."""
    ep = ExpParam(num_rotation_steps=8, roi_save_num_samples=12)
    P, scale = 4, 1.3
    roi = (17, 9, 17 + 43, 9 + 51)                       # x1, y1, x2, y2 -> 44 x 52 cells
    W, H = roi[2] - roi[0] + 1, roi[3] - roi[1] + 1
    pc = synth.part_conf(P)
    joints = synth.make_joints(P, seed=21, max_offset=5, sigma_range=(1.5, 3))
    cells, Tig = synth.compact_scores(ep, H, W, P, 7)
    det, hyp = od.findObjectRoiHelper(ep, pc, roi, scale, [cells[p, 0] for p in range(P)], Tig, joints)

    ep1 = ExpParam(num_rotation_steps=8, roi_save_num_samples=12, min_object_scale=scale, max_object_scale=scale)
    un = np.empty((P, 1, ep.num_rotation_steps, H, W), np.float32)
    for p in range(P):
        g = oracle.load_score_grid(cells[p, 0], Tig, H, W)
        g[g < 0] = np.float32(0.0001)                    # clip_scores_fill (aux.hpp:42-59)
        want = oracle.find_local_max(g, 12)              # rows (rot, x, y, score), reference order
        assert len(want) == len(det[p]) and len(want) > 0
        got = det[p]
        assert sorted(map(tuple, got[:, [2, 4, 5, 6]].tolist())) == \
            sorted((r, x + roi[0], y + roi[1], v) for r, x, y, v in want.tolist())
        assert np.all(got[:, 0] == 0) and np.all(got[:, 1] == np.float32(scale))
        un[p, 0] = oracle.prepare_unary(g)               # log of the clipped scores
    ref = oracle.infer(ep1, pc, joints, un.copy(), sparse=True, want_marginals=False, want_hyps=True)
    for p in range(P):
        w = ref["part_hyps"][p].copy()
        w[:, 4] += roi[0]
        w[:, 5] += roi[1]
        assert np.array_equal(hyp[p][0], w[0]), "part %d argmax" % p
        assert hyp[p][0][1] == np.float32(scale)
        assert sorted(map(tuple, hyp[p][1:].tolist())) == sorted(map(tuple, w[1:].tolist())), "part %d local maxima" % p


# ---- legacy POS_GAUSSIAN message (objectdetect_findpos.cpp:64-89, multi_array_filter.hpp:335-369) ------

@pytest.mark.parametrize("diag", [False, True], ids=["full_cov", "diag_cov"])
@pytest.mark.parametrize("sparse", [False, True], ids=["bilinear", "direct"])
def test_pos_joint_marginal_matches_oracle(diag, sparse):
    ep = ExpParam(num_rotation_steps=3)
    H, W = 44, 38
    rng = np.random.default_rng(11)
    child = (rng.standard_normal((3, H, W)) * 2 - 3).astype(np.float32)
    if sparse:
        child[rng.random(child.shape) < 0.8] = -1e6          # exp -> exactly 0: skipped by the TM_DIRECT scatter
    th = 0.7
    Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    Cm = np.diag([9.0, 4.0]) if diag else Rm @ np.diag([10.0, 3.0]) @ Rm.T
    Cm = (Cm + Cm.T) / 2
    for offset, scale in (((5.5, -3.25), 1.0), ((-2.0, 7.0), 1.3)):
        want_parent, want_child = oracle.pos_message(child, offset, Cm, scale, sparse)
        with _ctx(ep, 2, H, W) as ctx:
            got_parent, got_child = od.computePosJointMarginal(ctx, child, offset, Cm, scale, sparse)
        _cmp(got_parent, want_parent, "pos message parent")
        _cmp(got_child, want_child, "pos message child round trip")
        assert (want_parent > -1e5).sum() > 100


# ---- ps_config.fast_math: fused multiply-add taps, inside the north star's float tolerance ---------------------------

def test_fast_math_mode_within_north_star_tolerance():
    """BASELINE.json: "argmax part positions and rotations bit-exact, and marginals within 1e-4 relative".
    The tolerance applies to this mode only; the default mode is compared for exact equality everywhere else.
    Log-marginals pass through zero, so "relative" is taken against max(|ref|, 1): |got - ref| <= 1e-4 * max(|ref|, 1)."""
    RTOL = ATOL = 1e-4
    ep = ExpParam(num_rotation_steps=24, roi_save_num_samples=5)
    P, H, W = 10, 72, 64
    pc = synth.part_conf(P)
    joints = synth.make_joints(P, seed=3, max_offset=8, sigma_range=(1.5, 4))
    for img in range(2):
        un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, img))
        want = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
        with PsContext(ep, pc, H, W, fast_math=True) as ctx:
            res = od.computeRootPosteriorRot(ctx, [[un[p, 0]] for p in range(P)], joints, True, write_back_masked=False)
            assert np.array_equal(res.best_conf[:, :6], want["best_conf"][:, :6]), "argmax records, image %d" % img
            np.testing.assert_allclose(res.best_conf[:, 6], want["best_conf"][:, 6], rtol=RTOL)
            exact = 0
            for p in range(P):
                got, ref = ctx.marginal(p), want["marginals"][0, p]
                np.testing.assert_allclose(got, ref, rtol=RTOL, atol=ATOL, err_msg="marginal of part %d" % p)
                exact += int(np.array_equal(got, ref))
            np.testing.assert_allclose(res.root_part_posterior, want["root_post"], rtol=RTOL, atol=ATOL)
        assert exact < P, "fast_math produced the parity bits everywhere: is the mode wired?"


# ---- the CUDA path against the reference's OWN code (recorded outputs of oracle/_ref, tests/golden/ref_drivers.npz) ----
# The oracle is pinned to the reference's compiled sources in tests/test_oracle_vs_ref.py; these tests close the loop
# without the oracle in between: what libpsinfer.so computes is compared directly with what objectdetect_findrot.cpp /
# objectdetect_aux.cpp / objectdetect_findpos.cpp / objectdetect_icps.cpp computed on the same inputs.
import os as _os
import sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden"))
import ref_driver_cases as _dc  # noqa: E402

_DRV = np.load(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "ref_drivers.npz"))


@pytest.mark.parametrize("name", sorted(_dc.message_cases()))
def test_gpu_message_equals_reference_code_output(name):
    c = _dc.message_cases()[name]
    child, oi, oo, Cm, rm, rs, sc, sp = _dc.message_args(c)
    with _ctx(c["ep"], 2, c["H"], c["W"]) as ctx:
        got = ctx.message(child, oi, oo, Cm, rm, rs, sc, sp)
    _cmp(got, _DRV["msg_" + name], "computeRotJointMarginal " + name)


@pytest.mark.parametrize("name", sorted(_dc.infer_cases()))
def test_gpu_inference_equals_reference_code_output(name):
    c = _dc.infer_cases()[name]
    pc, joints, un = _dc.infer_args(c)
    ep, P, S = c["ep"], c["P"], c["ep"].num_scale_steps
    with PsContext(ep, pc, c["H"], c["W"], keep_all_scales=True) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            for s in range(S):
                ctx.set_unary(p, s, un[p, s])
        ctx.infer(sparse=True, local_max=True)
        assert np.array_equal(ctx.best_conf(), _DRV["inf_%s_best" % name]), "argmax records"
        _cmp(ctx.root_posterior(), _DRV["inf_%s_root" % name], "root posterior")
        for s in range(S):
            for p in range(P):
                _cmp(ctx.marginal(p, s), _DRV["inf_%s_marg" % name][s, p], "marginal part %d scale %d" % (p, s))
                _cmp(ctx.get_unary(p, s), _DRV["inf_%s_masked" % name][p, s], "masked unary part %d scale %d" % (p, s))
        for p in range(P):
            got, want = ctx.part_hyps(p), _DRV["inf_%s_hyps%d" % (name, p)]
            assert got.shape == want.shape and np.array_equal(got[0], want[0]), "argmax record of part %d" % p
            # the local-maximum list: when more than K exist the reference keeps the K highest through std::sort, which
            # leaves the choice among equal scores at the cut unspecified -- scores must agree as a multiset,
            # records strictly above the lowest kept score must agree as a set
            g, w = got[1:], want[1:]
            assert sorted(g[:, 6].tolist()) == sorted(w[:, 6].tolist()), "local-maximum scores of part %d" % p
            if len(w):
                cut = w[:, 6].min()
                above = lambda a: sorted(map(tuple, a[a[:, 6] > cut].tolist()))
                assert above(g) == above(w), "local maxima above the cut, part %d" % p


@pytest.mark.parametrize("name", sorted(_dc.pos_message_cases()))
def test_gpu_pos_message_equals_reference_code_output(name):
    c = _dc.pos_message_cases()[name]
    ep = ExpParam(num_rotation_steps=c["child"].shape[0])
    with _ctx(ep, 2, c["child"].shape[1], c["child"].shape[2]) as ctx:
        par, ch = ctx.pos_message(c["child"], c["offset"], c["C"], c["scale"], c["sparse"])
    _cmp(par, _DRV["posmsg_%s_parent" % name], "computePosJointMarginal parent")
    _cmp(ch, _DRV["posmsg_%s_child" % name], "computePosJointMarginal child")


@pytest.mark.parametrize("name", ["rot", "torso_prior", "dpm_one_grid", "dpm_per_rotation"])
def test_gpu_conditioning_adds_equal_reference_code_output(name):
    """`pos` is left to the oracle-based test: its table holds the one pow(float, int) whose rounding depends on the
    C++ dialect; the other adds do not."""
    import ctypes
    c = _dc.condition_cases()[name]
    pc, un = _dc.condition_inputs(c)
    ep, P, S, R, H, W = c["ep"], c["P"], c["ep"].num_scale_steps, c["ep"].num_rotation_steps, c["H"], c["W"]
    L = oracle.lib()   # host table builders of the product are compared with the oracle's elsewhere; here: the adds
    with PsContext(ep, pc, H, W) as ctx:
        for p in range(P):
            for s in range(S):
                ctx.set_unary(p, s, un[p, s])
        if name == "rot":
            for p in range(P):
                t = ctx.rot_score_table(float(c["params"][p, 0]), float(c["params"][p, 1])) if pc.is_detect[p] else np.zeros(R, np.float32)
                ctx.add_unary_table(p, t, 0, c["weight"])
        elif name == "torso_prior":
            root = [p for p in range(P) if pc.is_detect[p] and pc.is_root[p]][0]
            t = ctx.torso_prior_table(*[float(v) for v in c["params"]], c["weight"])
            ctx.add_unary_table(root, t, 2)
        else:
            ctx.add_unary_grid(c["pidx"], c["dpm"], 0, c["weight"])
        for p in range(P):
            for s in range(S):
                _cmp(ctx.get_unary(p, s), _DRV["cond_" + name][p, s], "%s part %d scale %d" % (name, p, s))


# ---- legacy POS_GAUSSIAN driver (objectdetect_findpos.cpp:118-334) ---------------------------------------
from tests.golden import ref_pos_driver_cases as _pc


@pytest.mark.parametrize("name", sorted(_pc.cases()))
def test_gpu_pos_gaussian_driver_equals_reference_code_output(name):
    """ps_infer with POS_GAUSSIAN joints = mergeRotations + computeRootPosterior: the root posterior of every scale
    against what the reference's own code returned.  The rotation sum is x87 long double there and an error-free
    double pair here: identical except, at most, a handful of last-bit cells."""
    ep, pc, joints, un, sparse = _pc.cases()[name]
    want = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pos_driver.npz"))[name + "/root_post"]
    P, S, R, H, W = un.shape
    with PsContext(ep, pc, H, W) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            for s in range(S):
                ctx.set_unary(p, s, un[p, s])
        ctx.infer(sparse=sparse, root_hyps=True)
        got = ctx.root_posterior()
        assert got.shape == want.shape
        assert np.array_equal(np.isneginf(got), np.isneginf(want))
        fin = np.isfinite(want)
        neq = (got != want) & fin
        assert neq.mean() <= 1e-4, "%d cells differ" % neq.sum()
        np.testing.assert_allclose(got[fin], want[fin], rtol=2e-7, atol=0)
        hyps = ctx.root_hyps()
        lm = oracle.find_local_max(want, 1000)
        assert len(hyps) == len(lm) or neq.any()
        from partapp_b200 import capi
        with pytest.raises(capi.PsInferError):
            ctx.best_conf()       # no downward pass on this path


def test_disc_ps_message_passing_has_no_border_strip():
    """partSampleWithPriorHelper (libDiscPS/disc_sample_with_prior.cpp:64-330) is computeRootPosteriorRot +
    computePartMarginals with the upright masking but WITHOUT the root's border strip: PS_INFER_NO_BORDER_STRIP on a
    context whose ExpParam asks for the strip must give what the oracle gives with strip_border_detections = 0."""
    ep_strip = ExpParam(num_rotation_steps=8, strip_border_detections=0.2)
    ep_plain = ExpParam(num_rotation_steps=8)
    P, H, W = 4, 40, 44
    un = oracle.prepare_unary(synth.raw_scores(ep_plain, H, W, P, 9))
    joints = synth.make_joints(P, seed=2, max_offset=8, sigma_range=(1.5, 4))
    pc = synth.part_conf(P)
    want = oracle.infer(ep_plain, pc, joints, un.copy(), sparse=True)
    stripped = oracle.infer(ep_strip, pc, joints, un.copy(), sparse=True)
    _, root = synth.tree(P)
    assert not np.array_equal(want["marginals"][0, root], stripped["marginals"][0, root])   # the strip matters on this input
    with PsContext(ep_strip, pc, H, W) as ctx:
        ctx.set_joints(joints)
        for p in range(P):
            ctx.set_unary(p, 0, un[p, 0])
        ctx.infer(sparse=True, no_border_strip=True)
        assert np.array_equal(ctx.best_conf(), want["best_conf"])
        for p in range(P):
            _cmp(ctx.marginal(p), want["marginals"][0, p], "no-strip marginal %d" % p)
        for p in range(P):
            ctx.set_unary(p, 0, un[p, 0])
        ctx.infer(sparse=True)
        _cmp(ctx.marginal(root), stripped["marginals"][0, root], "stripped marginal")
