"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): argmax part positions and rotations bit-exact; marginals within 1e-4 relative.
The kernels are written to be bit-identical to the oracle (same fp32 summation order, fp64 exp/log), so most
checks here are exact equality with a tiny allowance for exp/log last-bit differences between glibc and CUDA libm
(counted and bounded, see _cmp).
"""
import numpy as np
import pytest

import oracle
from partapp_b200 import ExpParam, Joint, PartConf, PsContext, synth
from partapp_b200 import objectdetect as od

pytestmark = pytest.mark.gpu

LZ = np.float32(-1e6)


def _cmp(got, want, what, max_ulp_frac=1e-4, rtol=1e-4):
    """Exact support of LOG_ZERO; values within rtol everywhere; report the fraction of non-identical cells,
    which must stay below max_ulp_frac (libm last-bit differences only)."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, what
    assert np.array_equal(got == LZ, want == LZ), what + ": LOG_ZERO support differs"
    neq = got != want
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
            _cmp(ctx.get_unary(p, 0), want[p, 0], "unary prep", max_ulp_frac=1e-5)


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
