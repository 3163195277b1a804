"""CPU tests of the oracle (test infrastructure).  The reference ships no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned by (i) an independent float64 brute-force statement of one message
built from the analytic pairwise term of libDiscPS/factors.cpp:261-381, (ii) structural properties (flip symmetry,
filter delta response, eigen-decomposition), (iii) regression fixtures in tests/golden/ generated from the oracle."""
import math
import os

import numpy as np
import pytest

import oracle
from partapp_b200 import synth
from partapp_b200.objectdetect import ExpParam, Joint

LZ = np.float32(-1e6)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def taps(sigma):
    buf = np.zeros(2001)
    arr = np.ascontiguousarray(buf)
    n = oracle.lib().orc_gaussian_filter(float(sigma), arr.ctypes.data_as(oracle._dp), arr.size)
    return arr[:n].copy()


def test_gaussian_taps_definition():
    # boost_math.cpp:104-117: k = floor(3 sigma + .5), centre 1, exp(-i^2 / (2 sigma^2)), never normalised
    for sigma in (0.4, 1.0, 2.5, 7.3, 16.0):
        f = taps(sigma)
        k = int(math.floor(3 * sigma + 0.5))
        assert len(f) == 2 * k + 1 and f[k] == 1.0
        for i in range(1, k + 1):
            assert f[k + i] == f[k - i] == math.exp(-i * i / (2 * sigma * sigma))


def test_bin_centres():
    ep = ExpParam(num_rotation_steps=24)
    e = oracle.exp_param(ep)
    L = oracle.lib()
    # partapp_aux.hpp:45-58: min + step*(0.5 + idx)
    assert L.orc_rot_from_index(e, 0) == -172.5 and L.orc_rot_from_index(e, 23) == 172.5
    assert L.orc_index_from_rot(e, -1e-6) == 11 and L.orc_index_from_rot(e, 1e-6) == 12
    ep1 = ExpParam(num_rotation_steps=1, min_part_rotation=0, max_part_rotation=0)
    assert L.orc_rot_from_index(oracle.exp_param(ep1), 0) == 0.0  # min == max special case


def test_eig2d_reconstructs_covariance():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = rng.standard_normal((2, 2))
        Cm = a @ a.T + 0.1 * np.eye(2)
        Cm[1, 0] = Cm[0, 1]
        V = np.zeros(4)
        E = np.zeros(4)
        c = np.ascontiguousarray(Cm.reshape(4))
        oracle.lib().orc_eig2d(c.ctypes.data_as(oracle._dp), V.ctypes.data_as(oracle._dp), E.ctypes.data_as(oracle._dp))
        V = V.reshape(2, 2)
        E = E.reshape(2, 2)
        assert E[0, 0] <= E[1, 1]  # smallest eigenvalue first (boost_math.cpp:62-66)
        np.testing.assert_allclose(V @ E @ V.T, Cm, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(V.T @ V, np.eye(2), atol=1e-12)


def test_diag_filter_delta_response_is_unnormalised_outer_product():
    # filter.hpp:212-321: a delta of height p keeps height p at its centre; response = p * fy (x) fx, zero padded
    g = np.zeros((31, 29), np.float32)
    g[15, 14] = 3.0
    out = oracle.gauss_filter_2d(g, [[4.0, 0], [0, 9.0]], sparse=True)
    fx, fy = taps(2.0).astype(np.float32), taps(3.0).astype(np.float32)
    want = np.zeros_like(g)
    nx, ny = len(fx) // 2, len(fy) // 2
    for dy in range(-ny, ny + 1):
        for dx in range(-nx, nx + 1):
            want[15 + dy, 14 + dx] = np.float32(np.float32(3.0) * fx[nx + dx]) * fy[ny + dy]
    assert np.array_equal(out, want)
    assert out[15, 14] == 3.0


def test_border_truncation_is_zero_padding():
    g = np.ones((12, 10), np.float32)
    out = oracle.gauss_filter_2d(g, [[1.0, 0], [0, 1.0]], sparse=False)
    f = taps(1.0)
    full = f.sum()
    np.testing.assert_allclose(out[6, 5], full * full, rtol=1e-6)
    np.testing.assert_allclose(out[0, 0], f[3:].sum() ** 2, rtol=1e-6)  # window clipped, taps not renormalised


def _brute_force_message(ep, child, off_in, off_out, Cdiag, rot_mean, rot_sigma):
    """Independent float64 statement: analytic unnormalised Gaussian in the rounded joint-position offset and the
    rotation-index offset (libDiscPS/factors.cpp:261-381), truncated like the reference's windows."""
    R, H, W = child.shape
    M = float(child.max())
    step = (ep.max_part_rotation - ep.min_part_rotation) / R * math.pi / 180
    shift = int(math.floor(-rot_mean / step + 0.5))
    sx, sy = math.sqrt(Cdiag[0]), math.sqrt(Cdiag[1])
    nx, ny = int(math.floor(3 * sx + 0.5)), int(math.floor(3 * sy + 0.5))
    sig_idx = rot_sigma / step
    nr = int(math.floor(3 * sig_idx + 0.5))
    if 2 * nr + 1 >= R:
        nr = ((R - 2 if R % 2 else R - 1) - 1) // 2

    def rnd(v):
        return int(math.floor(v + 0.5))

    def trans(r, off):
        a = np.float32((ep.min_part_rotation + (ep.max_part_rotation - ep.min_part_rotation) / R * (r + 0.5)) * math.pi / 180)
        ca, sa = math.cos(float(a)), math.sin(float(a))
        return ca * off[0] - sa * off[1], sa * off[0] + ca * off[1]

    out = np.full((R, H, W), -1e6)
    prob = np.exp(child.astype(np.float64) - M)
    for rp in range(R):
        ux, uy = trans(rp, off_out)
        for yp in range(H):
            for xp in range(W):
                jx, jy = rnd(xp + ux), rnd(yp + uy)  # joint position seen from the parent
                if not (0 <= jx < W and 0 <= jy < H):
                    continue
                acc = 0.0
                for k in range(-nr, nr + 1):
                    ra = (rp + k) % R            # rotation slice after the mean shift (wraps in the blur)
                    rc = ra - shift              # child slice that was written there (no wrap)
                    if not 0 <= rc < R:
                        continue
                    wr = math.exp(-k * k / (2 * sig_idx * sig_idx))
                    tx, ty = trans(rc, off_in)
                    for dy in range(-ny, ny + 1):
                        y2 = jy + dy
                        if not 0 <= y2 < H:
                            continue
                        wy = math.exp(-dy * dy / (2 * sy * sy))
                        for dx in range(-nx, nx + 1):
                            x2 = jx + dx
                            if not 0 <= x2 < W:
                                continue
                            xc, yc = rnd(x2 - tx), rnd(y2 - ty)   # child cell whose joint sits at (x2, y2)
                            if 0 <= xc < W and 0 <= yc < H:
                                acc += prob[rc, yc, xc] * wr * wy * math.exp(-dx * dx / (2 * sx * sx))
                out[rp, yp, xp] = (math.log(acc) if acc > 0 else -1e6) + M
    return out


def test_message_equals_bruteforce_sum_product():
    ep = ExpParam(num_rotation_steps=6)
    R, H, W = 6, 11, 12
    rng = np.random.default_rng(3)
    child = (rng.standard_normal((R, H, W)) * 2 - 3).astype(np.float32)
    args = dict(off_in=(2.3, -1.6), off_out=(-1.2, 2.7), Cm=[[1.7, 0], [0, 0.9]], rot_mean=0.9, rot_sigma=1.1)
    got = oracle.message(ep, child, args["off_in"], args["off_out"], args["Cm"], args["rot_mean"], args["rot_sigma"],
                         1.0, True)
    want = _brute_force_message(ep, child, args["off_in"], args["off_out"], (1.7, 0.9), args["rot_mean"],
                                args["rot_sigma"])
    assert np.array_equal(got == LZ, want == -1e6)
    m = got != LZ
    np.testing.assert_allclose(got[m], want[m], rtol=2e-5, atol=2e-5)


def test_flip_symmetry_of_a_message():
    # Mirroring the child belief in x and flipping the joint (aux.cpp:102-119) mirrors the message, provided the
    # rotation bins are mirrored too (rot -> -rot maps bin r to R-1-r for a symmetric range).
    ep = ExpParam(num_rotation_steps=8)
    R, H, W = 8, 21, 21
    rng = np.random.default_rng(5)
    child = (rng.standard_normal((R, H, W)) * 2 - 3).astype(np.float32)
    j = Joint(0, 1, [3.0, -2.0], [-2.0, 4.0], [[5.0, 0.0], [0.0, 3.0]], 0.4, 0.7)
    oj = oracle.joint(j)
    oracle.lib().orc_flip_joint(oj)
    assert oj.offset_c[0] == -3.0 and oj.offset_p[0] == 2.0 and oj.rot_mean == -0.4 and oj.C[1] == -0.0
    a = oracle.message(ep, child, j.offset_c, j.offset_p, j.C, j.rot_mean, j.rot_sigma, 1.0, True)
    mirrored = np.ascontiguousarray(child[::-1, :, ::-1])
    b = oracle.message(ep, mirrored, [oj.offset_c[0], oj.offset_c[1]], [oj.offset_p[0], oj.offset_p[1]],
                       [[oj.C[0], oj.C[1]], [oj.C[2], oj.C[3]]], oj.rot_mean, oj.rot_sigma, 1.0, True)
    bm = b[::-1, :, ::-1]
    # integer rounding of the shifts is not mirror-symmetric at .5 ties, so compare away from LOG_ZERO borders
    both = (a != LZ) & (bm != LZ)
    assert both.mean() > 0.5
    np.testing.assert_allclose(a[both], bm[both], rtol=5e-2, atol=0.3)


def test_find_local_max_rule():
    # aux.cpp:203-228: ties allowed in the 8-neighbourhood, strict against both dim-0 neighbours, no wrap
    g = np.zeros((3, 5, 5), np.float32)
    g[1, 2, 2] = 1.0
    g[0, 0, 0] = 2.0
    lm = oracle.find_local_max(g, 1000)
    cells = {(int(r[0]), int(r[1]), int(r[2])) for r in lm}
    assert (1, 2, 2) in cells and (0, 0, 0) in cells
    assert all(int(r[0]) in (0, 1) for r in lm)  # the zero plateau of slice 2 is not strictly above slice 1... except under (1,2,2)
    g2 = np.zeros((1, 4, 4), np.float32)
    assert len(oracle.find_local_max(g2, 1000)) == 16  # a plateau: every cell qualifies
    assert len(oracle.find_local_max(g2, 5)) == 5


def test_argmax_is_first_maximum():
    g = np.zeros((2, 3, 4), np.float32)
    g[0, 1, 2] = 5
    g[1, 0, 0] = 5
    idx, val = oracle.argmax(g)
    assert idx == 1 * 4 + 2 and val == 5.0


def test_unary_prep():
    raw = np.array([[[-0.3, 0.0, 0.5, 1.0]]], np.float32)
    out = oracle.prepare_unary(raw)
    assert out[0, 0, 1] == LZ and out[0, 0, 3] == 0.0
    assert out[0, 0, 0] == np.float32(math.log(float(np.float32(0.0001))))
    assert out[0, 0, 2] == np.float32(math.log(0.5))


def test_golden_fixtures_reproduce():
    f = os.path.join(GOLD, "messages.npz")
    if not os.path.exists(f):
        pytest.skip("run tests/golden/make_golden.py")
    from tests.golden.make_golden import CASES, run_case
    z = np.load(f)
    for name in CASES:
        got = run_case(name)
        assert np.array_equal(got, z[name]), name
