"""Pins the oracle against the reference's own code.

oracle/_ref/libps_ref_core.so is the reference's libMultiArray (transform / filter / pointwise), libBoostMath (eig2d,
Gaussian taps, homogeneous coordinates) and libPartApp/partapp_aux.hpp (bin centres), compiled UNMODIFIED from
/root/reference against container stand-ins for Boost / Qt / BLAS (oracle/ref_shim/, oracle/ref_core.cpp).
tests/golden/ref_core.npz holds its outputs on the cases of tests/golden/ref_cases.py.

* everywhere (this container, the GPU box): the oracle must reproduce the recorded reference outputs bit for bit;
* where the library exists: it must still produce the recorded outputs, and on extra random inputs the oracle must agree
  with it.

What this does NOT pin: the BLAS summation order (the stand-in cblas_sdot is the Netlib order,
the convention the oracle states), the libm of the authors' machine, and the driver routines that compose these
primitives (computeRotJointMarginal, computeRootPosteriorRot, computePartMarginals), which live in a translation unit
that cannot be compiled here."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_cases  # noqa: E402
from oracle import refcore  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_core.npz"))
CASES = ref_cases.cases()


def _same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def test_fixture_covers_every_case():
    assert sorted(GOLD.files) == sorted(CASES)
    kinds = {k for k, _ in CASES.values()}
    assert kinds == {"taps", "eig", "inverse", "bbox", "transform_fixed", "gauss2d", "gauss2d_offset", "wraparound",
                     "pointwise", "bins"}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_output(name):
    kind, args = CASES[name]
    got = ref_cases.run(ref_cases.OracleBackend(), kind, args)
    assert _same(got, GOLD[name]), "%s (%s): oracle differs from the reference's own code" % (name, kind)


@pytest.mark.skipif(not refcore.available(), reason="oracle/_ref not built (needs the reference tree)")
def test_reference_core_still_matches_fixture():
    for name, (kind, args) in CASES.items():
        assert _same(ref_cases.run(refcore, kind, args), GOLD[name]), name


@pytest.mark.skipif(not refcore.available(), reason="oracle/_ref not built (needs the reference tree)")
@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_reference_core_on_random_inputs(seed):
    rng = np.random.default_rng(1000 + seed)
    ob = ref_cases.OracleBackend()
    h, w = int(rng.integers(9, 40)), int(rng.integers(9, 40))
    grid = rng.random((h, w)).astype(np.float32)
    grid[rng.random(grid.shape) < rng.uniform(0.0, 0.9)] = 0.0
    th, a, b = rng.uniform(0, np.pi), rng.uniform(1, 40), rng.uniform(1, 40)
    Cm = ref_cases.rot(th) @ np.diag([a, b]) @ ref_cases.rot(th).T
    Cm = (Cm + Cm.T) / 2
    for sparse in (False, True):
        assert _same(ob.gauss_filter_2d(grid, Cm, sparse), refcore.gauss_filter_2d(grid, Cm, sparse))
        off = rng.uniform(-8, 8, 2)
        assert _same(ob.gauss_filter_2d_offset(grid, Cm, off, sparse), refcore.gauss_filter_2d_offset(grid, Cm, off, sparse))
    T = np.eye(3)
    T[:2, :2] = rng.uniform(0.5, 4.0) * ref_cases.rot(rng.uniform(-np.pi, np.pi))
    T[:2, 2] = rng.uniform(-10, 10, 2)
    for m in (0, 1, 2):
        assert _same(ob.transform_fixed(grid, (h + 5, w + 3), T, 0.0, m), refcore.transform_fixed(grid, (h + 5, w + 3), T, 0.0, m))
    assert _same(ob.hc_inverse(T), refcore.hc_inverse(T))
    n = int(rng.integers(4, 60))
    sig = rng.uniform(0.3, 12.0)
    first, ln = ref_cases.clipped_taps(sig, n)
    taps = refcore.gaussian_filter(sig)[first:first + ln].astype(np.float32)
    col = rng.random(n).astype(np.float32)
    assert _same(ob.filter_1d_wraparound(col, taps), refcore.filter_1d_wraparound(col, taps))
