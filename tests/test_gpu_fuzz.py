"""Randomised whole inferences through the C ABI against the CPU oracle (tools/gpu_fuzz.py): sizes that are and are not
multiples of four, 8 to 48 rotations, 2 to 6 parts, one or two scales, full / diagonal / mixed covariances, short and long
filters, dense and compact ingest (twice on one context), upright masks, border strips.  Exact comparison."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,cases,maxdim", [(11, 40, 150), (12, 6, 400)])
def test_random_inferences_are_bit_identical_to_the_oracle(seed, cases, maxdim, tmp_path, monkeypatch):
    spec = importlib.util.spec_from_file_location(
        "gpu_fuzz", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "gpu_fuzz.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.chdir(tmp_path)            # the tool writes gpurun_out/fuzz_<seed>.txt relative to the cwd
    assert mod.main(cases, seed, maxdim) == 0, open(tmp_path / "gpurun_out" / ("fuzz_%d.txt" % seed)).read()
