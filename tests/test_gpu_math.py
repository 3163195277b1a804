"""Exhaustive check that the table-driven exp/log on the device equals CUDA's fp64 libm narrowed to fp32 for
every fp32 input (all 2^32 bit patterns), and that the latter equals this host's glibc on a large sample."""
import numpy as np
import pytest

from partapp_b200 import ExpParam, PsContext, synth

pytestmark = pytest.mark.gpu


def test_fast_exp_log_equal_slow_for_every_float():
    with PsContext(ExpParam(num_rotation_steps=8), synth.part_conf(2), 8, 8) as ctx:
        tot = [0, 0, 0, 0]
        step = 1 << 30
        for first in range(0, 1 << 32, step):
            r = ctx.selftest_math(first, step)
            tot = [a + b for a, b in zip(tot, r)]
        assert tot[0] == 0, "exp_fast differs from (float)exp((double)x) on %d inputs" % tot[0]
        assert tot[2] == 0, "log_fast differs from (float)log((double)x) on %d inputs" % tot[2]
        assert tot[1] > 2_000_000_000 and tot[3] > 2_000_000_000
