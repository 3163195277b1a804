#!/usr/bin/env python
"""Records the outputs of the reference's OWN numeric core on the cases of ref_cases.py into ref_core.npz.

Needs oracle/_ref/libps_ref_core.so, i.e. the reference tree (`make -C oracle ref` compiles libMultiArray /
libBoostMath / partapp_aux.hpp from /root/reference, unmodified, against the container stand-ins of oracle/ref_shim/).
The fixture travels; tests/test_oracle_vs_ref.py replays the inputs through the oracle everywhere and demands the same
bits.      python tests/golden/make_ref_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_cases  # noqa: E402
from oracle import refcore  # noqa: E402

if not refcore.available():
    sys.exit("oracle/_ref/libps_ref_core.so is missing: run `make -C oracle ref` where /root/reference exists")
out = {name: ref_cases.run(refcore, kind, a) for name, (kind, a) in ref_cases.cases().items()}
np.savez_compressed(os.path.join(HERE, "ref_core.npz"), **out)
print("wrote %d reference outputs" % len(out))
