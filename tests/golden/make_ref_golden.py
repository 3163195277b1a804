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

# ---- the drivers: objectdetect_findrot.cpp / objectdetect_aux.cpp as the reference compiled them --------------------
import ref_driver_cases as dc  # noqa: E402

if not refcore.drivers_available():
    sys.exit("oracle/_ref/libps_ref_drivers.so is missing")
drv = {}
for name, c in dc.message_cases().items():
    child, oi, oo, Cm, rm, rs, sc, sp = dc.message_args(c)
    drv["msg_" + name] = refcore.message(c["ep"], child, oi, oo, Cm, rm, rs, sc, sp)
for name, c in dc.infer_cases().items():
    pc, joints, un = dc.infer_args(c)
    r = refcore.infer(c["ep"], pc, joints, un, sparse=True)
    drv["inf_%s_best" % name] = r["best_conf"]
    drv["inf_%s_root" % name] = r["root_post"]
    drv["inf_%s_marg" % name] = r["marginals"]
    drv["inf_%s_masked" % name] = un
    for p, h in enumerate(r["part_hyps"]):
        drv["inf_%s_hyps%d" % (name, p)] = h
for name, c in dc.local_max_cases().items():
    drv["lm_" + name] = refcore.find_local_max(c["grid"], c["K"])
for name, joints in dc.joint_cases().items():
    P = len(joints) + 1
    for flip in (0, 1):
        rows, di = refcore.load_joints(P, joints, flip)
        drv["joints_%s_flip%d" % (name, flip)] = rows
        drv["joints_%s_flip%d_detinv" % (name, flip)] = di
for name, c in dc.condition_cases().items():
    pc, un = dc.condition_inputs(c)
    drv["cond_" + name] = refcore.condition(c["ep"], pc, un, c["kind"], c.get("params"), c["weight"], c.get("pidx", 0), c.get("dpm"))
for name, c in dc.pos_message_cases().items():
    par, ch = refcore.pos_message(c["child"], c["offset"], c["C"], c["scale"], c["sparse"])
    drv["posmsg_%s_parent" % name] = par
    drv["posmsg_%s_child" % name] = ch
np.savez_compressed(os.path.join(HERE, "ref_drivers.npz"), **drv)
print("wrote %d reference driver outputs" % len(drv))
