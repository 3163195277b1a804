"""Generates tests/golden/*.npz from the CPU oracle (the reference has no fixtures and cannot run here; see
oracle/ps_oracle.cpp).  Inputs are seeded, so the fixture pins the oracle against accidental change and gives the
GPU tests byte-level targets that do not need the oracle at run time.

    python -m tests.golden.make_golden
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from partapp_b200 import synth  # noqa: E402
from partapp_b200.objectdetect import ExpParam  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name: (R, H, W, seed, dense, off_in, off_out, C, rot_mean, rot_sigma, scale, sparse)
CASES = {
    "diag_sparse": (6, 20, 24, 1, False, (2.2, -3.1), (-1.4, 2.6), [[4.0, 0], [0, 2.25]], 0.5, 0.8, 1.0, True),
    "full_sparse": (6, 20, 24, 2, False, (3.0, 1.5), (-2.5, -2.0), [[5.0, 2.0], [2.0, 3.0]], -0.3, 0.6, 1.0, True),
    "full_dense": (6, 20, 24, 3, True, (3.0, 1.5), (-2.5, -2.0), [[5.0, -2.0], [-2.0, 3.0]], 0.3, 1.4, 1.0, False),
    "scaled": (8, 18, 22, 4, False, (2.0, 2.0), (1.0, -3.0), [[3.0, 1.0], [1.0, 4.0]], 0.0, 0.5, 1.15, True),
    "no_rot_blur": (6, 16, 16, 5, False, (1.0, 1.0), (-1.0, 2.0), [[2.0, 0], [0, 2.0]], 0.0, 0.0, 1.0, True),
}


def case_input(name):
    R, H, W, seed, dense = CASES[name][:5]
    ep = ExpParam(num_rotation_steps=R)
    if dense:
        rng = np.random.default_rng(seed)
        g = (rng.standard_normal((R, H, W)) * 3 - 5).astype(np.float32)
    else:
        g = oracle.prepare_unary(synth.raw_scores(ep, H, W, 1, seed)[0, 0])
    return ep, g


def run_case(name):
    ep, g = case_input(name)
    oi, oo, Cm, rm, rs, sc, sparse = CASES[name][5:]
    return oracle.message(ep, g, oi, oo, Cm, rm, rs, sc, sparse)


def infer_case():
    ep = ExpParam(num_rotation_steps=6, roi_save_num_samples=10)
    P, H, W = 4, 24, 20
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 7))
    joints = synth.make_joints(P, seed=4, max_offset=5, sigma_range=(1.2, 2.5))
    pc = synth.part_conf(P)
    return ep, pc, joints, un


def main():
    np.savez_compressed(os.path.join(HERE, "messages.npz"), **{n: run_case(n) for n in CASES})
    ep, pc, joints, un = infer_case()
    res = oracle.infer(ep, pc, joints, un.copy(), sparse=True)
    np.savez_compressed(os.path.join(HERE, "infer.npz"), best_conf=res["best_conf"], root_post=res["root_post"],
                        marginals=res["marginals"])
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
