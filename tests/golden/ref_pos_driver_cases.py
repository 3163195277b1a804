"""Cases of the legacy POS_GAUSSIAN driver (mergeRotations + computeRootPosterior, objectdetect_findpos.cpp:118-334),
shared by the fixture writer and the tests."""
import numpy as np

import oracle
from partapp_b200 import ExpParam, Joint, PartConf, synth


def cases():
    out = {}
    # a: the generic 6-part tree (root with chains of two), two scales, sparse lattice unaries
    ep = ExpParam(num_rotation_steps=8, num_scale_steps=2, min_object_scale=0.9, max_object_scale=1.1)
    P, H, W = 6, 40, 36
    joints = synth.make_joints(P, seed=3, max_offset=6, sigma_range=(1.5, 3))
    for j in joints:
        j.type = 1
    un = oracle.prepare_unary(synth.raw_scores(ep, H, W, P, 1))
    out["tree6_two_scales_sparse"] = (ep, synth.part_conf(P), joints, un, True)
    # b: a branching inner node, one undetected leaf (its branch sends no message) and dense unaries, bilinear warp
    ep = ExpParam(num_rotation_steps=4)
    P, H, W = 5, 32, 44
    edges = [(1, 0), (2, 1), (3, 1), (4, 0)]            # part 1 has two children
    rng = np.random.default_rng(8)
    joints = []
    for c, p in edges:
        th = rng.uniform(0, np.pi)
        Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        Cm = Rm @ np.diag(rng.uniform(2, 9, 2)) @ Rm.T
        Cm[1, 0] = Cm[0, 1]
        joints.append(Joint(child_idx=c, parent_idx=p, offset_c=rng.uniform(-5, 5, 2).tolist(),
                            offset_p=rng.uniform(-5, 5, 2).tolist(), C=Cm.tolist(), rot_mean=0.0, rot_sigma=0.0, type=1))
    pc = PartConf(is_detect=[True, True, True, True, False], is_upright=[False] * P, is_root=[True] + [False] * (P - 1))
    un = (rng.standard_normal((P, 1, 4, H, W)) * 2 - 4).astype(np.float32)
    out["branching_dense_undetected_leaf"] = (ep, pc, joints, un, False)
    return out
