"""Driver-level cases on which the oracle is pinned against the reference's own objectdetect_findrot.cpp /
objectdetect_aux.cpp (oracle/_ref/libps_ref_drivers.so).  Shared by make_ref_golden.py and tests/test_oracle_vs_ref.py."""
import numpy as np

from partapp_b200 import ExpParam, synth


def message_cases():
    out = {}
    ep = ExpParam(num_rotation_steps=8)
    H, W = 30, 26
    jt = synth.make_joints(6, seed=11, max_offset=6, sigma_range=(1.5, 3))
    jd = synth.make_joints(6, seed=12, max_offset=6, sigma_range=(1.5, 3), diagonal=True)
    for name, j, sparse, up in [("full_sparse_up", jt[0], True, True), ("full_dense_down", jt[1], False, False),
                                ("diag_sparse_up", jd[2], True, True), ("diag_dense_down", jd[3], False, False),
                                ("full_dense_up_scaled", jt[4], False, True)]:
        out[name] = dict(ep=ep, H=H, W=W, joint=j, sparse=sparse, up=up, scale=1.25 if "scaled" in name else 1.0,
                         seed=len(out))
    return out


def message_args(c):
    """(child grid, off_in, off_out, C, rot_mean, rot_sigma, scale, sparse) exactly as the passes call the routine."""
    import oracle
    j = c["joint"]
    child = oracle.prepare_unary(synth.raw_scores(c["ep"], c["H"], c["W"], 1, c["seed"])[0, 0])
    if not c["sparse"]:
        rng = np.random.default_rng(c["seed"])
        child = (rng.standard_normal(child.shape) * 3 - 5).astype(np.float32)
    if c["up"]:
        return child, j.offset_c, j.offset_p, j.C, j.rot_mean, j.rot_sigma, c["scale"], c["sparse"]
    return child, j.offset_p, j.offset_c, j.C, -j.rot_mean, j.rot_sigma, c["scale"], c["sparse"]


def infer_cases():
    out = {}
    out["p4"] = dict(P=4, ep=ExpParam(num_rotation_steps=8, roi_save_num_samples=6), H=28, W=24, upright=False, seed=3)
    out["p6_two_scales_upright_strip"] = dict(
        P=6, ep=ExpParam(num_rotation_steps=8, roi_save_num_samples=4, num_scale_steps=2, min_object_scale=0.9,
                         max_object_scale=1.1, strip_border_detections=0.1), H=26, W=24, upright=True, seed=4)
    out["p10_r12"] = dict(P=10, ep=ExpParam(num_rotation_steps=12, roi_save_num_samples=3), H=30, W=28, upright=False, seed=5)
    return out


def infer_args(c):
    import oracle
    pc = synth.part_conf(c["P"], upright_root=c["upright"])
    joints = synth.make_joints(c["P"], seed=c["seed"], max_offset=5, sigma_range=(1.5, 3))
    un = oracle.prepare_unary(synth.raw_scores(c["ep"], c["H"], c["W"], c["P"], c["seed"]))
    return pc, joints, un


def local_max_cases():
    rng = np.random.default_rng(77)
    out = {}
    for name, (D, H, W, K) in {"all": (8, 18, 20, 1000), "top7": (8, 18, 20, 7), "tiny": (3, 7, 7, 2), "one_slice": (1, 12, 12, 50)}.items():
        g = rng.standard_normal((D, H, W)).astype(np.float32)
        g[rng.random(g.shape) < 0.2] = np.float32(-1e6)      # plateaus of LOG_ZERO: ties in the 8-neighbourhood rule
        out[name] = dict(grid=g, K=K)
    return out


def joint_cases():
    return {"tree10": synth.make_joints(10, seed=9), "tree6_diag": synth.make_joints(6, seed=2, diagonal=True)}
