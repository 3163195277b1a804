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


def condition_cases():
    """The conditioning adds.  Parameters are in the ranges the MATLAB predictors produce: rotation
    mean / variance in radians, position mean / variance in pixels relative to the detected root."""
    ep = ExpParam(num_rotation_steps=8, num_scale_steps=2, min_object_scale=0.9, max_object_scale=1.1)
    P, H, W = 4, 22, 26
    rng = np.random.default_rng(31)
    rot = np.stack([rng.uniform(-1.5, 1.5, P), rng.uniform(0.05, 0.8, P)], 1)
    pos = np.stack([rng.uniform(-6, 6, P), rng.uniform(-6, 6, P), rng.uniform(4, 60, P), rng.uniform(4, 60, P)], 1)
    root = np.array([12.0, 9.0])
    dpm1 = np.log(rng.uniform(1e-3, 1.0, (1, H, W))).astype(np.float32)
    dpmR = np.log(rng.uniform(1e-3, 1.0, (8, H, W))).astype(np.float32)
    base = dict(ep=ep, P=P, H=H, W=W, seed=6)
    return {"rot": dict(base, kind=0, params=rot, weight=0.35),
            "pos": dict(base, kind=1, params=np.concatenate([pos.ravel(), root]), weight=0.6),
            "torso_prior": dict(base, kind=2, params=np.array([1.0, -2.0, 40.0, 65.0]), weight=0.8),
            "dpm_one_grid": dict(base, kind=3, pidx=2, dpm=dpm1, weight=0.5),
            "dpm_per_rotation": dict(base, kind=3, pidx=1, dpm=dpmR, weight=1.5)}


def condition_inputs(c):
    import oracle
    pc = synth.part_conf(c["P"])
    un = oracle.prepare_unary(synth.raw_scores(c["ep"], c["H"], c["W"], c["P"], c["seed"]))
    return pc, un


def oracle_condition(c):
    """The same adds composed from the oracle's restatement (tables built like the reference, then broadcast adds).
    oracle/_ref is compiled as C++17, where pow(float, int) squares in double; the oracle's default is the authors'
    gnu++98 rule (fp32 square) -- see ps_oracle.cpp.  The comparison runs the oracle under the C++11 rule."""
    import ctypes as C
    import oracle
    L = oracle.lib()
    L.orc_set_pow_dialect(1)
    try:
        return _oracle_condition(c, L, C, oracle)
    finally:
        L.orc_set_pow_dialect(0)


def _oracle_condition(c, L, C, oracle):
    fp = C.POINTER(C.c_float)
    pc, un = condition_inputs(c)
    u = un.copy()
    P, S, R, H, W = u.shape
    root = [p for p in range(P) if pc.is_detect[p] and pc.is_root[p]][0]
    f = lambda a: a.ctypes.data_as(fp)
    if c["kind"] == 0:
        for p in range(P):
            t = np.zeros(R, np.float32)
            if pc.is_detect[p]:
                L.orc_rot_score_table(C.byref(oracle.exp_param(c["ep"])), float(c["params"][p, 0]), float(c["params"][p, 1]), f(t))
            for s in range(S):
                L.orc_add_rot_table(f(u[p, s]), R, H, W, f(t), C.c_float(c["weight"]))
    elif c["kind"] == 1:
        prm = c["params"][:4 * P].reshape(P, 4)
        rx, ry = c["params"][4 * P:]
        for p in range(P):
            t = np.zeros((H, W), np.float32)
            if pc.is_detect[p] and p != root:
                L.orc_pos_score_table(H, W, *[float(v) for v in prm[p]], float(rx), float(ry), f(t))
            for s in range(S):
                L.orc_add_pos_table(f(u[p, s]), R, H, W, f(t), C.c_float(c["weight"]))
    elif c["kind"] == 2:
        t = np.zeros((H, W), np.float32)
        L.orc_torso_prior_table(H, W, *[float(v) for v in c["params"]], C.c_float(c["weight"]), f(t))
        for s in range(S):
            L.orc_add_pos_table_unweighted(f(u[root, s]), R, H, W, f(t))
    else:
        g = np.ascontiguousarray(c["dpm"], np.float32)
        for s in range(S):
            L.orc_add_dpm_score(f(u[c["pidx"], s]), R, H, W, f(g), g.shape[0], C.c_float(c["weight"]))
    return u


def pos_message_cases():
    """The legacy POS_GAUSSIAN message (objectdetect_findpos.cpp:64-89) on [D][H][W] stacks of 2-D grids."""
    rng = np.random.default_rng(11)
    H, W = 34, 30
    dense = (rng.standard_normal((2, H, W)) * 2 - 3).astype(np.float32)
    sparse = dense.copy()
    sparse[rng.random(sparse.shape) < 0.8] = np.float32(-1e6)
    th = 0.7
    Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    full = Rm @ np.diag([10.0, 3.0]) @ Rm.T
    full = (full + full.T) / 2
    out = {}
    for cname, Cm in (("diag", np.diag([9.0, 4.0])), ("full", full)):
        for gname, g, sp in (("dense", dense, False), ("sparse", sparse, True)):
            for i, (off, sc) in enumerate((((5.5, -3.25), 1.0), ((-2.0, 7.0), 1.3))):
                out["%s_%s_%d" % (cname, gname, i)] = dict(child=g, offset=off, C=Cm, scale=sc, sparse=sp)
    return out
