"""Deterministic synthetic inputs of LSP shape -- no real fixture exists in the reference.

Unaries mimic PartApp::loadScoreGrid output (reference libPartApp/partapp.cpp:830-903): classifier scores on a
stride-4 lattice whose phase depends on the rotation, exactly 0 elsewhere; after clip_scores_fill + computeLogGrid
every unevaluated cell is exactly -1e6.  Joints follow the 10-part tree of src/scripts/matlab/getJointsParts.m
(or the 22-part tree of getJointsParts22.m) with random offsets, full covariances and rotation Gaussians.
Everything is a pure function of (seed, image index), so the host oracle and the device see identical bytes.
"""
import numpy as np

from .objectdetect import ExpParam, Joint, PartConf

# getJointsParts.m:18-26 -- (child, parent), 0-based part ids, in joint order; root = part 4 (torso)
TREE10 = [(0, 1), (1, 4), (2, 4), (3, 2), (5, 4), (6, 7), (7, 4), (8, 4), (9, 8)]
ROOT10 = 4
# getJointsParts22.m:30-56; root = part 10
TREE22 = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 10), (5, 10), (6, 5), (7, 6), (8, 7), (9, 8), (11, 10), (12, 13),
          (13, 14), (14, 15), (15, 16), (16, 10), (17, 10), (18, 17), (19, 18), (20, 19), (21, 20)]
ROOT22 = 10


def tree(num_parts):
    if num_parts == 10:
        return TREE10, ROOT10
    if num_parts == 22:
        return TREE22, ROOT22
    # generic: a root (part 0) with chains of length <= 2
    edges = []
    for p in range(1, num_parts):
        edges.append((p, 0) if p % 2 == 1 else (p, p - 1))
    return edges, 0


def part_conf(num_parts, upright_root=False):
    _, root = tree(num_parts)
    return PartConf(is_detect=[True] * num_parts,
                    is_upright=[upright_root and p == root for p in range(num_parts)],
                    is_root=[p == root for p in range(num_parts)])


def make_joints(num_parts, seed=7, diagonal=False, max_offset=40.0, sigma_range=(4.0, 16.0), type_id=0):
    """Random generic spatial model.  `type_id` selects one entry of a per-joint type table (poselet-conditioned
    model: the whole joint is swapped per image, reference objectdetect_aux.cpp:76-99)."""
    edges, _ = tree(num_parts)
    joints = []
    for jidx, (c, p) in enumerate(edges):
        rng = np.random.default_rng([seed, jidx, type_id])
        oc = rng.uniform(-max_offset, max_offset, 2)
        op = rng.uniform(-max_offset, max_offset, 2)
        s1, s2 = rng.uniform(sigma_range[0], sigma_range[1], 2)
        if diagonal:
            Cm = np.array([[s1 * s1, 0.0], [0.0, s2 * s2]])
        else:
            th = rng.uniform(0.0, np.pi)
            Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            Cm = Rm @ np.diag([s1 * s1, s2 * s2]) @ Rm.T
            Cm[1, 0] = Cm[0, 1]  # exactly symmetric (eig2d asserts m21 == m12, boost_math.cpp:51)
        joints.append(Joint(child_idx=c, parent_idx=p, offset_c=oc.tolist(), offset_p=op.tolist(), C=Cm.tolist(),
                            rot_mean=float(rng.uniform(-0.6, 0.6)), rot_sigma=float(rng.uniform(0.2, 0.8))))
    return joints


def raw_scores(exp_param: ExpParam, height, width, num_parts, imgidx, seed=1234, stride=4, bump=True):
    """Classifier-score grids [P][S][R][H][W] float32 as loadScoreGrid would return them (0 = not evaluated)."""
    R, S = exp_param.num_rotation_steps, exp_param.num_scale_steps
    out = np.zeros((num_parts, S, R, height, width), np.float32)
    yy, xx = np.mgrid[0:height, 0:width]
    for p in range(num_parts):
        rng = np.random.default_rng([seed + imgidx, p])
        cx, cy = rng.uniform(0.25, 0.75) * width, rng.uniform(0.25, 0.75) * height
        cr = int(rng.integers(0, R))
        for s in range(S):
            for r in range(R):
                ox, oy = r % stride, (r // stride) % stride
                ny = (height - oy + stride - 1) // stride
                nx = (width - ox + stride - 1) // stride
                u = rng.random((ny, nx))
                v = (u ** 4).astype(np.float32)
                neg = rng.random((ny, nx)) < 0.25
                v = np.where(neg, -v, v)
                v[v == 0] = np.float32(1e-3)  # an evaluated cell is never exactly 0
                if bump:
                    dr = min((r - cr) % R, (cr - r) % R)
                    g = np.exp(-((xx[oy::stride, ox::stride] - cx) ** 2 + (yy[oy::stride, ox::stride] - cy) ** 2) /
                               (2 * 12.0 ** 2) - dr * dr / 2.0)
                    v = np.maximum(v, (0.98 * g).astype(np.float32))
                out[p, s, r, oy::stride, ox::stride] = v
    return out


def log_unaries(exp_param: ExpParam, height, width, num_parts, imgidx, seed=1234, stride=4):
    """raw_scores pushed through the reference's unary prep (findrot.cpp:834-845) in numpy: x<0 -> 1e-4,
    0 -> -1e6, else log evaluated in double and narrowed."""
    raw = raw_scores(exp_param, height, width, num_parts, imgidx, seed, stride)
    v = np.where(raw < 0, np.float32(0.0001), raw)
    with np.errstate(divide="ignore"):
        lg = np.log(v.astype(np.float64)).astype(np.float32)
    return np.where(v == 0, np.float32(-1e6), lg).astype(np.float32)


def compact_scores(exp_param: ExpParam, height, width, num_parts, imgidx, seed=1234, stride=4, rotated=False):
    """The same classifier scores as raw_scores(), in the form the detector stores them (reference
    libPartApp/partapp.cpp:830-903): per part and scale a compact grid `cells[R][gh][gw]` (0 = not evaluated) and the
    grid->image transforms `Tig[R][3][3]`.  rotated=False: axis-aligned lattices whose scatter reproduces raw_scores()
    exactly; rotated=True: each rotation's lattice is rotated about the image centre like the real detector's."""
    R, S = exp_param.num_rotation_steps, exp_param.num_scale_steps
    gh, gw = (height + stride - 1) // stride, (width + stride - 1) // stride
    cells = np.zeros((num_parts, S, R, gh, gw), np.float32)
    Tig = np.zeros((R, 3, 3), np.float64)
    raw = raw_scores(exp_param, height, width, num_parts, imgidx, seed, stride)
    for r in range(R):
        ox, oy = r % stride, (r // stride) % stride
        sub = raw[:, :, r, oy::stride, ox::stride]
        cells[:, :, r, :sub.shape[2], :sub.shape[3]] = sub
        if rotated:
            th = 2 * np.pi * (r + 0.5) / R - np.pi
            c, s_ = np.cos(th), np.sin(th)
            cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
            A = np.array([[c * stride, -s_ * stride], [s_ * stride, c * stride]])
            t = np.array([cx, cy]) - A @ np.array([(gw - 1) / 2.0, (gh - 1) / 2.0])
            Tig[r] = [[A[0, 0], A[0, 1], t[0]], [A[1, 0], A[1, 1], t[1]], [0, 0, 1]]
        else:
            Tig[r] = [[stride, 0, ox], [0, stride, oy], [0, 0, 1]]
    return cells, Tig
