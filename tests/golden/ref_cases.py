"""The cases on which the oracle is pinned against the reference's own numeric core (oracle/_ref, built from
/root/reference by `make -C oracle ref`).  Shared by make_ref_golden.py (which records the reference's outputs into
ref_core.npz) and by the tests (which replay the inputs through the oracle)."""
import numpy as np

TM_NEAREST, TM_BILINEAR, TM_DIRECT = 0, 1, 2


def rot(th):
    return np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])


def cases():
    """name -> (kind, args).  Everything is derived from fixed seeds."""
    rng = np.random.default_rng(20131)
    out = {}
    for i, s in enumerate([0.2, 0.5, 1.0, 2.37, 4.0, 7.9, 16.0, 33.3]):
        out["taps_%d" % i] = ("taps", dict(sigma=s))
    covs = [np.diag([9.0, 4.0]), np.diag([4.0, 9.0]), np.diag([5.0, 5.0])]
    for th, a, b in [(0.3, 16.0, 4.0), (1.2, 3.0, 30.0), (2.5, 9.0, 9.5), (-0.7, 2.0, 2.5), (0.0001, 8.0, 3.0)]:
        c = rot(th) @ np.diag([a, b]) @ rot(th).T
        covs.append((c + c.T) / 2)
    for i, c in enumerate(covs):
        out["eig_%d" % i] = ("eig", dict(C=c))
    # affine maps: rotations + scalings + translations, integer and fractional
    Ts = []
    for th, s, tx, ty in [(0.0, 1.0, 3.0, -2.0), (0.0, 1.0, 2.5, 1.25), (0.4, 1.0, 5.0, 3.0), (-1.1, 1.3, 10.0, 12.0),
                          (0.0, 4.0, 1.0, 2.0), (0.2, 0.6, 4.0, 1.5)]:
        T = np.eye(3)
        T[:2, :2] = s * rot(th)
        T[:2, 2] = [tx, ty]
        Ts.append(T)
    for i, T in enumerate(Ts):
        out["inv_%d" % i] = ("inverse", dict(T=T))
        out["bbox_%d" % i] = ("bbox", dict(T=T, w=37, h=29))
    dense = rng.random((29, 37)).astype(np.float32)
    sparse = dense.copy()
    sparse[rng.random(sparse.shape) < 0.85] = 0.0
    for i, T in enumerate(Ts):
        for m, name in ((TM_NEAREST, "nearest"), (TM_BILINEAR, "bilinear"), (TM_DIRECT, "direct")):
            src = sparse if m == TM_DIRECT else dense
            out["xform_%s_%d" % (name, i)] = ("transform_fixed", dict(grid=src, out_shape=(33, 41), T=T, default=0.0, method=m))
    out["xform_default_nearest"] = ("transform_fixed", dict(grid=dense, out_shape=(20, 50), T=Ts[2], default=-1e6, method=TM_NEAREST))
    for i, c in enumerate(covs):
        out["g2d_dense_%d" % i] = ("gauss2d", dict(grid=dense, C=c, sparse=False))
        out["g2d_sparse_%d" % i] = ("gauss2d", dict(grid=sparse, C=c, sparse=True))
    for i, (c, off) in enumerate([(covs[0], (3.0, -2.0)), (covs[3], (5.5, 1.25)), (covs[4], (-4.0, 6.0)), (covs[2], (0.0, 0.0))]):
        out["g2doff_dense_%d" % i] = ("gauss2d_offset", dict(grid=dense, C=c, offset=off, sparse=False))
        out["g2doff_sparse_%d" % i] = ("gauss2d_offset", dict(grid=sparse, C=c, offset=off, sparse=True))
    for i, (n, sig) in enumerate([(24, 0.8), (24, 2.4), (24, 9.0), (48, 3.1), (8, 1.0), (13, 5.0)]):
        col = rng.random(n).astype(np.float32)
        out["wrap_%d" % i] = ("wraparound", dict(col=col, sigma=sig))
    for i, (mn, mx, n) in enumerate([(-180.0, 180.0, 24), (-180.0, 180.0, 48), (-22.5, 22.5, 3), (0.8, 1.2, 5), (0.0, 360.0, 7)]):
        out["bins_%d" % i] = ("bins", dict(mn=mn, mx=mx, n=n))
    vals = np.concatenate([rng.random(200).astype(np.float32), np.float32([0.0, 1.0, 1e-30, 3e38, 1e-45])])
    out["log_grid"] = ("pointwise", dict(op=0, a=vals))
    out["unary_prep"] = ("pointwise", dict(op=5, a=np.concatenate([vals, -vals[:50], np.float32([-0.0, -1e-30])])))
    out["exp_grid"] = ("pointwise", dict(op=1, a=np.concatenate([(-rng.random(200) * 120).astype(np.float32),
                                                                 np.float32([0.0, -103.9, -104.1, 88.0, -1e6])])))
    return out


def clipped_taps(sigma, n):
    """The tap vector computeRotJointMarginal hands to the wrap-around filter: tails clipped to < n (findrot.cpp:385-390)."""
    k = int(np.floor(3 * sigma + 0.5))
    ln = 2 * k + 1
    first = 0
    if ln >= n:
        c = ln // 2
        ln = n - 2 if n % 2 == 1 else n - 1
        first = c - ln // 2
    return first, ln


def run(backend, kind, a):
    """Evaluates one case with `backend` = the reference core (oracle.refcore) or the oracle adapter below."""
    if kind == "taps":
        return backend.gaussian_filter(a["sigma"])
    if kind == "eig":
        V, E = backend.eig2d(a["C"])
        return np.concatenate([V.ravel(), E.ravel()])
    if kind == "inverse":
        return backend.hc_inverse(a["T"])
    if kind == "bbox":
        return backend.transformed_bbox(a["T"], a["w"], a["h"])
    if kind == "transform_fixed":
        return backend.transform_fixed(a["grid"], a["out_shape"], a["T"], a["default"], a["method"])
    if kind == "gauss2d":
        return backend.gauss_filter_2d(a["grid"], a["C"], a["sparse"])
    if kind == "gauss2d_offset":
        return backend.gauss_filter_2d_offset(a["grid"], a["C"], a["offset"], a["sparse"])
    if kind == "wraparound":
        taps = backend.gaussian_filter(a["sigma"])
        first, ln = clipped_taps(a["sigma"], a["col"].size)
        return backend.filter_1d_wraparound(a["col"], taps[first:first + ln].astype(np.float32))
    if kind == "pointwise":
        return backend.pointwise(a["op"], a["a"])
    if kind == "bins":
        return backend.bins(a["mn"], a["mx"], a["n"])
    raise KeyError(kind)


class OracleBackend:
    """The same calls answered by the oracle's restatement (oracle/ps_oracle.cpp)."""

    def __init__(self):
        import ctypes as C
        import oracle
        self.o, self.C = oracle, C
        self.L = oracle.lib()
        self.dp, self.fp = C.POINTER(C.c_double), C.POINTER(C.c_float)

    def _d(self, a):
        a = np.ascontiguousarray(a, np.float64)
        return a, a.ctypes.data_as(self.dp)

    def gaussian_filter(self, sigma):
        out = np.empty(4096, np.float64)
        n = self.L.orc_gaussian_filter(float(sigma), out.ctypes.data_as(self.dp), out.size)
        return out[:n].copy()

    def eig2d(self, Cm):
        _c, pc = self._d(Cm)
        V, E = np.empty(4), np.empty(4)
        self.L.orc_eig2d(pc, V.ctypes.data_as(self.dp), E.ctypes.data_as(self.dp))
        return V.reshape(2, 2), E.reshape(2, 2)

    def hc_inverse(self, T):
        _t, pt = self._d(T)
        out = np.empty(9)
        self.L.orc_hc_inverse(pt, out.ctypes.data_as(self.dp))
        return out.reshape(3, 3)

    def transformed_bbox(self, T, w, h):
        _t, pt = self._d(T)
        out = np.empty(4)
        self.L.orc_transformed_bbox(pt, int(w), int(h), out.ctypes.data_as(self.dp))
        return out

    def transform_fixed(self, grid, out_shape, T, default_value, method):
        g = np.ascontiguousarray(grid, np.float32)
        out = np.empty(out_shape, np.float32)
        _t, pt = self._d(T)
        self.L.orc_transform_fixed(g.ctypes.data_as(self.fp), g.shape[0], g.shape[1], out.ctypes.data_as(self.fp),
                                   out.shape[0], out.shape[1], pt, self.C.c_float(default_value), int(method))
        return out

    def gauss_filter_2d(self, grid, Cm, sparse):
        return self.o.gauss_filter_2d(grid, Cm, sparse)

    def gauss_filter_2d_offset(self, grid, Cm, offset, sparse):
        g = np.ascontiguousarray(grid, np.float32)
        out = np.empty_like(g)
        _c, pc = self._d(Cm)
        _o, po = self._d(offset)
        self.L.orc_gauss_filter_2d_offset(g.ctypes.data_as(self.fp), out.ctypes.data_as(self.fp), g.shape[0], g.shape[1], pc,
                                          po, int(bool(sparse)))
        return out

    def filter_1d_wraparound(self, col, taps):
        c = np.ascontiguousarray(col, np.float32)
        f = np.ascontiguousarray(taps, np.float32)
        out = np.empty_like(c)
        self.L.orc_filter_1d_wraparound(c.ctypes.data_as(self.fp), out.ctypes.data_as(self.fp), c.size,
                                        f.ctypes.data_as(self.fp), f.size)
        return out

    def pointwise(self, op, a):
        a = np.ascontiguousarray(a, np.float32).copy()
        if op == 5:
            return self.o.prepare_unary(a)
        self.L.orc_pointwise(int(op), a.ctypes.data_as(self.fp), a.size)
        return a

    def bins(self, mn, mx, n):
        from types import SimpleNamespace
        ep = SimpleNamespace(num_rotation_steps=n, min_part_rotation=mn, max_part_rotation=mx, num_scale_steps=1,
                             min_object_scale=1.0, max_object_scale=1.0, strip_border_detections=0.0, roi_save_num_samples=1)
        e = self.o.exp_param(ep)
        centres = np.array([self.L.orc_rot_from_index(self.C.byref(e), i) for i in range(n)])
        back = np.array([float(self.L.orc_index_from_rot(self.C.byref(e), float(c))) for c in centres])
        return np.concatenate([centres, back])
