"""Python host mirror of the reference's `object_detect` inference interface, over the C ABI.

Names and argument meaning follow src/libs/libPictStruct/objectdetect.h (reference):
  computeRotJointMarginal  objectdetect.h:277-282 / objectdetect_findrot.cpp:292-456
  computeRootPosteriorRot  objectdetect.h:269-274 / objectdetect_findrot.cpp:470-727
  getMaxStates             objectdetect_findrot.cpp:73-110
  findLocalMax             objectdetect_aux.cpp:193-261
Grids are numpy float32 C-order [rotation][y][x] (host) -- exactly the reference's FloatGrid3 layout.
Where the reference asserts, these raise PsInferError.  All compute happens in libpsinfer.so on the GPU.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import capi

LOG_ZERO = np.float32(-1e6)  # libBoostMath/boost_math.h:23


@dataclass
class ExpParam:
    """The ExpParam fields the path reads (ExpParam.proto:189-195, :229, :282)."""
    num_rotation_steps: int = 48
    min_part_rotation: float = -180.0
    max_part_rotation: float = 180.0
    num_scale_steps: int = 1
    min_object_scale: float = 1.0
    max_object_scale: float = 1.0
    strip_border_detections: float = 0.0
    roi_save_num_samples: int = 1000
    interpolate: bool = False  # score grids mapped with TM_BILINEAR instead of TM_DIRECT (partapp.cpp:889-894)


@dataclass
class PartConf:
    """PartConfig.part[i].{is_detect,is_upright,is_root} (PartConfig.proto)."""
    is_detect: Sequence[bool]
    is_upright: Sequence[bool]
    is_root: Sequence[bool]

    @property
    def num_parts(self):
        return len(self.is_detect)


@dataclass
class Joint:
    """object_detect::Joint (objectdetect.h:54-86), 0-based part indices."""
    child_idx: int
    parent_idx: int
    offset_c: Sequence[float]
    offset_p: Sequence[float]
    C: Sequence[Sequence[float]]
    rot_mean: float = 0.0
    rot_sigma: float = 0.0
    type: int = capi.PS_JOINT_ROT_GAUSSIAN

    def to_c(self):
        j = capi.ps_joint()
        j.type = int(self.type)
        j.child_idx = int(self.child_idx)
        j.parent_idx = int(self.parent_idx)
        j.offset_c[0], j.offset_c[1] = float(self.offset_c[0]), float(self.offset_c[1])
        j.offset_p[0], j.offset_p[1] = float(self.offset_p[0]), float(self.offset_p[1])
        Cm = np.asarray(self.C, dtype=np.float64).reshape(4)
        for i in range(4):
            j.C[i] = float(Cm[i])
        j.rot_mean = float(self.rot_mean)
        j.rot_sigma = float(self.rot_sigma)
        return j

    @staticmethod
    def from_c(j):
        return Joint(j.child_idx, j.parent_idx, [j.offset_c[0], j.offset_c[1]], [j.offset_p[0], j.offset_p[1]],
                     [[j.C[0], j.C[1]], [j.C[2], j.C[3]]], j.rot_mean, j.rot_sigma, j.type)

    def flipped(self):
        """loadJoints' flip branch (objectdetect_aux.cpp:102-119)."""
        lib = capi.load_library()
        j = self.to_c()
        lib.ps_flip_joint(C.byref(j))
        return Joint.from_c(j)


def make_config(exp_param: ExpParam, part_conf: PartConf, height: int, width: int, device: int = 0,
                root_idx: int = -1, keep_all_scales: bool = False, fast_math: bool = False):
    cfg = capi.ps_config()
    cfg.device = device
    cfg.num_parts = part_conf.num_parts
    cfg.num_rotation_steps = exp_param.num_rotation_steps
    cfg.min_part_rotation = exp_param.min_part_rotation
    cfg.max_part_rotation = exp_param.max_part_rotation
    cfg.num_scale_steps = exp_param.num_scale_steps
    cfg.min_object_scale = exp_param.min_object_scale
    cfg.max_object_scale = exp_param.max_object_scale
    cfg.height = height
    cfg.width = width
    cfg.root_idx = root_idx
    for i in range(min(part_conf.num_parts, capi.PS_MAX_PARTS)):
        cfg.is_detect[i] = 1 if part_conf.is_detect[i] else 0
        cfg.is_upright[i] = 1 if part_conf.is_upright[i] else 0
        cfg.is_root[i] = 1 if part_conf.is_root[i] else 0
    cfg.strip_border_detections = exp_param.strip_border_detections
    cfg.roi_save_num_samples = int(exp_param.roi_save_num_samples)
    cfg.keep_all_scales = 1 if keep_all_scales else 0
    cfg.interpolate = 1 if getattr(exp_param, "interpolate", False) else 0
    cfg.fast_math = 1 if fast_math else 0
    return cfg


def rot_from_index(exp_param: ExpParam, idx: int) -> float:
    """partapp_aux.hpp:123-129"""
    cfg = make_config(exp_param, PartConf([True], [False], [True]), 1, 1)
    return capi.load_library().ps_rot_from_index(C.byref(cfg), idx)


def scale_from_index(exp_param: ExpParam, idx: int) -> float:
    """partapp_aux.hpp:86-92"""
    cfg = make_config(exp_param, PartConf([True], [False], [True]), 1, 1)
    return capi.load_library().ps_scale_from_index(C.byref(cfg), idx)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected grid of shape %s, got %s" % (shape, a.shape))
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PsContext:
    """One ps_ctx: all tree levels of one image resident on one GPU."""

    def __init__(self, exp_param: ExpParam, part_conf: PartConf, height: int, width: int, device: int = 0,
                 root_idx: int = -1, keep_all_scales: bool = False, fast_math: bool = False):
        self.lib = capi.load_library()
        self.exp_param = exp_param
        self.part_conf = part_conf
        self.cfg = make_config(exp_param, part_conf, height, width, device, root_idx, keep_all_scales, fast_math)
        self.R, self.S = exp_param.num_rotation_steps, exp_param.num_scale_steps
        self.H, self.W, self.P = height, width, part_conf.num_parts
        h = C.c_void_p()
        st = self.lib.ps_create(C.byref(self.cfg), C.byref(h))
        if st != capi.PS_OK:
            raise capi.PsInferError(st, (self.lib.ps_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.ps_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st):
        if st != capi.PS_OK:
            raise capi.PsInferError(st, (self.lib.ps_last_error(self.h) or b"").decode())

    # -- model -------------------------------------------------------------------------------------
    def set_joints(self, joints: Sequence[Joint]):
        arr = (capi.ps_joint * len(joints))(*[j.to_c() for j in joints])
        self._check(self.lib.ps_set_joints(self.h, arr, len(joints)))

    def set_stream(self, cuda_stream_handle: Optional[int]):
        self._check(self.lib.ps_set_stream(self.h, C.c_void_p(cuda_stream_handle or 0)))

    def synchronize(self):
        self._check(self.lib.ps_synchronize(self.h))

    # -- unaries -----------------------------------------------------------------------------------
    def set_unary(self, part, scale, grid, raw_scores=False):
        g = _f32(grid, (self.R, self.H, self.W))
        self._check(self.lib.ps_set_unary(self.h, part, scale, _ptr(g), capi.PS_MEM_HOST, int(raw_scores)))
        self.synchronize()  # g may be a temporary

    def set_unary_device(self, part, scale, dev_ptr, raw_scores=False):
        self._check(self.lib.ps_set_unary(self.h, part, scale, C.c_void_p(dev_ptr), capi.PS_MEM_DEVICE, int(raw_scores)))

    def set_unary_pinned(self, part, scale, host_ptr, raw_scores=False):
        """Asynchronous H2D from pinned host memory (caller keeps the buffer alive until synchronize)."""
        self._check(self.lib.ps_set_unary(self.h, part, scale, C.c_void_p(host_ptr), capi.PS_MEM_HOST, int(raw_scores)))

    def set_unary_compact(self, part, scale, cells, Tig, device_ptr=None):
        """loadScoreGrid on the device: cells [R][gh][gw] (0 = not evaluated), Tig [R][3][3] doubles."""
        Tig = np.ascontiguousarray(Tig, np.float64)
        if Tig.shape != (self.R, 3, 3):
            raise ValueError("Tig must be [R][3][3]")
        if device_ptr is not None:
            gh, gw = cells
            self._check(self.lib.ps_set_unary_compact(self.h, part, scale, C.c_void_p(device_ptr), gh, gw,
                                                      Tig.ctypes.data_as(C.POINTER(C.c_double)), capi.PS_MEM_DEVICE))
            return
        g = _f32(cells)
        if g.ndim != 3 or g.shape[0] != self.R:
            raise ValueError("cells must be [R][gh][gw]")
        self._check(self.lib.ps_set_unary_compact(self.h, part, scale, _ptr(g), g.shape[1], g.shape[2],
                                                  Tig.ctypes.data_as(C.POINTER(C.c_double)), capi.PS_MEM_HOST))
        self.synchronize()

    def set_unaries_compact(self, parts, scales, cells, Tig, gh=None, gw=None, pointers=None, device=False):
        """loadScoreGrid for several (part, scale) grids on ONE lattice in one fill + one scatter launch.  `cells`: list of
        [R][gh][gw] arrays, or -- with `pointers`, gh, gw -- raw addresses in device (device=True) or pinned host memory."""
        Tig = np.ascontiguousarray(Tig, np.float64)
        if Tig.shape != (self.R, 3, 3):
            raise ValueError("Tig must be [R][3][3]")
        n = len(parts)
        keep = None
        if pointers is None:
            keep = [_f32(c) for c in cells]
            gh, gw = keep[0].shape[1:]
            if any(k.shape != (self.R, gh, gw) for k in keep):
                raise ValueError("every grid must be [R][gh][gw] with the same gh, gw")
            pointers = [k.ctypes.data for k in keep]
        ps = (C.c_int * n)(*[int(p) for p in parts])
        ss = (C.c_int * n)(*[int(s) for s in scales])
        ptrs = (C.c_void_p * n)(*[int(p) for p in pointers])
        self._check(self.lib.ps_set_unaries_compact(self.h, n, ps, ss, ptrs, int(gh), int(gw),
                                                    Tig.ctypes.data_as(C.POINTER(C.c_double)),
                                                    capi.PS_MEM_DEVICE if device else capi.PS_MEM_HOST))
        if keep is not None:
            self.synchronize()

    def set_unary_compact_raw(self, part, scale, cells, Tig):
        """The same mapping stopped after clip_scores_fill (objectdetect_roi.cpp:205-215): scores, not logs."""
        Tig = np.ascontiguousarray(Tig, np.float64)
        g = _f32(cells)
        if g.ndim != 3 or g.shape[0] != self.R or Tig.shape != (self.R, 3, 3):
            raise ValueError("cells must be [R][gh][gw], Tig [R][3][3]")
        self._check(self.lib.ps_set_unary_compact_raw(self.h, part, scale, _ptr(g), g.shape[1], g.shape[2],
                                                      Tig.ctypes.data_as(C.POINTER(C.c_double)), capi.PS_MEM_HOST))
        self.synchronize()

    def log_unary(self, part, scale):
        """computeLogGrid in place on the resident grid (objectdetect_roi.cpp:240-242)."""
        self._check(self.lib.ps_log_unary(self.h, part, scale))

    def unary_local_max(self, part, scale, max_n):
        """findLocalMax on the resident grid: rows (rotidx, x, y, score)."""
        out = np.empty((max(max_n, 1), 4), np.float32)
        n = C.c_int()
        self._check(self.lib.ps_unary_local_max(self.h, part, scale, max_n, out.ctypes.data_as(C.POINTER(C.c_float)),
                                                C.byref(n)))
        return out[:n.value].copy()

    def set_unary_compact_pinned(self, part, scale, host_ptr, gh, gw, Tig):
        """Asynchronous variant for pinned host memory (caller keeps the buffer alive until synchronize)."""
        self._check(self.lib.ps_set_unary_compact(self.h, part, scale, C.c_void_p(host_ptr), gh, gw,
                                                  Tig.ctypes.data_as(C.POINTER(C.c_double)), capi.PS_MEM_HOST))

    def get_unary(self, part, scale):
        out = np.empty((self.R, self.H, self.W), np.float32)
        self._check(self.lib.ps_get_unary(self.h, part, scale, _ptr(out), capi.PS_MEM_HOST))
        return out

    def add_unary_table(self, part, table, kind, weight=1.0):
        t = _f32(table)
        self._check(self.lib.ps_add_unary_table(self.h, part, t.ctypes.data_as(C.POINTER(C.c_float)), kind, weight))

    def add_unary_tables(self, part, tables, kinds, weights, pointers=None, device=False):
        """Up to 4 conditioning tables applied in order in one pass (rotation score, position score, torso prior:
        findrot.cpp:913-949).  `tables`: numpy arrays, or -- with `pointers` -- raw addresses of float32 tables that
        already live in device memory (device=True) or in pinned host memory (device=False)."""
        n = len(kinds)
        keep = None
        if pointers is None:
            keep = [_f32(t) for t in tables]
            pointers = [t.ctypes.data for t in keep]
        ptrs = (C.c_void_p * n)(*[int(p) for p in pointers])
        ks = (C.c_int * n)(*[int(k) for k in kinds])
        ws = (C.c_float * n)(*[float(w) for w in weights])
        self._check(self.lib.ps_add_unary_tables(self.h, part, n, ptrs, ks, ws,
                                                 capi.PS_MEM_DEVICE if device else capi.PS_MEM_HOST))
        if keep is not None:
            self.synchronize()   # pageable sources are staged at call time; stay conservative for numpy temporaries

    # host builders of the conditioning tables (same arithmetic as the reference, which builds them on the CPU)
    def rot_score_table(self, mu, var):
        """getRotScoreGrid (objectdetect_icps.cpp:228-281): table[R] for one part."""
        t = np.empty(self.R, np.float32)
        self.lib.ps_rot_score_table(C.byref(self.cfg), float(mu), float(var), t.ctypes.data_as(C.POINTER(C.c_float)))
        return t

    def pos_score_table(self, mu_x, mu_y, var_x, var_y, root_x, root_y):
        """getPosScoreGrid (objectdetect_icps.cpp:366-423): table[H][W] for one non-root part."""
        t = np.empty((self.H, self.W), np.float32)
        self.lib.ps_pos_score_table(self.H, self.W, float(mu_x), float(mu_y), float(var_x), float(var_y), float(root_x),
                                    float(root_y), t.ctypes.data_as(C.POINTER(C.c_float)))
        return t

    def torso_prior_table(self, mu_x, mu_y, var_x, var_y, weight):
        """setTorsoPosPrior (objectdetect_icps.cpp:137-191): the weighted table[H][W] added to the root."""
        t = np.empty((self.H, self.W), np.float32)
        self.lib.ps_torso_prior_table(self.H, self.W, float(mu_x), float(mu_y), float(var_x), float(var_y), float(weight),
                                      t.ctypes.data_as(C.POINTER(C.c_float)))
        return t

    def add_unary_grid(self, part, grid, mode, weight=1.0):
        """addDPMScore (mode 0) / addLoadDPMScore (mode 1), objectdetect_icps.cpp:445-524; grid [1 or R][H][W]."""
        g = _f32(grid)
        if g.ndim == 2:
            g = g[None]
        self._check(self.lib.ps_add_unary_grid(self.h, part, _ptr(g), g.shape[0], mode, weight, capi.PS_MEM_HOST))

    # -- inference ---------------------------------------------------------------------------------
    def infer(self, sparse=True, local_max=False, root_hyps=False, keep_unaries=False, no_border_strip=False):
        flags = (capi.PS_INFER_SPARSE if sparse else 0) | (capi.PS_INFER_LOCAL_MAX if local_max else 0) | \
                (capi.PS_INFER_ROOT_HYPS if root_hyps else 0) | (capi.PS_INFER_KEEP_UNARIES if keep_unaries else 0) | \
                (capi.PS_INFER_NO_BORDER_STRIP if no_border_strip else 0)
        self._check(self.lib.ps_infer(self.h, flags))

    infer_async = infer  # ps_infer only enqueues device work; getters synchronise

    def max_states(self, local_max=False):
        self._check(self.lib.ps_max_states(self.h, capi.PS_INFER_LOCAL_MAX if local_max else 0))

    def best_conf(self):
        out = np.empty((self.P, capi.PS_HYP_VEC), np.float32)
        self._check(self.lib.ps_get_best_conf(self.h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def part_hyps(self, part):
        cap = int(self.exp_param.roi_save_num_samples) + 1
        out = np.empty((cap, capi.PS_HYP_VEC), np.float32)
        n = C.c_int()
        self._check(self.lib.ps_get_part_hyps(self.h, part, out.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return out[:n.value].copy()

    def marginal(self, part, scale=None):
        scale = self.S - 1 if scale is None else scale
        out = np.empty((self.R, self.H, self.W), np.float32)
        self._check(self.lib.ps_get_marginal(self.h, part, scale, _ptr(out), capi.PS_MEM_HOST))
        return out

    def root_posterior(self):
        out = np.empty((self.S, self.H, self.W), np.float32)
        self._check(self.lib.ps_get_root_posterior(self.h, _ptr(out), capi.PS_MEM_HOST))
        return out

    def root_hyps(self, cap=1000):
        out = np.empty((cap, 4), np.float32)
        n = C.c_int()
        self._check(self.lib.ps_get_root_hyps(self.h, out.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return out[:n.value].copy()

    def plan_info(self, joint, downward, scale=0):
        out = (C.c_int * 10)()
        self._check(self.lib.ps_get_plan_info(self.h, joint, int(downward), scale, out))
        keys = ("diag", "rows", "cols", "rot_taps", "x_taps", "y_taps", "rot_shift", "shift_flags", "x_cells", "y_cells")
        return dict(zip(keys, [int(v) for v in out]))

    def selftest_math(self, first_bits, count):
        """(exp mismatches, exp tested, log mismatches, log tested) over fp32 bit patterns [first, first+count)."""
        out = (C.c_ulonglong * 4)()
        self._check(self.lib.ps_selftest_math(self.h, first_bits, count, out))
        return tuple(int(v) for v in out)

    def eval_math(self, op, first_bits, count):
        """Device exp (op 0) / log (op 1) of the path on fp32 bit patterns [first, first + count)."""
        out = np.empty(count, np.float32)
        self._check(self.lib.ps_eval_math(self.h, op, first_bits, count, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def launch_count(self):
        return int(self.lib.ps_launch_count(self.h))

    def profile_enable(self, on=True):
        self._check(self.lib.ps_profile_enable(self.h, int(on)))

    def profile_read(self):
        """{kernel class: (total device ms, launches)} since the last read."""
        cap = 32
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        cnt = (C.c_longlong * cap)()
        n = C.c_int()
        self._check(self.lib.ps_profile_read(self.h, cap, names, ms, cnt, C.byref(n)))
        return {names[i].decode(): (ms[i], int(cnt[i])) for i in range(n.value)}

    # -- seams -------------------------------------------------------------------------------------
    def message(self, log_prob_child, offset_in, offset_out, Cm, rot_mean, rot_sigma, scale, sparse):
        g = _f32(log_prob_child, (self.R, self.H, self.W))
        out = np.empty_like(g)
        oi = (C.c_double * 2)(*[float(v) for v in offset_in])
        oo = (C.c_double * 2)(*[float(v) for v in offset_out])
        cc = (C.c_double * 4)(*[float(v) for v in np.asarray(Cm, np.float64).reshape(4)])
        self._check(self.lib.ps_message(self.h, _ptr(g), _ptr(out), capi.PS_MEM_HOST, oi, oo, cc, float(rot_mean),
                                        float(rot_sigma), float(scale), int(bool(sparse))))
        return out

    def pos_message(self, log_prob_child, offset, Cm, scale, sparse):
        """computePosJointMarginal on each [H][W] slice; returns (log_prob_parent, rewritten log_prob_child)."""
        g = _f32(log_prob_child, (self.R, self.H, self.W)).copy()
        out = np.empty_like(g)
        oo = (C.c_double * 2)(*[float(v) for v in offset])
        cc = (C.c_double * 4)(*[float(v) for v in np.asarray(Cm, np.float64).reshape(4)])
        self._check(self.lib.ps_pos_message(self.h, _ptr(g), _ptr(out), capi.PS_MEM_HOST, oo, cc, float(scale),
                                            int(bool(sparse))))
        return out, g

    def find_local_max(self, grid, max_n):
        g = _f32(grid)
        d0, h, w = g.shape
        out = np.empty((max(max_n, 1), 4), np.float32)
        n = C.c_int()
        self._check(self.lib.ps_find_local_max(self.h, _ptr(g), capi.PS_MEM_HOST, d0, h, w, max_n,
                                               out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n)))
        return out[:n.value].copy()


# ---- free functions with the reference's names ----------------------------------------------------

def computeRotJointMarginal(ctx: PsContext, log_prob_child, offset_c_10, offset_p_01, Cm, rot_mean, rot_sigma,
                            scale, bIsSparse):
    """objectdetect_findrot.cpp:292-456: returns log_prob_parent."""
    return ctx.message(log_prob_child, offset_c_10, offset_p_01, Cm, rot_mean, rot_sigma, scale, bIsSparse)


@dataclass
class RootPosteriorResult:
    root_part_posterior: np.ndarray                 # [S][H][W]
    best_conf: np.ndarray                           # [P][7], argmax of every part at the last scale
    best_part_hyp: List[np.ndarray] = field(default_factory=list)  # per part rows of 7 (argmax + local maxima)


def computeRootPosteriorRot(ctx: PsContext, log_part_detections, joints: Sequence[Joint], bIsSparse=True,
                            local_max=False, write_back_masked=True) -> RootPosteriorResult:
    """objectdetect_findrot.cpp:470-727.  log_part_detections[p][s] are [R][H][W] arrays; like the reference they
    are masked in place (upright slices, root border strip) when write_back_masked is set."""
    ctx.set_joints(joints)
    for p in range(ctx.P):
        for s in range(ctx.S):
            ctx.set_unary(p, s, log_part_detections[p][s])
    ctx.infer(sparse=bIsSparse, local_max=local_max)
    res = RootPosteriorResult(ctx.root_posterior(), ctx.best_conf())
    res.best_part_hyp = [ctx.part_hyps(p) for p in range(ctx.P)]
    if write_back_masked:
        for p in range(ctx.P):
            for s in range(ctx.S):
                log_part_detections[p][s][...] = ctx.get_unary(p, s)
    return res


def getMaxStates(ctx: PsContext, log_part_detections, local_max=False):
    """objectdetect_findrot.cpp:73-110 (use_pairwise: false)."""
    for p in range(ctx.P):
        ctx.set_unary(p, 0, log_part_detections[p][0])
    ctx.max_states(local_max=local_max)
    return ctx.best_conf(), [ctx.part_hyps(p) for p in range(ctx.P)]


def findLocalMax(ctx: PsContext, log_prob_grid, max_hypothesis_number):
    """objectdetect_aux.cpp:193-261: rows of (dim0, x, y, score)."""
    return ctx.find_local_max(log_prob_grid, max_hypothesis_number)


def computePosJointMarginal(ctx: PsContext, log_prob_child, offset, Cm, scale, bIsSparse):
    """objectdetect_findpos.cpp:64-89 (legacy POS_GAUSSIAN joints): returns (log_prob_parent, log_prob_child')."""
    return ctx.pos_message(log_prob_child, offset, Cm, scale, bIsSparse)


def findObjectRoiHelper(exp_param: ExpParam, part_conf: PartConf, roi, scale, score_grids, Tig, joints, device=0,
                        root_idx=-1):
    """The inference half of object_detect::findObjectRoiHelper (objectdetect_roi.cpp:45-278) for a region of interest
    whose detector responses are given: `roi` = (x1, y1, x2, y2) after the reference's border extension and clamping
    (:141-148), `score_grids[p]` = the compact ScoreGrid of part p, [R][gh][gw], `Tig` [R][3][3] its grid->ROI map
    (:201-203).  Computing the responses (computeDescriptorGridRoi / computeScoreGrid, :180-199) is the detector and
    stays outside.  Returns (best_part_det, best_part_hyp): per part, PartHyp rows
    [scaleidx, scale, rotidx, rot_deg, x, y, score] with the ROI offset added (:230-236, :265-271).

    The reference forces one scale equal to `scale` for the inference (:82-85) and runs computeRootPosteriorRot sparse
    without saved marginals (:246-262)."""
    x1, y1, x2, y2 = [int(v) for v in roi]
    W, H = abs(x2 - x1) + 1, abs(y2 - y1) + 1
    ep = ExpParam(**{**exp_param.__dict__, "min_object_scale": float(scale), "max_object_scale": float(scale),
                     "num_scale_steps": 1})
    K = int(ep.roi_save_num_samples)
    P = part_conf.num_parts
    with PsContext(ep, part_conf, H, W, device=device, root_idx=root_idx) as ctx:
        best_part_det = []
        for p in range(P):
            ctx.set_unary_compact_raw(p, 0, score_grids[p], Tig)            # TM_DIRECT + clip_scores_fill
            rows = ctx.unary_local_max(p, 0, K)                             # findLocalMax on the scores (:226-228)
            det = np.zeros((len(rows), 7), np.float32)
            for i, (r, x, y, v) in enumerate(rows):
                det[i] = [0, np.float32(scale_from_index(ep, 0)), r, np.float32(rot_from_index(ep, int(r))), x + x1, y + y1, v]
            best_part_det.append(det)
            ctx.log_unary(p, 0)                                             # computeLogGrid (:240-242)
        ctx.set_joints(joints)
        ctx.infer(sparse=True, local_max=True)
        best_part_hyp = []
        for p in range(P):
            h = ctx.part_hyps(p).copy()
            h[:, 4] += x1
            h[:, 5] += y1
            best_part_hyp.append(h)
    return best_part_det, best_part_hyp
