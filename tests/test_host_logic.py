"""Host logic of the dataset loop and its multi-process sharding (gloo, world_size 2, CPU)."""
import os
import socket
import sys

import numpy as np
import pytest

from partapp_b200 import dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_first_last_and_shards_cover_the_range():
    assert dataset.init_firstidx_lastidx(1000) == (0, 999)
    assert dataset.init_firstidx_lastidx(1000, 10, 5) == (10, 14)
    assert dataset.init_firstidx_lastidx(12, 10, 50) == (10, 11)
    for n, world in ((1000, 8), (7, 2), (3, 8), (1, 1)):
        seen = []
        for r in range(world):
            lo, hi = dataset.shard_range(0, n - 1, r, world)
            seen += list(range(lo, hi + 1))
        assert seen == list(range(n))


def test_output_file_names():
    assert dataset.get_object_hyp_filename(12, False) == "/object_hyp_imgidx12_o0_spmnone.pbuf"
    assert dataset.pose_est_filename(7) == "/pose_est_imgidx0007.mat"


def _fake_infer(imgidx, flip):
    out = np.zeros((3, 7), np.float32)
    out[:, 4] = imgidx
    out[:, 5] = 2 * imgidx + int(flip)
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def gather(packed):
        t = torch.from_numpy(packed)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [o.numpy() for o in outs]

    res = dataset.find_object_dataset(_fake_infer, 0, 6, flip_orientation=True, rank=rank, world=world, gather=gather)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, sorted(res.keys()), float(sum(v[:, 5].sum() for v in res.values()))))


def test_two_rank_shard_and_gather_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    want_keys = sorted((i, f) for i in range(7) for f in (False, True))
    want_sum = float(sum(3 * (2 * i + int(f)) for i, f in want_keys))
    for rank, keys, total in got:
        assert keys == want_keys, "rank %d gathered %s" % (rank, keys)
        assert total == want_sum


def test_single_rank_loop_without_gather():
    res = dataset.find_object_dataset(_fake_infer, 2, 4)
    assert sorted(res) == [(2, False), (3, False), (4, False)]


def test_pcp_rule():
    from partapp_b200 import parteval
    pp = parteval.PartParam(window_size_x=20, window_size_y=60, pos_offset_x=10, pos_offset_y=30)
    gt = [0, 1.0, 11, -7.5, 100, 200, 0.0]
    assert parteval.is_gt_match(gt, gt, pp)
    near = [0, 1.0, 11, -7.5, 110, 210, 0.0]          # 14 px off both endpoints, threshold 0.5 * 60 = 30
    far = [0, 1.0, 11, -7.5, 140, 200, 0.0]           # 40 px off
    rot = [0, 1.0, 17, 82.5, 100, 200, 0.0]           # a quarter turn moves the endpoints by ~42 px
    assert parteval.is_gt_match(gt, near, pp) and not parteval.is_gt_match(gt, far, pp)
    assert not parteval.is_gt_match(gt, rot, pp)
    assert parteval.pcp_identical([gt, gt], [near, far], [pp]) == 0.5


# ---- eval_segments: the reference's PCP evaluation (parteval.cpp:1162-1345) ----------------------------

_PART_CONF = """
# two sticks: a vertical limb between annopoints 1-2, a horizontal one between 3-4
part { part_id: 1 part_pos: 1 part_pos: 2 part_x_axis_from: 1 part_x_axis_to: 2 part_x_axis_offset: -90
       ext_y_pos: 6 ext_y_neg: 4 is_root: true }
part { part_id: 2 part_pos: 3 part_pos: 4 part_x_axis_from: 3 part_x_axis_to: 4 part_x_axis_offset: -90 }
joint { child_idx: 2 parent_idx: 1 type: "Gaussian" }
"""
# the detection window of part 1 is the 100 px stick plus the part's y extension (4 above, 6 below), like the windows
# the reference derives from the part configuration; the evaluation strips the extension again (parteval.cpp:836-837)
_WINDOW_PARAM = """
part { part_id: 1 window_size_x: 40 window_size_y: 110 pos_offset_x: 20 pos_offset_y: 54 }
part { part_id: 2 window_size_x: 30 window_size_y: 80 pos_offset_x: 15 pos_offset_y: 40 }
"""
_ANNOLIST = """<annotationlist>
<annotation><image><name>im0.png</name></image>
  <annorect><x1>10</x1><y1>10</y1><x2>200</x2><y2>200</y2>
    <annopoints>
      <point><id>1</id><x>100</x><y>50</y><is_visible>1</is_visible></point>
      <point><id>2</id><x>100</x><y>150</y><is_visible>1</is_visible></point>
      <point><id>3</id><x>60</x><y>120</y><is_visible>1</is_visible></point>
      <point><id>4</id><x>140</x><y>120</y><is_visible>1</is_visible></point>
    </annopoints>
  </annorect>
</annotation>
<annotation><image><name>im1.png</name></image>
  <annorect><x1>0</x1><y1>0</y1><x2>1</x2><y2>1</y2>
    <annopoints>
      <point><id>1</id><x>100</x><y>50</y></point>
      <point><id>2</id><x>100</x><y>150</y></point>
    </annopoints>
  </annorect>
</annotation>
</annotationlist>
"""


def test_eval_segments_pcp(tmp_path):
    from partapp_b200 import parteval as pe
    (tmp_path / "part_conf.txt").write_text(_PART_CONF)
    (tmp_path / "window_param.txt").write_text(_WINDOW_PARAM)
    (tmp_path / "test.al").write_text(_ANNOLIST)
    conf = pe.load_part_conf(str(tmp_path / "part_conf.txt"))
    win = pe.load_window_param(str(tmp_path / "window_param.txt"))
    annos = pe.load_annolist(str(tmp_path / "test.al"))
    assert [a.image for a in annos] == ["im0.png", "im1.png"] and annos[0].rects[0].points[4] == (140, 120)
    assert conf[0].part_pos == [1, 2] and conf[0].part_x_axis_offset == -90 and win[1].window_size_y == 80

    gt = pe.get_part_bbox(annos[0].rects[0], conf[0], 1.0)
    np.testing.assert_allclose(gt.part_pos, [100, 100])
    np.testing.assert_allclose(gt.part_y_axis, [0, 1], atol=1e-12)          # the limb direction
    assert abs(gt.min_proj_y - (-50 - 4)) < 1e-9 and abs(gt.max_proj_y - (50 + 6)) < 1e-9

    def conf_rows(x0, rot0, x1, y1, rot1):
        # best_conf rows: [scaleidx, scale, rotidx, rot_deg, x, y, score]
        return np.array([[0, 1.0, 0, rot0, x0, 100, 0.5], [0, 1.0, 0, rot1, x1, y1, 0.25]], np.float32)

    runs = {0: conf_rows(100, 0.0, 100, 120, -90.0),   # both sticks exactly on the annotation
            1: conf_rows(151, 0.0, 0, 0, 0.0)}         # 51 px off with a 100 px stick: miss; part 2 not annotated
    res = pe.eval_segments(annos, conf, win, lambda i: runs[i], 0, 1, save_dir=str(tmp_path / "seg_endpoints"))
    assert (res.seg_correct, res.seg_total) == (2, 3) and res.per_part_total == [2, 1] and res.per_part_correct == [1, 1]
    assert abs(res.ratio - 2 / 3) < 1e-12
    np.testing.assert_allclose(res.endpoints[0][0], [100, 150, 100, 50, 1, 1], atol=1e-9)   # bottom, top, match, has gt
    np.testing.assert_allclose(res.endpoints[0][1, :4], [140, 120, 60, 120], atol=1e-4)
    assert res.endpoints[1][1, 5] == 0 and res.endpoints[1][2, 0] == 0.0
    import scipy.io
    np.testing.assert_allclose(scipy.io.loadmat(str(tmp_path / "seg_endpoints" / "endpoints_0001.mat"))["endpoints"],
                               res.endpoints[1])
    # the 0.5 * length threshold is strict on both ends
    runs[1] = conf_rows(149, 0.0, 0, 0, 0.0)
    assert pe.eval_segments(annos, conf, win, lambda i: runs[i], 1, 1).seg_correct == 1
    runs[1] = conf_rows(100, 90.0, 0, 0, 0.0)          # right place, wrong orientation
    assert pe.eval_segments(annos, conf, win, lambda i: runs[i], 1, 1).seg_correct == 0


def test_model_parts_to_evaluation_parts():
    """vis_eval_helper's conversions (parteval.cpp:520-857): joint parts merge into limbs whose stick ends are the
    joints; directly copied parts lose the window's y extension; the 14-part model snaps the limb axis to a bin."""
    from partapp_b200 import parteval as pe
    pp = pe.PartParam(window_size_x=30, window_size_y=40, pos_offset_x=15, pos_offset_y=20)

    def box(x, y, rot=0.0, scale=1.0):
        return pe.bbox_from_hyp(np.array([0, scale, 0, rot, x, y, 1.0], np.float32), pp)

    ev = [pe.PartDef(part_id=i + 1, ext_x_pos=15.0, ext_y_pos=5.0, ext_y_neg=3.0) for i in range(10)]
    # human_full_joints: 18 joint parts; limb 0 = parts 0 (lower joint) and 1 (upper joint) of a vertical stick
    pos = [(50 + 7 * i, 100 + 11 * i) for i in range(18)]
    pos[0], pos[1] = (60, 200), (60, 140)
    boxes = [box(x, y) for x, y in pos]
    model = [pe.PartDef(part_id=i + 1, ext_x_pos=10.0 + i) for i in range(18)]
    out = pe.convert_eval_bboxes("human_full_joints", boxes, [1.0] * 18, ev, model)
    assert len(out) == 10
    top, bot, seg = pe.get_bbox_endpoints(out[0])
    np.testing.assert_allclose(out[0].part_pos, [60, 170])
    np.testing.assert_allclose([top, bot], [[60, 140], [60, 200]], atol=1e-9)      # the two joints
    assert abs(seg - 60) < 1e-9
    # torso (model part 8) and head (17) are copied, their windows shortened by the evaluation part's extension
    assert abs(out[4].min_proj_y - (-20 + 3)) < 1e-9 and abs(out[4].max_proj_y - (20 - 5)) < 1e-9
    np.testing.assert_allclose(out[5].part_pos, pos[17])
    # x shrink: 0.7 * 15 / ext_x_pos of model part pidx (30 for the root slot 4, 20 for slot 5)
    assert abs(out[0].max_proj_x - 15 * float(np.float32(0.7 * 15 / 10.0))) < 1e-6
    assert abs(out[4].max_proj_x - 15 * float(np.float32(0.7 * 30 / 14.0))) < 1e-6
    # inputs are not modified
    assert boxes[8].min_proj_y == -20

    # human_full_torso4: the torso is the mean of four corner parts
    pos22 = [(10 * i, 5 * i) for i in range(22)]
    pos22[16], pos22[17], pos22[18], pos22[19] = (100, 200), (140, 200), (140, 100), (100, 100)   # lower-left .. upper-left
    out = pe.convert_eval_bboxes("human_full_torso4", [box(x, y) for x, y in pos22], [1.0] * 22, ev,
                                 [pe.PartDef(ext_x_pos=12.0)] * 22)
    np.testing.assert_allclose(out[4].part_pos, [120, 150])
    assert (out[4].min_proj_y, out[4].max_proj_y, out[4].min_proj_x, out[4].max_proj_x) == (-50.0, 50.0, -20.0, 20.0)

    # human_full_14_parts: the limb axis is the rotation bin centre nearest to atan2 of the joint difference
    pos14 = [(20 * i, 300 - 9 * i) for i in range(14)]
    pos14[0], pos14[1] = (100, 100), (100, 160)          # difference (0, -60): atan2 = -90 deg
    out = pe.convert_eval_bboxes("human_full_14_parts", [box(x, y) for x, y in pos14], [1.0] * 14, ev,
                                 [pe.PartDef(ext_x_pos=12.0)] * 14, rot_range=(-180.0, 180.0, 24))
    assert len(out) == 10
    ang = np.degrees(np.arctan2(out[0].part_y_axis[1], out[0].part_y_axis[0]))
    assert abs(ang - (-82.5)) < 1e-9 or abs(ang - (-97.5)) < 1e-9     # bins are 15 deg wide, centres at +-7.5 + 15 k
    assert abs(ang - (-97.5)) < 1e-9                                   # first minimum in ascending bin order
    assert (out[0].min_proj_y, out[0].max_proj_y) == (-30.0, 30.0)     # the larger of the x / y spans, centred

    # every other type: one evaluation part per model part, extension scaled by the hypothesis scale
    out = pe.convert_eval_bboxes("human_full", [box(5, 5, scale=2.0)] * 10, [2.0] * 10, ev, ev)
    assert abs(out[3].min_proj_y - (-40 + 6)) < 1e-9 and abs(out[3].max_proj_y - (40 - 10)) < 1e-9


def test_eval_segments_experiment_layout(tmp_path):
    """The experiment-level entry reads the same files, in the same places, as `partapp --eval_segments`."""
    import scipy.io
    from partapp_b200 import parteval as pe
    (tmp_path / "part_conf.txt").write_text(_PART_CONF)
    (tmp_path / "test.al").write_text(_ANNOLIST)
    sub = tmp_path / "logs" / "exp-pcp"
    (sub / "class").mkdir(parents=True)
    (sub / "part_marginals").mkdir()
    (sub / "class" / "window_param.txt").write_text("train_object_height: 200\n" + _WINDOW_PARAM)
    (tmp_path / "exp-pcp.txt").write_text('test_dataset: "test.al"\nlog_dir: "./logs"\npart_conf: "part_conf.txt"\n'
                                          "num_rotation_steps: 24\n")
    rows0 = np.array([[0, 1, 0, 0, 100, 100, 1], [0, 1, 0, -90, 100, 120, 1]], np.float32)
    rows1 = np.array([[0, 1, 0, 0, 180, 100, 1], [0, 1, 0, 0, 0, 0, 1]], np.float32)
    for i, rows in enumerate((rows0, rows1)):
        scipy.io.savemat(str(sub / "part_marginals" / ("pose_est_imgidx%04d.mat" % i)), {"best_conf": rows})
    r = pe.eval_segments_experiment(str(tmp_path / "exp-pcp.txt"))
    assert (r.seg_correct, r.seg_total) == (2, 3)
    assert (sub / "part_marginals" / "seg_endpoints" / "endpoints_0001.mat").exists()
    r = pe.eval_segments_experiment(str(tmp_path / "exp-pcp.txt"), first=1, numimgs=1, save_endpoints=False)
    assert (r.seg_correct, r.seg_total) == (0, 1)


# ---- Gaussian work lists: brute-force coverage on the CPU ---------------------------------------

def _work_lists(H, W, Cm, scale=1.0):
    import ctypes
    from partapp_b200 import ExpParam, capi
    from partapp_b200.objectdetect import PartConf, make_config
    lib = capi.load_library()
    cfg = make_config(ExpParam(num_rotation_steps=8), PartConf([True], [False], [True]), H, W)
    dims = (ctypes.c_int * 6)()
    T34 = (ctypes.c_double * 6)()
    cap = 1 << 16
    xl, yl = (ctypes.c_int * (3 * cap))(), (ctypes.c_int * (3 * cap))()
    Cc = (ctypes.c_double * 4)(*np.asarray(Cm, np.float64).ravel())
    rc = lib.ps_plan_work_lists(ctypes.byref(cfg), Cc, scale, dims, T34, xl, yl, cap)
    assert rc == 0
    EH, EW, nx, ny, nxl, nyl = [int(v) for v in dims]
    xs = np.array(xl[:3 * nxl], np.int64).reshape(-1, 3)
    ys = np.array(yl[:3 * nyl], np.int64).reshape(-1, 3)
    return EH, EW, nx, ny, np.array(T34[:]), xs, ys


@pytest.mark.parametrize("k", range(10))
def test_work_lists_cover_every_cell_the_read_back_touches(k):
    """plan_message's work lists against a brute-force statement of what is needed: the bilinear read-back of every
    image pixel touches up to 2 x 2 eigen-frame cells (transform.hpp:196-238) -- all of them must lie in a listed
    group of the y pass, and every x-pass output such a cell's y window reaches must lie in a listed group of the x
    pass.  Also: the lists stay well below the full bounding box."""
    rng = np.random.default_rng(900 + k)
    H, W = [(150, 230), (260, 140), (64, 64), (600, 400), (33, 500), (129, 191), (400, 37), (300, 300), (90, 70),
            (512, 256)][k]
    th = rng.uniform(0, np.pi) if k else 0.6
    s1, s2 = rng.uniform(1.0, 12.0, 2)
    Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    Cm = Rm @ np.diag([s1 * s1, s2 * s2]) @ Rm.T
    Cm[1, 0] = Cm[0, 1]
    scale = [1.0, 1.2, 0.8][k % 3]
    EH, EW, nx, ny, T34, xs, ys = _work_lists(H, W, Cm, scale)
    ycov = np.zeros((EH + 64, EW + 64), bool)          # [ey][ex]; lists may run past the grid edge
    for row0, strip, ng in ys:                          # y pass: rows walk y, strips walk x
        assert row0 % 8 == 0 and 1 <= ng <= 8
        ycov[row0:row0 + 8 * ng, strip * 64:strip * 64 + 64] = True
    xcov = np.zeros((EH + 64, EW + 64), bool)
    for row0, strip, ng in xs:                          # x pass: rows walk x, strips walk y
        assert row0 % 8 == 0 and 1 <= ng <= 8
        xcov[strip * 64:strip * 64 + 64, row0:row0 + 8 * ng] = True
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    x1 = T34[0] * xx + T34[1] * yy + T34[2]
    y1 = T34[3] * xx + T34[4] * yy + T34[5]
    ix, iy = np.floor(x1).astype(np.int64), np.floor(y1).astype(np.int64)
    touched = np.zeros((EH + 64, EW + 64), bool)
    for dx in (0, 1):
        for dy in (0, 1):
            cx, cy = ix + dx, iy + dy
            ok = (cx >= 0) & (cx < EW) & (cy >= 0) & (cy < EH)
            touched[cy[ok], cx[ok]] = True
    assert touched.sum() >= 0.9 * H * W                 # the read-back really lands inside the eigen-frame grid
    assert not (touched & ~ycov).any(), "a cell the read-back touches is not in the y list"
    # x-pass outputs the y pass needs: every cell within ny rows of a touched cell, same column
    need_x = np.zeros_like(touched)
    cols_any = np.flatnonzero(touched.any(axis=0))
    for ex in cols_any:
        rows = np.flatnonzero(touched[:, ex])
        lo, hi = max(0, rows.min() - ny), min(EH - 1, rows.max() + ny)
        # the footprint of a convex rectangle in one column is an interval, so is its dilation
        need_x[lo:hi + 1, ex] = True
    assert not (need_x[:EH, :EW] & ~xcov[:EH, :EW]).any(), "an x-pass output the y pass reads is not in the x list"
    if min(H, W) >= 256:                                # the point of the lists: the work stays close to the image's
        assert ycov[:EH, :EW].sum() <= 1.5 * H * W and xcov[:EH, :EW].sum() <= 1.8 * H * W


def _walks(H, W, Cm, scale=1.0):
    import ctypes
    from partapp_b200 import ExpParam, capi
    from partapp_b200.objectdetect import PartConf, make_config
    lib = capi.load_library()
    cfg = make_config(ExpParam(num_rotation_steps=8), PartConf([True], [False], [True]), H, W)
    dims = (ctypes.c_int * 7)()
    cap, mcap = 4096, 1 << 16
    wl = (ctypes.c_int * (4 * cap))()
    ml = (ctypes.c_ubyte * mcap)()
    nm = ctypes.c_int(0)
    Cc = (ctypes.c_double * 4)(*np.asarray(Cm, np.float64).ravel())
    rc = lib.ps_plan_walks(ctypes.byref(cfg), Cc, scale, dims, wl, cap, ml, mcap, ctypes.byref(nm))
    assert rc == 0
    EH, EW, nx, ny, halo, lag, nw = [int(v) for v in dims]
    assert nw <= cap and nm.value <= mcap
    return EH, EW, nx, ny, halo, lag, np.array(wl[:4 * nw], np.int64).reshape(-1, 4), np.array(ml[:nm.value], np.uint8)


@pytest.mark.parametrize("k", range(10))
def test_fused_gaussian_walks_cover_every_cell_the_read_back_touches(k):
    """The walks of the fused x+y Gaussian kernel (k_gauss_xy) against the same brute-force statement: every
    eigen-frame cell the bilinear read-back touches lies in its strip's y interval, every x-filtered value within the y
    reach of such a cell (same column) lies in an x block whose mask keeps its 8-column group, the ring of 64 (K + 1)
    rows holds a whole y window, and each strip appears once, longest first."""
    rng = np.random.default_rng(900 + k)
    H, W = [(150, 230), (260, 140), (64, 64), (600, 400), (33, 500), (129, 191), (400, 37), (300, 300), (90, 70),
            (512, 256)][k]
    th = rng.uniform(0, np.pi) if k else 0.6
    s1, s2 = rng.uniform(1.0, 12.0, 2)
    Rm = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    Cm = Rm @ np.diag([s1 * s1, s2 * s2]) @ Rm.T
    Cm[1, 0] = Cm[0, 1]
    scale = [1.0, 1.2, 0.8][k % 3]
    EH, EW, nx, ny, halo, lag, walks, masks = _walks(H, W, Cm, scale)
    _, _, _, _, T34, _, _ = _work_lists(H, W, Cm, scale)
    assert halo % 8 == 0 and ny <= halo < ny + 8 and 64 * lag >= ny + halo and 64 * (lag - 1) < ny + halo
    assert 64 * (lag + 1) >= 64 + 2 * ny                 # a y window of 64 + 2 ny rows fits the ring
    assert len(set(walks[:, 0].tolist())) == len(walks) and (np.diff(walks[:, 2]) <= 0).all()
    pad = 64 * (lag + 2)
    ycov = np.zeros((EH + 2 * pad, EW + 64), bool)       # [ey + pad][ex]
    xcov = np.zeros_like(ycov)
    for strip, row0, ng, moff in walks:
        assert row0 % 8 == 0 and ng >= 1
        ycov[pad + row0:pad + row0 + 8 * ng, strip * 64:strip * 64 + 64] = True
        nxb = (ng + 7) // 8 + lag
        for j in range(nxb):
            r0 = row0 - halo + 64 * j
            for g in range(8):
                if masks[moff + j] >> g & 1:
                    xcov[pad + r0:pad + r0 + 64, strip * 64 + 8 * g:strip * 64 + 8 * g + 8] = True
        # the last y block's window ends inside the last x block
        assert row0 + 8 * ng - 1 + ny <= row0 - halo + 64 * nxb - 1
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    x1 = T34[0] * xx + T34[1] * yy + T34[2]
    y1 = T34[3] * xx + T34[4] * yy + T34[5]
    ix, iy = np.floor(x1).astype(np.int64), np.floor(y1).astype(np.int64)
    touched = np.zeros_like(ycov)
    for dx in (0, 1):
        for dy in (0, 1):
            cx, cy = ix + dx, iy + dy
            ok = (cx >= 0) & (cx < EW) & (cy >= 0) & (cy < EH)
            touched[cy[ok] + pad, cx[ok]] = True
    assert not (touched & ~ycov).any(), "a cell the read-back touches is not in a walk"
    need_x = np.zeros_like(touched)
    for ex in np.flatnonzero(touched.any(axis=0)):
        rows = np.flatnonzero(touched[:, ex]) - pad
        lo, hi = max(0, rows.min() - ny), min(EH - 1, rows.max() + ny)   # rows outside the grid are zeros by construction
        need_x[pad + lo:pad + hi + 1, ex] = True
    assert not (need_x & ~xcov).any(), "an x-filtered value the y pass reads is masked out"
    if min(H, W) >= 256:
        assert ycov[pad:pad + EH, :EW].sum() <= 1.5 * H * W and xcov[pad:pad + EH, :EW].sum() <= 1.9 * H * W


def test_evaluator_unary_and_roi_hypothesis_sources(tmp_path):
    """EVAL_TYPE_UNARIES (parteval.cpp:326-383): the estimate of a part is the first maximum of its detector score grids
    in (rotation, x, y) scan order after the TM_DIRECT scatter of loadScoreGrid; eval_segments_roi (:1779-1889): one
    pose_est file per region of interest, a part is correct if it matches any annotated person."""
    import oracle
    from partapp_b200 import ExpParam, synth
    from partapp_b200 import parteval as pe
    ep = ExpParam(num_rotation_steps=8)
    H, W, P = 40, 36, 2
    cells, Tig = synth.compact_scores(ep, H, W, P, 3, rotated=True)
    for p in range(P):
        want = oracle.load_score_grid(cells[p, 0], Tig, H, W)                      # the oracle's loadScoreGrid
        got = pe.load_score_grid_direct(cells[p, 0], Tig, H, W)
        assert np.array_equal(got, want)
        # a tie: the same maximum at two places -> the (rotation, x, y) scan order decides, not the flat [r][y][x] order
        g = got.copy()
        g[:] = 0
        g[2, 30, 5] = g[2, 4, 9] = 7.0
        row = pe.unary_best_hyp(g, 0, 1.0, (-180.0, 180.0, 8))
        assert (row[2], row[4], row[5], row[6]) == (2, 5, 30, 7.0)                 # x = 5 comes before x = 9
        assert abs(row[3] - (-180 + 45 * 2.5)) < 1e-6
    # ROI evaluation: two people annotated, two regions of interest, each estimate matches one of them
    pd = [pe.PartDef(1, [1, 2], [1], [2], -90.0, 0, 0, 0, 0)]
    win = [pe.PartParam(20, 40, 10, 20)]
    person = lambda x, y: pe.AnnoRect(points={1: (x, y), 2: (x, y + 40)})
    annos = [pe.Annotation("a.png", [person(100, 100), person(200, 120)])]
    confs = {(0, 0): np.array([[0, 1, 0, 0, 100, 120, 1]], np.float32),   # upright stick centred on person 1
             (0, 1): np.array([[0, 1, 0, 0, 201, 141, 1]], np.float32),   # on person 2
             }
    r = pe.eval_segments_roi(annos, [2], pd, win, lambda i, k: confs[(i, k)], 0, 0)
    assert (r.seg_correct, r.seg_total) == (2, 2)
    confs[(0, 1)] = np.array([[0, 1, 0, 0, 300, 141, 1]], np.float32)    # far from both
    r = pe.eval_segments_roi(annos, [2], pd, win, lambda i, k: confs[(i, k)], 0, 0)
    assert (r.seg_correct, r.seg_total) == (1, 2)


def test_eval_segments_experiment_unary_source(tmp_path):
    """`eval_type="unaries"` on an experiment directory in the reference's formats: the hypotheses are the maxima of the
    score-grid files, and the annotated stick of each part is placed on that maximum, so every part is correct; moving the
    annotation of one part away makes exactly that part wrong."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import make_experiment
    import oracle
    from partapp_b200 import parteval as pe
    P, R, H, W = 6, 8, 48, 40
    info = make_experiment.make(str(tmp_path), num_images=2, P=P, R=R, H=H, W=W)
    win = "".join("part { window_size_x: 8 window_size_y: 16 pos_offset_x: 4 pos_offset_y: 8 }\n" for _ in range(P))
    with open(os.path.join(info["base"], "class", "window_param.txt"), "a") as f:
        f.write(win)
    conf = "".join("part {\n part_id: %d\n part_pos: %d\n part_pos: %d\n part_x_axis_from: %d\n part_x_axis_to: %d\n"
                   " part_x_axis_offset: -90\n}\n" % (p + 1, 2 * p + 1, 2 * p + 2, 2 * p + 1, 2 * p + 2) for p in range(P))
    (tmp_path / "part_conf_eval.txt").write_text(conf)
    with open(info["expopt"], "a") as f:
        f.write('part_conf_eval: "part_conf_eval.txt"\n')

    def write_al(shift_part=None):
        out = "<annotationlist>\n"
        for i in range(2):
            pts = ""
            for p in range(P):
                g = oracle.load_score_grid(info["cells"][i][p, 0], info["Tig"], H, W)
                row = pe.unary_best_hyp(g, 0, 1.0, (-180.0, 180.0, R))
                b = pe.bbox_from_hyp(row, pe.PartParam(8, 16, 4, 8))
                e1, e2, _ = pe.get_bbox_endpoints(b)
                if shift_part == (i, p):
                    e1, e2 = e1 + 30, e2 + 30
                for k, e in enumerate((e1, e2)):
                    pts += "<point><id>%d</id><x>%d</x><y>%d</y></point>" % (2 * p + 1 + k, round(e[0]), round(e[1]))
            out += ("<annotation><image><name>images/im%04d.png</name></image><annorect><x1>0</x1><y1>0</y1><x2>9</x2><y2>9</y2>"
                    "<annopoints>%s</annopoints></annorect></annotation>\n" % (i, pts))
        (tmp_path / "test.al").write_text(out + "</annotationlist>\n")

    write_al()
    r = pe.eval_segments_experiment(info["expopt"], eval_type="unaries", save_endpoints=False)
    assert (r.seg_correct, r.seg_total) == (2 * P, 2 * P)
    write_al(shift_part=(1, 3))
    r = pe.eval_segments_experiment(info["expopt"], eval_type="unaries", save_endpoints=False)
    assert (r.seg_correct, r.seg_total) == (2 * P - 1, 2 * P) and r.per_part_correct[3] == 1


def test_eval_segments_experiment_disc_ps_source(tmp_path):
    """EVAL_TYPE_DISC_PS (parteval.cpp:384-493): per part the sample with the first strictly largest posterior -- among
    all samples, or among those of one subject (vect_didx) -- becomes the PartHyp that is evaluated; files and
    directories as libDiscPS leaves them."""
    import scipy.io
    from partapp_b200 import parteval as pe
    (tmp_path / "part_conf.txt").write_text(_PART_CONF)
    (tmp_path / "test.al").write_text(_ANNOLIST)
    sub = tmp_path / "logs" / "exp-dps"
    (sub / "class").mkdir(parents=True)
    (sub / "class" / "window_param.txt").write_text("train_object_height: 200\n" + _WINDOW_PARAM)
    sdir = tmp_path / "dai" / "part_marginals_samples"
    post_dir = sub / "part_marginals_samples_post"
    post_dir.mkdir(parents=True)
    (tmp_path / "exp-dps.txt").write_text('test_dataset: "test.al"\nlog_dir: "./logs"\npart_conf: "part_conf.txt"\n'
                                          'num_rotation_steps: 24\ndai_samples_dir: "./dai/part_marginals_samples"\n')
    rot = lambda r: -180.0 + 360.0 / 24 * (0.5 + r)
    # image 0, part 0: three samples, the second and third tie on the posterior -> the second (first maximum) wins;
    # rotation index 11 is 0 +- 7.5 degrees: the upright stick at (100, 100) matches the annotation
    samples = {(0, 0): ([0, 0, 0], [3, 11, 20], [10, 100, 100], [10, 100, 30], [0.1, 0.7, 0.7], [0, 1, 0]),
               (0, 1): ([0, 0], [5, 17], [120, 120], [100, 100], [0.9, 0.2], [0, 1]),
               (1, 0): ([0], [11], [100], [180], [0.5], [0]),
               (1, 1): ([0], [11], [0], [0], [0.5], [0])}
    for (i, p), (si, ri, iy, ix, post, didx) in samples.items():
        d = sdir / ("samples_imgidx%04d" % i)
        d.mkdir(parents=True, exist_ok=True)
        scipy.io.savemat(str(d / ("samples_pidx%d.mat" % p)),
                         {"vect_scale_idx": np.array(si, np.float64), "vect_rot_idx": np.array(ri, np.float64),
                          "vect_iy": np.array(iy, np.float64), "vect_ix": np.array(ix, np.float64),
                          "vect_didx": np.array(didx, np.float64)})
    for i in range(2):
        scipy.io.savemat(str(post_dir / ("samples_imgidx%04d_post.mat" % i)),
                         {"samples_post_part%d" % p: np.array(samples[(i, p)][4], np.float64).reshape(-1, 1) for p in range(2)})
    row = pe.disc_ps_best_hyp(*samples[(0, 0)][:5], (1.0, 1.0, 1), (-180.0, 180.0, 24))
    assert row.tolist() == [0, 1, 11, np.float32(rot(11)), 100, 100, np.float32(0.7)]
    # the same through the experiment entry, against eval_segments fed with hand-built PartHyp rows
    annos = pe.load_annolist(str(tmp_path / "test.al"))
    conf = pe.load_part_conf(str(tmp_path / "part_conf.txt"))
    win = pe.load_window_param(str(sub / "class" / "window_param.txt"))
    hand = {0: np.array([[0, 1, 11, rot(11), 100, 100, 0.7], [0, 1, 5, rot(5), 100, 120, 0.9]], np.float32),
            1: np.array([[0, 1, 11, rot(11), 180, 100, 0.5], [0, 1, 11, rot(11), 0, 0, 0.5]], np.float32)}
    want = pe.eval_segments(annos, conf, win, lambda i: hand[i], 0, 1, 1.0, rot_range=(-180.0, 180.0, 24))
    got = pe.eval_segments_experiment(str(tmp_path / "exp-dps.txt"), eval_type="disc_ps")
    assert (got.seg_correct, got.seg_total, got.per_part_correct) == (want.seg_correct, want.seg_total, want.per_part_correct)
    assert got.seg_correct >= 1                                        # the upright stick of image 0 is found
    assert (post_dir / "seg_endpoints" / "endpoints_0000.mat").exists()
    # one subject only (vect_didx): subject 1 of part 1 is its second sample; subject 0 of part 0 are samples 0 and 2
    row1 = pe.disc_ps_best_hyp(*samples[(0, 1)][:5], (1.0, 1.0, 1), (-180.0, 180.0, 24), samples[(0, 1)][5], 1)
    assert row1[2] == 17 and row1[6] == np.float32(0.2)
    row0 = pe.disc_ps_best_hyp(*samples[(0, 0)][:5], (1.0, 1.0, 1), (-180.0, 180.0, 24), samples[(0, 0)][5], 0)
    assert row0[2] == 20 and row0[6] == np.float32(0.7) and (row0[4], row0[5]) == (30, 100)
    hand0 = {0: np.array([[0, 1, 20, rot(20), 30, 100, 0.7], [0, 1, 5, rot(5), 100, 120, 0.9]], np.float32)}
    want0 = pe.eval_segments(annos, conf, win, lambda i: hand0[i], 0, 0, 1.0, eval_didx=0, rot_range=(-180.0, 180.0, 24))
    got0 = pe.eval_segments_experiment(str(tmp_path / "exp-dps.txt"), first=0, numimgs=1, eval_type="disc_ps", eval_didx=0,
                                       save_endpoints=False)
    assert (got0.seg_correct, got0.seg_total, got0.per_part_correct) == (want0.seg_correct, want0.seg_total, want0.per_part_correct)
    assert got0.per_part_correct[0] == 0                               # the sample at x = 30 misses the stick at x = 100
