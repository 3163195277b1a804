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
