"""End-to-end drop-in check: the C++ host CLI (`psinfer_partapp --expopt X --find_obj`) reads an experiment directory
in the reference's on-disk formats and writes pose_est / part_post / log_part_posterior_final .mat files and the
HypothesisList .pbuf; everything is compared with the CPU oracle run on the same inputs."""
import os
import subprocess

import numpy as np
import pytest
import scipy.io

import oracle
from partapp_b200 import synth
from tests.make_experiment import make
from tests.test_host_formats import _parse_hyps

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "partapp_b200", "psinfer_partapp")


@pytest.fixture(scope="module")
def cli(pslib):
    subprocess.run(["make", "-C", os.path.join(ROOT, "partapp_b200", "csrc", "host")], check=True, capture_output=True)
    return CLI


def _oracle_image(info, i):
    ep, P, H, W = info["ep"], info["P"], info["H"], info["W"]
    un = np.stack([np.stack([oracle.prepare_unary(oracle.load_score_grid(info["cells"][i][p, s], info["Tig"], H, W))
                             for s in range(info["S"])]) for p in range(P)])
    return oracle.infer(ep, synth.part_conf(P), info["joints"], np.ascontiguousarray(un), sparse=True, want_hyps=True)


def test_find_obj_cli_matches_oracle(cli, tmp_path):
    info = make(str(tmp_path / "exp"), num_images=2,
                extra_expopt="save_part_marginals_local_max: true\nsave_part_marginals: true\n")
    r = subprocess.run([cli, "--expopt", info["expopt"], "--find_obj"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pm = os.path.join(info["base"], "part_marginals")
    for i in range(2):
        want = _oracle_image(info, i)
        best = scipy.io.loadmat(os.path.join(pm, "pose_est_imgidx%04d.mat" % i))["best_conf"]
        assert best.shape == (info["P"], 7) and best.dtype == np.float32
        assert np.array_equal(best[:, :6], want["best_conf"][:, :6])
        np.testing.assert_allclose(best[:, 6], want["best_conf"][:, 6], rtol=1e-6)
        post = scipy.io.loadmat(os.path.join(pm, "part_post_imgidx%04d.mat" % i))
        for p in range(info["P"]):
            got, ref = post["part%d" % p], want["part_hyps"][p]
            assert got.shape == ref.shape and np.array_equal(got[0, :6], ref[0, :6])
        g = scipy.io.loadmat(os.path.join(pm, "log_part_posterior_final_imgidx%d_scaleidx0_o0_pidx1.mat" % i))["log_prob_grid"]
        ref = want["marginals"][0, 1]
        assert g.shape == ref.shape and np.array_equal(g == -1e6, ref == -1e6)
        np.testing.assert_allclose(g, ref, rtol=1e-4)
        hyps = _parse_hyps(open(os.path.join(info["base"], "object_hyp", "object_hyp_imgidx%d_o0_spmnone.pbuf" % i), "rb").read())
        lm = oracle.find_local_max(want["root_post"], 1000)
        assert len(hyps) == len(lm)
        keep = lm[:, 3] > -5e5
        ref_xy = sorted((int(x + info["bbox_offset"][0]), int(y + info["bbox_offset"][1])) for _, x, y, _ in lm[keep])
        got_xy = sorted((int(h[1]), int(h[2])) for h in hyps if h[4] > -5e5)
        assert got_xy == ref_xy                                       # bbox offset applied, truncated (findrot.cpp:1044-1045)
        assert all(abs(h[3] - 1.0) < 1e-6 and h.get(5, 0) == 0 for h in hyps)


def test_find_obj_cli_first_numimgs_and_errors(cli, tmp_path):
    info = make(str(tmp_path / "exp2"), num_images=3)
    r = subprocess.run([cli, "--expopt", info["expopt"], "--find_obj", "--first", "1", "--numimgs", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pm = os.path.join(info["base"], "part_marginals")
    assert sorted(os.listdir(pm)) == ["pose_est_imgidx0001.mat"]
    r = subprocess.run([cli, "--expopt", info["expopt"], "--train_class"], capture_output=True, text=True)
    assert r.returncode == 2 and "unsupported option" in r.stderr
    os.remove(os.path.join(info["base"], "spatial", "joint_2_1.mat"))
    r = subprocess.run([cli, "--expopt", info["expopt"], "--find_obj"], capture_output=True, text=True)
    assert r.returncode == 1 and "joint_2_1.mat" in r.stderr


def _conditioned_unaries(info, i):
    """findrot.cpp:849-949 with the oracle's pieces, in the reference's order."""
    import ctypes
    ep, P, R, H, W = info["ep"], info["P"], info["R"], info["H"], info["W"]
    c = info["cond"][i]
    L = oracle.lib()
    fp = ctypes.POINTER(ctypes.c_float)
    un = np.stack([oracle.prepare_unary(oracle.load_score_grid(info["cells"][i][p, 0], info["Tig"], H, W)) for p in range(P)])
    un = np.ascontiguousarray(un[:, None])
    root = info["root_idx"]
    with np.errstate(divide="ignore"):
        logt = np.where(c["dpm_torso"] == 0, np.float32(-1e6), np.log(c["dpm_torso"].astype(np.float64)).astype(np.float32))
    logt = np.ascontiguousarray(logt[None], np.float32)
    L.orc_add_dpm_score(un[root, 0].ctypes.data_as(fp), R, H, W, logt.ctypes.data_as(fp), 1, ctypes.c_float(0.5))
    if c["head"] < P:
        L.orc_add_load_dpm_score(un[c["head"], 0].ctypes.data_as(fp), R, H, W, c["dpm_head"].ctypes.data_as(fp), 1, ctypes.c_float(0.4))
    else:   # pidx_only outside the part list: the reference then adds the head grid to EVERY part (icps.cpp:465)
        for p in range(P):
            L.orc_add_load_dpm_score(un[p, 0].ctypes.data_as(fp), R, H, W, c["dpm_head"].ctypes.data_as(fp), 1, ctypes.c_float(0.4))
    for p in range(P):
        g = np.ascontiguousarray(c["dpm_unary"][p])
        L.orc_add_load_dpm_score(un[p, 0].ctypes.data_as(fp), R, H, W, g.ctypes.data_as(fp), R, ctypes.c_float(0.3))
    t = np.zeros(R, np.float32)
    for p in range(P):
        L.orc_rot_score_table(ctypes.byref(oracle.exp_param(ep)), ctypes.c_double(c["rot"][p, 0]), ctypes.c_double(c["rot"][p, 1] ** 2),
                              t.ctypes.data_as(fp))
        L.orc_add_rot_table(un[p, 0].ctypes.data_as(fp), R, H, W, t.ctypes.data_as(fp), ctypes.c_float(0.8))
    t2 = np.zeros((H, W), np.float32)
    for p in range(P):
        if p == root:
            continue
        q = c["pos"][p]
        L.orc_pos_score_table(H, W, ctypes.c_double(q[0]), ctypes.c_double(q[1]), ctypes.c_double(q[2] ** 2), ctypes.c_double(q[3] ** 2),
                              ctypes.c_double(c["rootpos"][0]), ctypes.c_double(c["rootpos"][1]), t2.ctypes.data_as(fp))
        L.orc_add_pos_table(un[p, 0].ctypes.data_as(fp), R, H, W, t2.ctypes.data_as(fp), ctypes.c_float(0.6))
    pr = c["prior"]
    L.orc_torso_prior_table(H, W, ctypes.c_double(pr[0]), ctypes.c_double(pr[1]), ctypes.c_double(pr[2]), ctypes.c_double(pr[3]),
                            ctypes.c_float(0.7), t2.ctypes.data_as(fp))
    L.orc_add_pos_table_unweighted(un[root, 0].ctypes.data_as(fp), R, H, W, t2.ctypes.data_as(fp))
    return un


def test_find_obj_cli_conditioned_model_reads_predictor_outputs(cli, tmp_path):
    """The poselet-conditioned full model through the drop-in host: rotation / position parameters, the torso detection,
    the torso prior and three kinds of DPM score grids are read from the files the reference's MATLAB side writes
    (objectdetect_icps.cpp:193-606) and added in the reference's order (findrot.cpp:849-949)."""
    info = make(str(tmp_path / "expc"), num_images=2, P=6, R=8, H=44, W=40, conditioning=True,
                extra_expopt="save_part_marginals: true\n")
    r = subprocess.run([cli, "--expopt", info["expopt"], "--find_obj"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pm = os.path.join(info["base"], "part_marginals")
    for i in range(2):
        un = _conditioned_unaries(info, i)
        want = oracle.infer(info["ep"], synth.part_conf(info["P"]), info["joints"], un, sparse=True)
        best = scipy.io.loadmat(os.path.join(pm, "pose_est_imgidx%04d.mat" % i))["best_conf"]
        assert np.array_equal(best[:, :6], want["best_conf"][:, :6]), "image %d" % i
        # addLoadDPMScore goes through logf (glibc < 1 ulp vs the device's correctly rounded one): scores to 1e-5
        np.testing.assert_allclose(best[:, 6], want["best_conf"][:, 6], rtol=1e-5)
        g = scipy.io.loadmat(os.path.join(pm, "log_part_posterior_final_imgidx%d_scaleidx0_o0_pidx0.mat" % i))["log_prob_grid"]
        np.testing.assert_allclose(g, want["marginals"][0, 0], rtol=1e-4)
    # use_gt_torso (icps.cpp:292-300): the position tables are centred on the annotated root part instead of the torso
    # detection -- part_pos of the first annotated rectangle, two points -> their mean, truncated to int
    root = info["root_idx"]
    exp_dir = str(tmp_path / "expc")
    conf = open(os.path.join(exp_dir, "part_conf.txt")).read()
    marker = "part_id: %d\n  part_pos: %d\n" % (root + 1, root)
    assert conf.count(marker) == 1
    open(os.path.join(exp_dir, "part_conf.txt"), "w").write(
        conf.replace(marker, "part_id: %d\n  part_pos: 3\n  part_pos: 7\n" % (root + 1)))
    gt = [((11, 20), (18, 25)), ((30, 9), (21, 14))]
    with open(os.path.join(exp_dir, "test.al"), "w") as f:
        f.write("<annotationlist>\n")
        for i in range(2):
            pts = "".join("<point><id>%d</id><x>%d</x><y>%d</y></point>" % (k, x, y) for k, (x, y) in zip((3, 7), gt[i]))
            f.write("<annotation><image><name>images/im%04d.png</name></image><annorect><annopoints>%s</annopoints></annorect>"
                    "</annotation>\n" % (i, pts))
        f.write("</annotationlist>\n")
    with open(info["expopt"], "a") as f:
        f.write("use_gt_torso: true\n")
    r = subprocess.run([cli, "--expopt", info["expopt"], "--find_obj"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for i in range(2):
        info["cond"][i]["rootpos"] = (int((gt[i][0][0] + gt[i][1][0]) * 0.5), int((gt[i][0][1] + gt[i][1][1]) * 0.5))
        want = oracle.infer(info["ep"], synth.part_conf(info["P"]), info["joints"], _conditioned_unaries(info, i), sparse=True)
        best = scipy.io.loadmat(os.path.join(pm, "pose_est_imgidx%04d.mat" % i))["best_conf"]
        assert np.array_equal(best[:, :6], want["best_conf"][:, :6]), "use_gt_torso image %d" % i
        np.testing.assert_allclose(best[:, 6], want["best_conf"][:, 6], rtol=1e-5)
    # a missing predictor file is an error that names the file, not a silent skip
    os.remove(os.path.join(str(tmp_path / "expc"), "pred_data_test", "testlist_params_pos_imgidx_1.mat"))
    r = subprocess.run([cli, "--expopt", info["expopt"], "--find_obj", "--first", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "testlist_params_pos_imgidx_1.mat" in r.stderr


def test_find_obj_cli_multi_worker_outputs_are_byte_identical(cli, tmp_path):
    """findObjectDataset with several worker threads / contexts (and, where the box has them, several GPUs) writes the
    same bytes as the single-threaded loop of the reference (aux.cpp:344-402)."""
    import filecmp
    info_a = make(str(tmp_path / "a"), num_images=8, P=4, R=8, H=48, W=40)
    info_b = make(str(tmp_path / "b"), num_images=8, P=4, R=8, H=48, W=40)
    ra = subprocess.run([cli, "--expopt", info_a["expopt"], "--find_obj", "--gpus", "1", "--contexts", "1"], capture_output=True, text=True)
    assert ra.returncode == 0, ra.stderr
    import torch
    ng = min(2, torch.cuda.device_count())
    rb = subprocess.run([cli, "--expopt", info_b["expopt"], "--find_obj", "--gpus", str(ng), "--contexts", "3"], capture_output=True, text=True)
    assert rb.returncode == 0, rb.stderr
    for sub in ("part_marginals", "object_hyp"):
        da, db = os.path.join(info_a["base"], sub), os.path.join(info_b["base"], sub)
        names = sorted(os.listdir(da))
        assert names == sorted(os.listdir(db)) and len(names) == 8
        for n in names:
            assert filecmp.cmp(os.path.join(da, n), os.path.join(db, n), shallow=False), n
    # process-level shards as in the reference (--distribute, main.cpp:175-184): batch 1 of 4 = images 2..3
    info_c = make(str(tmp_path / "c"), num_images=8, P=4, R=8, H=48, W=40)
    rc = subprocess.run([cli, "--expopt", info_c["expopt"], "--find_obj", "--distribute", "--ncpu", "4", "--batch_num", "1"],
                        capture_output=True, text=True)
    assert rc.returncode == 0, rc.stderr
    assert sorted(os.listdir(os.path.join(info_c["base"], "part_marginals"))) == ["pose_est_imgidx0002.mat", "pose_est_imgidx0003.mat"]
