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
