"""Writes a synthetic partapp experiment directory in the reference's own on-disk formats (the reference ships none:
code_test.zip is absent): expopt / part_conf protobuf text, an .al image list with real PNG files, joint_<c>_<p>.mat,
imgidx<i>-pidx<p>-o0-scoregrid.mat (cell_scoregrid + transform_Ti2/T2g), written with scipy.io.savemat -- an
implementation independent of the C++ reader under test."""
import os
import struct
import zlib

import numpy as np
import scipy.io

from partapp_b200 import synth
from partapp_b200.objectdetect import ExpParam


def write_png(path, width, height):
    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    raw = b"".join(b"\x00" + b"\x80" * width for _ in range(height))
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, 8, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def make(root, num_images=2, P=4, R=8, H=48, W=40, S=1, seed=5, extra_expopt="", conditioning=False, link_after=0):
    os.makedirs(root, exist_ok=True)
    ep = ExpParam(num_rotation_steps=R, num_scale_steps=S, roi_save_num_samples=20,
                  min_object_scale=1.0 if S == 1 else 0.9, max_object_scale=1.0 if S == 1 else 1.1)
    edges, root_idx = synth.tree(P)
    joints = synth.make_joints(P, seed=seed, max_offset=6, sigma_range=(1.5, 3))
    log_dir = os.path.join(root, "log_dir")
    sub = "exp-synth"
    base = os.path.join(log_dir, sub)
    for d in ("spatial", "test_scoregrid", "class"):
        os.makedirs(os.path.join(base, d), exist_ok=True)
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    # images + annotation list (.al XML, libAnnotation)
    names = []
    for i in range(num_images):
        nm = "images/im%04d.png" % i
        write_png(os.path.join(root, nm), W, H)
        names.append(nm)
    with open(os.path.join(root, "test.al"), "w") as f:
        f.write("<annotationlist>\n" + "".join(
            "<annotation><image><name>%s</name></image></annotation>\n" % n for n in names) + "</annotationlist>\n")
    # part_conf (PartConfig.proto text): ids are 1-based
    with open(os.path.join(root, "part_conf.txt"), "w") as f:
        for p in range(P):
            f.write("part {\n  part_id: %d\n  part_pos: %d\n  is_detect: true\n  is_root: %s\n}\n" %
                    (p + 1, p, "true" if p == root_idx else "false"))
        for (c, p) in edges:
            f.write('joint {\n  child_idx: %d\n  parent_idx: %d\n  type: "RotGaussian"\n}\n' % (c + 1, p + 1))
    with open(os.path.join(base, "class", "window_param.txt"), "w") as f:
        f.write("train_object_height: 200\nbbox_offset_x: 3.7\nbbox_offset_y: -2.2\n")
    if conditioning:
        extra_expopt += ('pred_unary_rot: true\npred_unary_rot_weight: 0.8\npred_unary_pos: true\npred_unary_pos_weight: 0.6\n'
                         'use_torso_pos_prior: true\ntorso_pos_prior_weight: 0.7\nuse_dpm_torso: true\ndpm_torso_weight: 0.5\n'
                         'use_dpm_unary: true\ndpm_unary_weight: 0.3\ndo_dpm_rot: true\nuse_dpm_head: true\ndpm_head_weight: 0.4\n'
                         'rootidx_det: 2\ntorso_det_test_dir: "./torso_det"\ntest_dpm_torso_dir: "./dpm_torso"\n'
                         'test_dpm_unary_dir: "./dpm_unary"\npred_data_test_dir: "./pred_data_test"\n')
    # expopt; relative paths are resolved against this file (partapp.cpp:112-139)
    with open(os.path.join(root, "exp-synth.txt"), "w") as f:
        f.write('# synthetic experiment\ntest_dataset: "test.al"\nlog_dir: "./log_dir"\npart_conf: "part_conf.txt"\n'
                "num_rotation_steps: %d\nmin_part_rotation: -180\nmax_part_rotation: 180\nnum_scale_steps: %d\n"
                "min_object_scale: %g\nmax_object_scale: %g\nroi_save_num_samples: 20\nforce_recompute_scores: false\n%s"
                % (R, S, ep.min_object_scale, ep.max_object_scale, extra_expopt))
    # joints (objectdetect_learnparam.cpp:61-90 save_joint layout), compressed like mat_open(.., "wz")
    for j in joints:
        scipy.io.savemat(os.path.join(base, "spatial", "joint_%d_%d.mat" % (j.child_idx + 1, j.parent_idx + 1)), {
            "type": float(j.type), "child_idx": float(j.child_idx + 1), "parent_idx": float(j.parent_idx + 1),
            "offset_c": np.asarray(j.offset_c, np.float64).reshape(2, 1),
            "offset_p": np.asarray(j.offset_p, np.float64).reshape(2, 1),
            "C": np.asarray(j.C, np.float64), "rot_mean": float(j.rot_mean), "rot_sigma": float(j.rot_sigma)},
            do_compression=True)
    # score grids (partapp.cpp:792-799 naming; cell_scoregrid{scale,rot}, transform_Ti2/T2g [S][R][3][3] single)
    cells_all, tig_all = [], None
    for i in range(num_images):
        if link_after and i >= link_after:   # big experiments: later images reuse the files of image i % link_after
            cells_all.append(cells_all[i % link_after])
            for p in range(P):
                os.symlink("imgidx%d-pidx%d-o0-scoregrid.mat" % (i % link_after, p),
                           os.path.join(base, "test_scoregrid", "imgidx%d-pidx%d-o0-scoregrid.mat" % (i, p)))
            continue
        cells, Tig = synth.compact_scores(ep, H, W, P, i)
        cells_all.append(cells)
        tig_all = Tig
        for p in range(P):
            cg = np.empty((S, R), dtype=object)
            for s in range(S):
                for r in range(R):
                    cg[s, r] = cells[p, s, r]
            Ti2 = np.zeros((S, R, 3, 3), np.float32)
            T2g = np.zeros((S, R, 3, 3), np.float32)
            for s in range(S):
                for r in range(R):
                    # Tig = Ti2 * T2g with integer entries -> the double product is exact
                    Ti2[s, r] = [[1, 0, Tig[r, 0, 2]], [0, 1, Tig[r, 1, 2]], [0, 0, 1]]
                    T2g[s, r] = [[Tig[r, 0, 0], 0, 0], [0, Tig[r, 1, 1], 0], [0, 0, 1]]
            scipy.io.savemat(os.path.join(base, "test_scoregrid", "imgidx%d-pidx%d-o0-scoregrid.mat" % (i, p)),
                             {"cell_scoregrid": cg, "transform_Ti2": Ti2, "transform_T2g": T2g}, do_compression=True)
    cond = write_conditioning(root, base, num_images, P, R, H, W, root_idx, seed) if conditioning else None
    return {"cond": cond, "ep": ep, "joints": joints, "P": P, "R": R, "S": S, "H": H, "W": W, "cells": cells_all, "Tig": tig_all,
            "expopt": os.path.join(root, "exp-synth.txt"), "base": base, "root_idx": root_idx,
            "bbox_offset": (3.7, -2.2)}


def write_conditioning(root, base, num_images, P, R, H, W, root_idx, seed):
    """The per-image predictor outputs the conditioned model reads (objectdetect_icps.cpp:193-226, :283-324, :326-363,
    :445-486, :550-581) plus class/torso_pos_prior.mat, written the way MATLAB writes them (double matrices, cell arrays).
    Returns what was written so a test can rebuild the expected unaries."""
    out = []
    for d in ("pred_data_test", "torso_det", "dpm_torso", "dpm_unary/head"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    for p in range(P):
        os.makedirs(os.path.join(root, "dpm_unary", "pidx_%04d" % p), exist_ok=True)
    prior = np.array([[2.0, -3.0, 150.0, 220.0]])
    scipy.io.savemat(os.path.join(base, "class", "torso_pos_prior.mat"), {"params": prior})
    head = 11 if P == 22 else (1 if P == 12 else 5)
    for i in range(num_images):
        rng = np.random.default_rng([seed, 77, i])
        rot = np.column_stack([rng.uniform(-1, 1, P), rng.uniform(0.3, 0.8, P)])            # mu, sigma
        pos = np.column_stack([rng.uniform(-8, 8, P), rng.uniform(-8, 8, P), rng.uniform(6, 12, P), rng.uniform(6, 12, P)])
        clus = rng.integers(1, 4, (P, 1)).astype(np.float64)
        pd = os.path.join(root, "pred_data_test")
        scipy.io.savemat(os.path.join(pd, "testlist_params_rot_imgidx_%d.mat" % i), {"rot_test": rot})
        scipy.io.savemat(os.path.join(pd, "testlist_pred_rot_imgidx_%d.mat" % i), {"clusidx_test": clus})
        scipy.io.savemat(os.path.join(pd, "testlist_params_pos_imgidx_%d.mat" % i), {"pos_test": pos})
        scipy.io.savemat(os.path.join(pd, "testlist_pred_pos_imgidx_%d.mat" % i), {"clusidx_test": clus})
        best = np.zeros((3, 7))
        best[2, 4:6] = [W * 0.5 + rng.uniform(-3, 3), H * 0.5 + rng.uniform(-3, 3)]           # row rootidx_det = 2: x, y
        scipy.io.savemat(os.path.join(root, "torso_det", "pose_est_imgidx%04d.mat" % i), {"best_conf": best})
        torso = rng.uniform(0.0, 1.0, (H, W)).astype(np.float32)
        torso[:2] = 0.0                                                                      # zeros -> LOG_ZERO
        scipy.io.savemat(os.path.join(root, "dpm_torso", "imgidx_%04d.mat" % (i + 1)), {"scoregrid": torso}, do_compression=True)
        unary = []
        for p in range(P):
            g = rng.uniform(-0.1, 1.0, (R, H, W)).astype(np.float32)
            cell = np.empty((R, 1), dtype=object)
            for r in range(R):
                cell[r, 0] = g[r]
            scipy.io.savemat(os.path.join(root, "dpm_unary", "pidx_%04d" % p, "imgidx_%04d.mat" % (i + 1)), {"scoregrid": cell},
                             do_compression=True)
            unary.append(g)
        hg = rng.uniform(-0.1, 1.0, (1, H, W)).astype(np.float32)
        cell = np.empty((1, 1), dtype=object)
        cell[0, 0] = hg[0]
        scipy.io.savemat(os.path.join(root, "dpm_unary", "head", "imgidx_%04d.mat" % (i + 1)), {"scoregrid": cell})
        out.append({"rot": rot, "pos": pos, "rootpos": (int(best[2, 4]), int(best[2, 5])), "dpm_torso": torso,
                    "dpm_unary": unary, "dpm_head": hg, "head": head, "prior": prior[0]})
    return out
