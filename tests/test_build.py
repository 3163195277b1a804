"""CPU checks of the build: the C-ABI library exists, loads without a GPU, exports every symbol include/psinfer.h
declares, fails loudly (no CPU fallback), and its Gaussian kernels keep both roundings of a tap (FFMA2 + FADD2)."""
import ctypes
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "psinfer.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(pslib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(pslib, s), "libpsinfer.so does not export %s" % s


def test_ctypes_prototypes_cover_header():
    from partapp_b200 import capi
    assert sorted(capi.PROTOTYPES) == declared_symbols()


def test_version_and_pure_host_entry_points(pslib):
    from partapp_b200 import capi
    assert b"sm_100a" in pslib.ps_version()
    cfg = capi.ps_config()
    cfg.num_rotation_steps, cfg.min_part_rotation, cfg.max_part_rotation = 24, -180.0, 180.0
    cfg.num_scale_steps, cfg.min_object_scale, cfg.max_object_scale = 5, 0.8, 1.2
    assert pslib.ps_rot_from_index(ctypes.byref(cfg), 0) == -172.5
    assert abs(pslib.ps_scale_from_index(ctypes.byref(cfg), 2) - 1.0) < 1e-7
    assert pslib.ps_index_from_rot(ctypes.byref(cfg), 0.5) == 12
    j = capi.ps_joint()
    j.type = capi.PS_JOINT_ROT_GAUSSIAN
    j.offset_c[0], j.offset_c[1], j.rot_mean = 3.0, 4.0, 0.25
    j.C[0], j.C[1], j.C[2], j.C[3] = 2.0, 0.5, 0.5, 1.0
    pslib.ps_flip_joint(ctypes.byref(j))
    assert (j.offset_c[0], j.offset_c[1], j.rot_mean) == (-3.0, 4.0, -0.25)
    assert (j.C[0], j.C[1], j.C[2], j.C[3]) == (2.0, -0.5, -0.5, 1.0)


def test_conditioning_tables_match_oracle(pslib, oracle_lib):
    import numpy as np
    from partapp_b200 import ExpParam, capi
    from partapp_b200.objectdetect import PartConf, make_config
    ep = ExpParam(num_rotation_steps=24)
    cfg = make_config(ep, PartConf([True], [False], [True]), 30, 20)
    fp = ctypes.POINTER(ctypes.c_float)
    a, b = np.zeros(24, np.float32), np.zeros(24, np.float32)
    pslib.ps_rot_score_table(ctypes.byref(cfg), 0.3, 0.2, a.ctypes.data_as(fp))
    oracle_lib.lib().orc_rot_score_table(ctypes.byref(oracle_lib.exp_param(ep)), 0.3, 0.2, b.ctypes.data_as(fp))
    assert np.array_equal(a, b)
    a, b = np.zeros((30, 20), np.float32), np.zeros((30, 20), np.float32)
    pslib.ps_pos_score_table(30, 20, 2.0, -3.0, 40.0, 60.0, 9.0, 14.0, a.ctypes.data_as(fp))
    oracle_lib.lib().orc_pos_score_table(30, 20, 2.0, -3.0, 40.0, 60.0, 9.0, 14.0, b.ctypes.data_as(fp))
    assert np.array_equal(a, b)
    pslib.ps_torso_prior_table(30, 20, 1.0, 2.0, 50.0, 80.0, 0.7, a.ctypes.data_as(fp))
    oracle_lib.lib().orc_torso_prior_table(30, 20, 1.0, 2.0, 50.0, 80.0, 0.7, b.ctypes.data_as(fp))
    assert np.array_equal(a, b)


def test_no_cpu_fallback_without_gpu(pslib):
    """Without a CUDA device ps_create must fail with PS_ERR_CUDA -- never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from partapp_b200 import ExpParam, PsContext, PsInferError, capi, synth
    with pytest.raises(PsInferError) as ei:
        PsContext(ExpParam(num_rotation_steps=8), synth.part_conf(2), 8, 8)
    assert ei.value.status == capi.PS_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "partapp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "ps_oracle" not in text and "from oracle" not in text, f


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_gaussian_kernels_keep_both_roundings(pslib):
    """Parity arithmetic in SASS: in the packed Gaussian kernels every FFMA2 (multiply with a run-time -0.0 addend) is
    paired with an FADD2; a lone FFMA2 would mean ptxas contracted multiply and add."""
    from partapp_b200 import capi
    out = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels = re.split(r"\n\s*Function : ", out)
    seen = fast = 0
    for k in kernels[1:]:
        name = k.split("\n", 1)[0]
        if not any(t in name for t in ("k_conv_cols2", "k_conv_rows2", "k_rotconv3", "k_conv_cols_tma", "k_rotconv4", "k_gauss_xy")):
            continue
        ffma2, fadd2 = len(re.findall(r"\bFFMA2\b", k)), len(re.findall(r"\bFADD2\b", k))
        assert not re.search(r"\bFFMA\b", k), "%s contains a scalar FFMA" % name
        if re.search(r"k_conv_cols_tma2ILi\d+ELb1E|k_rotconv4I.*ELb1EEE|k_gauss_xyILb1E", name):  # the ps_config.fast_math instantiations
            fast += 1
            assert ffma2 > 0 and fadd2 == 0, "%s: fast-math variant with %d FADD2" % (name, fadd2)
            continue
        seen += 1
        assert ffma2 > 0 and ffma2 == fadd2, "%s: %d FFMA2 vs %d FADD2" % (name, ffma2, fadd2)
    assert seen >= 6 and fast >= 2


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")
def test_kernel_footprints_leave_room_for_co_resident_blocks(pslib):
    """The shapes DESIGN.md section 5 measured its way to: the Gaussian kernel fits two blocks into half of an SM's
    register file (56 registers x 288 threads), the full-slice-group resampling kernels stay at 32 registers (eight
    blocks per SM, four beside two Gaussian blocks), the TMA wait sleeps on its barrier instead of polling it, and the
    Gaussian kernel really is fed by cp.async.bulk.tensor (UTMALDG)."""
    from partapp_b200 import capi
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", capi.LIB_PATH], capture_output=True, text=True,
                         check=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", res):
        regs[m.group(1)] = int(m.group(2))
    conv = [r for n, r in regs.items() if "k_conv_cols_tma2ILi8" in n]
    assert len(conv) == 2 and max(conv) <= 56, conv
    full = [r for n, r in regs.items() if "k_resample_bilinearILi8ELb1" in n or "k_warp_direct2ILi8ELb1" in n]
    assert len(full) == 2 and max(full) <= 32, full
    full = [r for n, r in regs.items() if "k_resample_bilinear_bILi8" in n or "k_warp_direct_bILi8" in n]   # level-batched
    assert len(full) == 2 and max(full) <= 32, full
    # the cfg-2 shapes (the fused-multiply-add build of the 7-tap rotation filter may take 48)
    assert all(r <= 40 for n, r in regs.items() if "k_epilogue3" in n or ("k_rotconv4ILi24E" in n and "ELb0EEE" in n))
    assert all(r <= 48 for n, r in regs.items() if "k_rotconv4ILi24E" in n)
    # fused x+y Gaussian (parity / fast_math x static / dynamic work order): 64 registers -- two 256-thread blocks of the
    # fixed order take exactly half of the 64 K registers and leave room for four 256-thread blocks of the 32-register
    # memory-bound kernels
    fused = [r for n, r in regs.items() if "k_gauss_xy" in n]
    assert len(fused) == 4 and max(fused) <= 64, fused
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    for k in re.split(r"\n\s*Function : ", sass)[1:]:
        if "k_conv_cols_tma2ILi8" in k.split("\n", 1)[0] or "k_gauss_xy" in k.split("\n", 1)[0]:
            assert "UTMALDG.3D" in k and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in k and "NANOSLEEP.SYNCS" in k
