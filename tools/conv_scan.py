#!/usr/bin/env python
"""Efficiency of the Gaussian passes against the fp32-pipe ceiling as a function of the filter length.
Runs single messages (ps_message) with a rotated covariance diag(s1^2, s2^2) at cfg-2 size and reads the per-class
device times of the instrumented launches.  usage (on a GPU box): python tools/conv_scan.py"""
import math, os, sys
import numpy as np
sys.path.insert(0, '.')
from partapp_b200 import ExpParam, PartConf, PsContext

R, H, W = 24, 600, 400
ep = ExpParam(num_rotation_steps=R)
pc = PartConf([True, True], [False, False], [True, False])
ctx = PsContext(ep, pc, H, W, device=0)
rng = np.random.default_rng(0)
g = np.log(rng.random((R, H, W), dtype=np.float32) + 1e-3).astype(np.float32)
PEAK = 18.0e12
th = 0.5
rot = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
CASES = [(0.3, 0.31), (1, 1.05), (2, 2.1), (4, 4.2), (6, 6.3), (8, 8.4), (12, 12.5), (16, 16.5), (4, 16), (8, 16)]
if os.environ.get('CONV_SCAN_CASES'):
    CASES = [CASES[int(i)] for i in os.environ['CONV_SCAN_CASES'].split(',')]
for s1, s2 in CASES:
    Cm = rot @ np.diag([s1 * s1, s2 * s2]) @ rot.T
    ctx.message(g, (10, -5), (3, 7), Cm, 0.1, 0.4, 1.0, False)  # warm: plan, maps
    ctx.profile_enable(True)
    ctx.profile_read()
    for _ in range(5):
        ctx.message(g, (10, -5), (3, 7), Cm, 0.1, 0.4, 1.0, False)
    p = ctx.profile_read()
    ctx.profile_enable(False)
    nx, ny = 2 * int(math.floor(3 * s1 + 0.5)) + 1, 2 * int(math.floor(3 * s2 + 0.5)) + 1
    c, s = abs(math.cos(th)), abs(math.sin(th))
    EW, EH = math.ceil((W - 1) * c + (H - 1) * s), math.ceil((W - 1) * s + (H - 1) * c)
    line = f'sigma=({s1},{s2}) taps=({nx},{ny}) eigen={EH}x{EW}'
    for cls, taps in (('conv_rows', nx), ('conv_cols', ny)):
        if cls in p:
            ms = p[cls][0] / p[cls][1]
            full = R * EH * EW * taps          # every eigen-frame cell
            live = R * H * W * taps            # cells inside the rotated image footprint (what the tile lists keep, roughly)
            line += f' | {cls} {ms * 1e3:6.1f} us  all-cells {full / ms / 1e9 / PEAK * 1e12:5.2f}  footprint {live / ms / 1e9 / PEAK * 1e12:5.2f}'
    print(line)
    others = {k: round(v[0] / v[1] * 1e3, 1) for k, v in p.items() if k not in ('conv_rows', 'conv_cols')}
    print('    other kernels (us):', others)
