"""Timing of the semantic row: one 640x480 frame worth of labelled points per call (C3-style), GPU vs the compiled
reference on this host.  Not a bench.py line; numbers go to DESIGN.md."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import oracle
from pyslam_b200 import VoxelBlockSemanticGrid, VoxelBlockSemanticProbabilisticGrid
from pyslam_b200 import synthetic as S

cfg = S.CONFIGS["C2"]
rng = np.random.default_rng(0)
frames = []
for i in range(6):
    d, c, T = S.render_frame(cfg, i)
    fx, fy, cx, cy = cfg.K
    valid = (d > 0) & (d < cfg.depth_trunc)
    z = d[valid].astype(np.float64)
    rows, cols = np.where(valid)
    pc = np.column_stack([(cols - cx) * z / fx, (rows - cy) * z / fy, z])
    Twc = S.inv_T(T)
    pw = np.ascontiguousarray((Twc[:3, :3] @ pc.T + Twc[:3, 3].reshape(3, 1)).T, np.float32)
    col = np.ascontiguousarray(c[valid] / 255.0, np.float32)
    cls = (1 + np.argmax(c[valid].astype(np.int32), axis=1)).astype(np.int32)
    ins = (cls * 10 + rows // 120).astype(np.int32)
    frames.append((pw, col, cls, ins, z.astype(np.float32)))
print("points/frame", len(frames[0][0]), "image", d.shape)
for name, gcls, kind in (("voting", VoxelBlockSemanticGrid, "voting"),
                         ("probabilistic", VoxelBlockSemanticProbabilisticGrid, "probabilistic")):
    g = gcls(0.015, 8, capacity_blocks=1 << 15)
    g.integrate(*frames[0])
    t0 = time.perf_counter()
    for f in frames[1:]:
        g.integrate(*f)
    tg = (time.perf_counter() - t0) / (len(frames) - 1)
    t0 = time.perf_counter()
    v = g.get_voxels(1, 0.0)
    tv = time.perf_counter() - t0
    line = f"{name}: gpu integrate {tg * 1e3:.2f} ms/frame (host arrays in, sync), get_voxels {tv * 1e3:.2f} ms ({len(v.points)} voxels)"
    if oracle.have_ref_semantic():
        r = oracle.RefSemanticGrid(0.015, kind)
        r.integrate(*frames[0])
        t0 = time.perf_counter()
        for f in frames[1:]:
            r.integrate(*f)
        tr = (time.perf_counter() - t0) / (len(frames) - 1)
        t0 = time.perf_counter()
        rv = r.get_voxels(1, 0.0)
        trv = time.perf_counter() - t0
        line += f" | reference cpu integrate {tr * 1e3:.2f} ms/frame, get_voxels {trv * 1e3:.2f} ms ({len(rv['points'])} voxels)"
    print(line)
