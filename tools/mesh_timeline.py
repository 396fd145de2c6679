"""Populate the C2 map (300 frames, or --frames N) and extract the mesh / point cloud a few times: the command the
ncu launch lists of the mesh_* kernels are captured from.   python tools/mesh_timeline.py [--frames N] [--config C2]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pyslam_b200 import B200TsdfVolume  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--config", default="C2")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
cfg, depth, color, Tcw = bench.load_frames(a.config, a.frames, 0, 1)
vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=1 << 19)
vol.integrate_batch(depth, color, cfg.K, Tcw)
vol.synchronize()
import ctypes as C
nv, nt = C.c_int64(0), C.c_int64(0)
for r in range(a.reps):
    t0 = time.perf_counter()
    vol._L.b2v_extract_mesh(vol._h, C.byref(nv), C.byref(nt))
    print(f"extract_mesh {1e3 * (time.perf_counter() - t0):.3f} ms: {vol.num_blocks()} blocks, {nv.value} vertices, {nt.value} triangles")
print("mesh stats", vol.last_mesh_stats())
t0 = time.perf_counter()
vol._L.b2v_extract_points(vol._h, C.byref(nv))
print(f"extract_points {1e3 * (time.perf_counter() - t0):.3f} ms: {nv.value} points")
print("point stats", vol.last_mesh_stats())
