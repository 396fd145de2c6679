"""Host-side cost of the ingest path: time to ENQUEUE one 300-frame step (no synchronisation) vs the time until the GPU
has finished it.  python tools/host_enqueue.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, torch, bench
from pyslam_b200 import B200TsdfVolume
from pyslam_b200.sharding import FrameIngest
cfg, depth, color, Tcw = bench.load_frames("C2", 300, 0, 1)
vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=1 << 19)
vol.set_group_size(32)
D = torch.from_numpy(depth).pin_memory(); Cc = torch.from_numpy(color).pin_memory()
ing = FrameIngest(vol, chunk_frames=64)
for _ in range(3):
    ing.integrate_batch(D, Cc, cfg.K, Tcw); ing.synchronize()
enq, tot = [], []
for _ in range(5):
    t0 = time.perf_counter(); ing.integrate_batch(D, Cc, cfg.K, Tcw); t1 = time.perf_counter(); ing.synchronize(); t2 = time.perf_counter()
    enq.append(t1 - t0); tot.append(t2 - t0)
print("host enqueue ms per 300-frame step", [round(1e3 * x, 2) for x in enq], "total ms", [round(1e3 * x, 2) for x in tot])
# device-resident frames: pure library call cost (no H2D)
d = torch.from_numpy(depth).cuda(); c = torch.from_numpy(color).cuda()
for _ in range(2):
    vol.integrate_batch(d, c, cfg.K, Tcw); vol.synchronize()
e2 = []
for _ in range(5):
    t0 = time.perf_counter(); vol.integrate_batch(d, c, cfg.K, Tcw); t1 = time.perf_counter(); vol.synchronize(); t2 = time.perf_counter()
    e2.append((round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t0), 2)))
print("b2v_integrate_batch(300 device frames): (enqueue ms, total ms)", e2)
