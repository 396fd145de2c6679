import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from pyslam_b200 import B200TsdfVolume, synthetic as S
import bench
cfg, depth, color, Tcw = bench.load_frames("C2", 64, 0, 1)
vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=1<<19)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
d = torch.from_numpy(depth).cuda(); c = torch.from_numpy(color).cuda()
for _ in range(3): vol.integrate_batch(d, c, cfg.K, Tcw, stream=st.cuda_stream)
torch.cuda.synchronize()
vol.profile_enable(True)
vol.integrate_batch(d, c, cfg.K, Tcw, stream=st.cuda_stream)
torch.cuda.synchronize()
print(vol.profile_read())
