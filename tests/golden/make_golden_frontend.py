"""Golden vectors of the per-frame prep rows, produced by the REFERENCE'S OWN Python code
imported from /root/reference (build container only):

    python tests/golden/make_golden_frontend.py

frontend_T0.npz
  filter_shadow_points(depth)                         pyslam/utilities/depth.py:103-146   (unmodified function)
  depth2pointcloud(depth_filtered, color, ...)        pyslam/utilities/depth.py:45-85     (unmodified function)
  world transform + float32 casts                     pyslam/dense/volumetric_integrator_voxel_grid.py:262-281
                                                      (restated here: 3 lines of numpy)
  and the dump of the compiled, unmodified reference VoxelBlockGrid fed those points (oracle/_ref).
"""

import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pyslam_b200 import synthetic as S  # noqa: E402
from tests._util import GOLDEN, sort_dump  # noqa: E402

REF_DEPTH_PY = "/root/reference/pyslam/utilities/depth.py"


def load_reference_depth_module():
    spec = importlib.util.spec_from_file_location("pyslam_ref_depth", REF_DEPTH_PY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_depth_module()
    cfg = S.CONFIGS["T0"]
    n = 3
    frames = [S.render_frame(cfg, i) for i in range(n)]
    out = dict(K=cfg.K, voxel_size=cfg.voxel_size, max_depth=cfg.depth_trunc,
               depth=np.stack([f[0] for f in frames]), color=np.stack([f[1] for f in frames]),
               Tcw=np.stack([f[2] for f in frames]))
    grid = oracle.RefGrid(cfg.voxel_size, 8)
    grid_nf = oracle.RefGrid(cfg.voxel_size, 8)
    for i, (d, c, T) in enumerate(frames):
        filtered = ref.filter_shadow_points(d, delta_depth=None)          # reference function, unmodified
        out[f"filtered_{i}"] = filtered
        for tag, dd, g in (("f", filtered, grid), ("nf", d, grid_nf)):
            pc = ref.depth2pointcloud(np.ascontiguousarray(dd, np.float32), c, cfg.fx, cfg.fy, cfg.cx, cfg.cy,
                                      cfg.depth_trunc)                    # reference function, unmodified
            inv_pose = S.inv_T(T)
            points_world = (inv_pose[:3, :3] @ pc.points.T + inv_pose[:3, 3].reshape(3, 1)).T  # voxel_grid.py:262-265
            pts = np.ascontiguousarray(points_world, dtype=np.float32)
            cols = np.ascontiguousarray(pc.colors, dtype=np.float32)
            g.integrate(pts, cols)
            if tag == "f":
                out[f"points_{i}"] = pts
                out[f"colors_{i}"] = cols
    for tag, g in (("f", grid), ("nf", grid_nf)):
        d = sort_dump(g.dump_blocks())
        out[f"{tag}_keys"], out[f"{tag}_hashes"], out[f"{tag}_count"] = d["keys"], d["hashes"], d["count"]
        out[f"{tag}_pos_sum"], out[f"{tag}_col_sum"] = d["pos_sum"], d["col_sum"]
    np.savez_compressed(os.path.join(GOLDEN, "frontend_T0.npz"), **out)
    nrem = [int((out[f"filtered_{i}"] != frames[i][0]).sum()) for i in range(n)]
    print("frontend_T0: shadow filter removed", nrem, "pixels;", len(out["f_keys"]), "blocks (filtered),",
          len(out["nf_keys"]), "blocks (unfiltered)")


if __name__ == "__main__":
    main()
