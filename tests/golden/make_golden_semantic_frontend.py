"""Golden vectors of the semantic integrator's per-frame front-end ("per-frame semantic
pipeline"), produced in the build container by the REFERENCE'S OWN Python functions imported from /root/reference
(pyslam/utilities/depth.py: filter_shadow_points :103-146, depth2pointcloud :45-85 with semantic / object-id
images), the 3-line world transform of volumetric_integrator_voxel_semantic_grid.py:411-436 restated in numpy, and
the UNMODIFIED compiled VoxelBlockSemanticProbabilisticGrid (oracle/_ref):
    python tests/golden/make_golden_semantic_frontend.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pyslam_b200 import synthetic as S  # noqa: E402
from tests._util import GOLDEN, sort_dump  # noqa: E402
from tests.golden.make_golden_frontend import load_reference_depth_module  # noqa: E402

VOXEL = 0.06


def main():
    ref = load_reference_depth_module()
    cfg = S.CONFIGS["T0"]
    rng = np.random.default_rng(21)
    n = 3
    out = dict(K=cfg.K, voxel_size=VOXEL, max_depth=cfg.depth_trunc, depth_threshold=1.5, depth_decay_rate=0.6)
    g = oracle.RefSemanticGrid(VOXEL, "probabilistic")
    g.set_depth_threshold(1.5)
    g.set_depth_decay_rate(0.6)
    depth_l, color_l, cls_l, obj_l, T_l = [], [], [], [], []
    for i in range(n):
        d, c, T = S.render_frame(cfg, i)
        h, w = d.shape
        cls_img = (1 + np.argmax(c.astype(np.int32), axis=2)).astype(np.int32)
        obj_img = (cls_img * 10 + (np.arange(h)[:, None] // 36)).astype(np.int32)
        flip = rng.random((h, w)) < 0.1
        cls_img = np.where(flip, rng.integers(-1, 2, (h, w)), cls_img).astype(np.int32)
        obj_img = np.where(flip, rng.integers(-1, 2, (h, w)), obj_img).astype(np.int32)
        filtered = ref.filter_shadow_points(d, delta_depth=None)                                  # reference, unmodified
        depth_filtered = np.ascontiguousarray(filtered, dtype=np.float32)                          # semantic_grid.py:344-346
        pc = ref.depth2pointcloud(depth_filtered, c, cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.depth_trunc,
                                  semantic_image=cls_img, object_ids_image=obj_img)               # reference, unmodified
        depths = np.ascontiguousarray(pc.points[:, 2], dtype=np.float32)                           # :408-409
        inv_pose = S.inv_T(T)
        points_world = (inv_pose[:3, :3] @ pc.points.T + inv_pose[:3, 3].reshape(3, 1)).T          # :411-415
        colors = np.ascontiguousarray(pc.colors, dtype=np.float32)                                 # :424-426
        points = np.ascontiguousarray(points_world, dtype=np.float32)                              # :434-436 (float32 default)
        g.integrate(points, colors, np.ascontiguousarray(pc.semantics, np.int32),
                    np.ascontiguousarray(pc.object_ids, np.int32), depths)
        depth_l.append(d), color_l.append(c), cls_l.append(cls_img), obj_l.append(obj_img), T_l.append(T)
    out.update(depth=np.stack(depth_l), color=np.stack(color_l), class_image=np.stack(cls_l),
               object_image=np.stack(obj_l), Tcw=np.stack(T_l))
    d = sort_dump(g.dump_blocks(8))
    assert d["aux"].max() <= 8
    for k, v in d.items():
        out[k] = v
    np.savez_compressed(os.path.join(GOLDEN, "semantic_frontend_T0.npz"), **out)
    print("semantic_frontend_T0:", len(d["keys"]), "blocks,", int((d["count"] > 0).sum()), "voxels, max labels",
          int(d["aux"].max()), "size", os.path.getsize(os.path.join(GOLDEN, "semantic_frontend_T0.npz")))


if __name__ == "__main__":
    main()
