"""Golden vectors of the instance -> object association row (per-frame semantic pipeline),
produced in the build container by the UNMODIFIED compiled reference (oracle/_ref/libref_semantic.so):
    python tests/golden/make_golden_semantic_assoc.py

For each kind ("vote", "prob") the reference integrator's loop body
(pyslam/dense/volumetric_integrator_voxel_semantic_grid.py:349-461) is replayed over 4 synthetic T0 frames whose
2-D instance ids change from frame to frame: assign_object_ids_to_instance_ids (with carving) -> remap_instance_ids
-> depth2pointcloud -> integrate.  Stored: the label images, every frame's instance -> object map, the final block
dump and the object-id allocator's final value."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pyslam_b200 import remap_instance_ids, synthetic as S  # noqa: E402
from tests._util import GOLDEN, sort_dump  # noqa: E402

VOXEL = 0.06
N_FRAMES = 4
PARAMS = dict(depth_threshold=0.08, do_carving=True, min_vote_ratio=0.5, min_votes=3, depth_max=4.0, depth_min=0.05)


def label_images(cfg, i, rng):
    d, c, T = S.render_frame(cfg, i)
    h, w = d.shape
    cls_img = (1 + np.argmax(c.astype(np.int32), axis=2)).astype(np.int32)
    band = (np.arange(w)[None, :] * 3 // w).astype(np.int32)
    inst_img = np.where(cls_img == 1, 0, 20 * (i + 1) + cls_img * 3 + band).astype(np.int32)   # class 1 = "stuff"
    flip = rng.random((h, w)) < 0.05
    cls_img = np.where(flip, -1, cls_img).astype(np.int32)                                     # unlabelled pixels
    return d, c, T, cls_img, inst_img


def frame_points(d, c, K, Tcw, max_depth, cls_img, obj_img):
    """The reference front-end's arithmetic with an explicit (BLAS-free) transform, float32 points."""
    T = S.inv_T(Tcw)
    valid = (d > 0) & (d < max_depth)
    z = d[valid].astype(np.float64)
    rows, cols = np.where(valid)
    x, y = (cols - K[2]) * z * (1.0 / K[0]), (rows - K[3]) * z * (1.0 / K[1])
    pw = np.stack([x * T[r, 0] + y * T[r, 1] + z * T[r, 2] + T[r, 3] for r in range(3)], axis=1).astype(np.float32)
    return pw, (c[valid] / 255.0).astype(np.float32), cls_img[valid], obj_img[valid], d[valid]


def main():
    oracle.build()
    cfg = S.CONFIGS["T0"]
    out = dict(voxel_size=VOXEL, n_frames=N_FRAMES, K=cfg.K, max_depth=cfg.depth_trunc,
               **{f"param_{k}": v for k, v in PARAMS.items()})
    for tag, kind in (("vote", "voting"), ("prob", "probabilistic")):
        rng = np.random.default_rng(33)
        oracle.RefSemanticGrid.set_next_object_id(1)
        g = oracle.RefSemanticGrid(VOXEL, kind)
        g.set_depth_threshold(10.0)
        for i in range(N_FRAMES):
            d, c, T, cls_img, inst_img = label_images(cfg, i, rng)
            Kf = np.array(cfg.K, np.float32)
            m = g.assign_object_ids_to_instance_ids(Kf, cfg.width, cfg.height, T, PARAMS["depth_max"],
                                                    PARAMS["depth_min"], cls_img, inst_img, d,
                                                    PARAMS["depth_threshold"], PARAMS["do_carving"],
                                                    PARAMS["min_vote_ratio"], PARAMS["min_votes"])
            obj_img = remap_instance_ids(inst_img, m)
            g.integrate(*frame_points(d, c, cfg.K, T, cfg.depth_trunc, cls_img, obj_img))
            keys = np.array(sorted(m), np.int32)
            out[f"{tag}_map_inst_{i}"] = keys
            out[f"{tag}_map_obj_{i}"] = np.array([m[k] for k in keys], np.int32)
            if tag == "vote":
                out.update({f"depth_{i}": d, f"color_{i}": c, f"Tcw_{i}": T, f"class_image_{i}": cls_img,
                            f"instance_image_{i}": inst_img})
            print(tag, "frame", i, "map", m)
        dmp = sort_dump(g.dump_blocks(8))
        for k in ("keys", "count", "object_id", "class_id", "confidence", "aux"):
            out[f"{tag}_{k}"] = dmp[k]
        out[f"{tag}_next_object_id"] = oracle.RefSemanticGrid.get_next_object_id()
        occ = dmp["count"] > 0
        print(tag, "voxels", int(occ.sum()), "object ids", np.unique(dmp["object_id"][occ]).tolist(),
              "next id", out[f"{tag}_next_object_id"])
    np.savez_compressed(os.path.join(GOLDEN, "semantic_assoc_T0.npz"), **out)
    print("size", os.path.getsize(os.path.join(GOLDEN, "semantic_assoc_T0.npz")))


if __name__ == "__main__":
    main()
