"""Golden vectors of the semantic-fusion row, produced by the UNMODIFIED compiled
reference grids (oracle/_ref/libref_semantic.so) in the build container:
    python tests/golden/make_golden_semantic.py

semantic_T0.npz  per kind ("vote", "prob"): the label / depth streams fed to
                 VoxelBlockSemanticGrid / VoxelBlockSemanticProbabilisticGrid over 4 T0 frames (points built
                 like the reference front-end, volumetric_integrator_voxel_semantic_grid.py:392-461) and the
                 reference's block dump: keys, BlockKeyHash, count, float64 position sums, colour sums, label,
                 confidence, counter / label evidence in std::map order, plus get_voxels(min_count=2, 0.4)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pyslam_b200 import synthetic as S  # noqa: E402
from tests._util import GOLDEN, sort_dump  # noqa: E402

N_FRAMES = 4
VOXEL = 0.06


def semantic_frame(cfg, i, rng, dtype):
    """points (world), colors [0,1] f32, class ids, object ids, camera depths of one synthetic frame."""
    d, c, T = S.render_frame(cfg, i)
    fx, fy, cx, cy = cfg.K
    valid = (d > 0.0) & (d < cfg.depth_trunc)
    z = d[valid].astype(np.float64)
    rows, cols = np.where(valid)
    pc = np.column_stack([(cols - cx) * z / fx, (rows - cy) * z / fy, z])
    Twc = S.inv_T(T)
    pw = (Twc[:3, :3] @ pc.T + Twc[:3, 3].reshape(3, 1)).T
    rgb = c[valid]
    cls = 1 + np.argmax(rgb.astype(np.int32), axis=1).astype(np.int32)          # "class" from the dominant channel
    inst = (cls * 10 + (rows // 36)).astype(np.int32)                             # two objects per class
    flip = rng.random(len(cls)) < 0.12                                            # 12 % label noise
    cls = np.where(flip, rng.integers(0, 2, len(cls)), cls).astype(np.int32)
    inst = np.where(flip, rng.integers(-1, 1, len(cls)), inst).astype(np.int32)   # incl. the invalid id -1
    depths = (z * 2.0).astype(np.float32)                                         # stretch past the thresholds
    return (np.ascontiguousarray(pw, dtype), np.ascontiguousarray(rgb / 255.0, np.float32), cls, inst, depths)


def main():
    oracle.build()
    assert oracle.have_ref_semantic()
    cfg = S.CONFIGS["T0"]
    out = dict(voxel_size=VOXEL, n_frames=N_FRAMES)
    for tag, kind, dtype, thr, rate in (("vote", "voting", np.float32, 3.0, 0.0),
                                         ("prob", "probabilistic", np.float64, 2.95, 2.0)):
        rng = np.random.default_rng(7)
        g = oracle.RefSemanticGrid(VOXEL, kind)
        g.set_depth_threshold(thr)
        if kind == "probabilistic":
            g.set_depth_decay_rate(rate)
        out[f"{tag}_depth_threshold"], out[f"{tag}_depth_decay_rate"] = thr, rate
        for i in range(N_FRAMES):
            p, col, cls, inst, dep = semantic_frame(cfg, i, rng, dtype)
            g.integrate(p, col, cls, inst, dep)
            out.update({f"{tag}_points_{i}": p, f"{tag}_colors_{i}": col, f"{tag}_cls_{i}": cls,
                        f"{tag}_inst_{i}": inst, f"{tag}_depths_{i}": dep})
        d = sort_dump(g.dump_blocks(8))
        for k, v in d.items():
            out[f"{tag}_{k}"] = v
        v = g.get_voxels(2, 0.4)
        order = np.lexsort((v["points"][:, 2], v["points"][:, 1], v["points"][:, 0]))
        for k, a in v.items():
            out[f"{tag}_voxels_{k}"] = a[order]
        occ = d["count"] > 0
        assert kind == "voting" or d["aux"].max() <= 8, "keep the golden scenario within 8 label pairs per voxel"
        print(tag, "blocks", len(d["keys"]), "voxels", int(occ.sum()), "max count", int(d["count"].max()),
              "labels/voxel max", int(d["aux"].max()) if kind == "probabilistic" else "-",
              "get_voxels(2,0.4)", len(v["points"]), "mean conf", float(d["confidence"][occ].mean()))
    np.savez_compressed(os.path.join(GOLDEN, "semantic_T0.npz"), **out)
    print("size", os.path.getsize(os.path.join(GOLDEN, "semantic_T0.npz")))


if __name__ == "__main__":
    main()
