"""Regenerate the committed golden fixtures (run in the build container, where /root/reference
exists):  python tests/golden/make_golden.py

tsdf_T0.npz     inputs (depth, colour, poses of 4 T0 frames) + the outputs of BOTH restatements of Open3D's
                ScalableTSDFVolume(voxel, trunc, RGB8, 16, 4): sorted block keys, reference BlockKeyHash, voxel
                planes, per-frame touched sets, canonical welded mesh from oracle/tsdf_oracle.c (block based, the
                kernels' arithmetic contract v3), checked here - before anything is written - to be bit-identical
                in keys, tsdf, weight, mesh topology and vertex positions to oracle/open3d_order.c (the literal
                unit-based restatement, whose float64 colours are stored as `o3d_rgb64`).  Open3D itself is not
                available -> parity vs a RUNNING Open3D stays unpinned.   `python make_golden.py tsdf` remakes
                this file only.
refgrid_T0.npz  world-space float32 points / colours derived from the same frames the way the
                reference front-end does (pyslam/utilities/depth.py:45-85,
                pyslam/dense/volumetric_integrator_voxel_grid.py:262-281) + the outputs of the
                UNMODIFIED compiled reference volumetric::VoxelBlockGrid (oracle/_ref):
                block keys, BlockKeyHash, counts, position / colour sums, get_voxels(min_count).
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pyslam_b200 import synthetic as S  # noqa: E402
from tests._util import GOLDEN, sort_dump, sorted_keys  # noqa: E402

N_FRAMES = 4


def frontend_points(depth, rgb, K, Tcw, max_depth):
    """depth2pointcloud + world transform exactly as the reference front-end (float64 -> float32)."""
    fx, fy, cx, cy = K
    valid = (depth > 0.0) & (depth < max_depth)
    z = depth[valid].astype(np.float64)
    rows, cols = np.where(valid)
    x = (cols - cx) * z * (1.0 / fx)
    y = (rows - cy) * z * (1.0 / fy)
    pts = np.column_stack([x, y, z])
    colors = rgb[valid] / 255.0
    Twc = S.inv_T(Tcw)
    R, t = Twc[:3, :3], Twc[:3, 3]
    pw = np.stack([pts[:, 0] * R[a, 0] + pts[:, 1] * R[a, 1] + pts[:, 2] * R[a, 2] + t[a]
                   for a in range(3)], axis=1)
    return np.ascontiguousarray(pw, np.float32), np.ascontiguousarray(colors, np.float32)


def main():
    oracle.build()
    cfg = S.CONFIGS["T0"]
    frames = [S.render_frame(cfg, i) for i in range(N_FRAMES)]
    out = dict(n_frames=N_FRAMES, K=cfg.K, voxel_size=cfg.voxel_size, sdf_trunc=cfg.sdf_trunc,
               depth_trunc=cfg.depth_trunc,
               depth=np.stack([f[0] for f in frames]), color=np.stack([f[1] for f in frames]),
               Tcw=np.stack([f[2] for f in frames]))
    o = oracle.TsdfOracle(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc)
    for i, (d, c, T) in enumerate(frames):
        o.integrate(d, c, cfg.K, T)
        out[f"touched_{i}"] = sorted_keys(o.last_touched())
    dump = sort_dump(o.dump_blocks())
    out.update(keys=dump["keys"], hashes=dump["hashes"], vox=dump["vox"])
    m = o.extract_mesh()
    cm = oracle.canonical_mesh(m["vertices"], m["colors"], m["edges"], m["triangles"])
    out.update(mesh_vertices=cm["vertices"], mesh_colors=cm["colors"], mesh_edges=cm["edges"],
               mesh_triangles=cm["triangles"])
    # the literal Open3D-order restatement must agree before the fixture is written
    o3 = oracle.Open3DOrderVolume(cfg.voxel_size, cfg.sdf_trunc, 16, 4)
    for d, c, T in frames:
        o3.integrate(d, c, cfg.K, T, cfg.depth_trunc)
    d3 = sort_dump(o3.dump_blocks())
    assert np.array_equal(d3["keys"], dump["keys"])
    assert np.array_equal(d3["vox"][:, :2], dump["vox"][:, :2].astype(np.float64)), "tsdf / weight differ"
    assert np.abs(d3["vox"][:, 2:] - dump["vox"][:, 2:]).max() < 1e-4
    m3 = o3.extract_triangle_mesh()
    c3 = oracle.canonical_mesh(m3["vertices"], m3["colors"], m3["edges"], m3["triangles"])
    assert np.array_equal(c3["edges"], cm["edges"]) and np.array_equal(c3["triangles"], cm["triangles"])
    assert np.array_equal(c3["vertices"], cm["vertices"]) and np.abs(c3["colors"] - cm["colors"]).max() < 1e-6
    out.update(o3d_rgb64=d3["vox"][:, 2:], o3d_mesh_colors=c3["colors"], contract=3)
    np.savez_compressed(os.path.join(GOLDEN, "tsdf_T0.npz"), **out)
    print("tsdf_T0:", len(dump["keys"]), "blocks,", len(cm["vertices"]), "vertices,",
          len(cm["triangles"]), "triangles")
    if len(sys.argv) > 1 and sys.argv[1] == "tsdf":
        return

    assert oracle.have_ref(), "the compiled reference is required to make refgrid_T0.npz"
    g = oracle.RefGrid(cfg.voxel_size, 8)
    pts_all, col_all, counts = [], [], []
    for d, c, T in frames[:3]:
        p, col = frontend_points(d, c, cfg.K, T, cfg.depth_trunc)
        g.integrate(p, col)
        pts_all.append(p)
        col_all.append(col)
        counts.append(len(p))
    # spatial queries + carving with the 4th frame's camera
    d3, _, T3 = frames[3]
    Kf = np.array(cfg.K, np.float32)
    fp, fc = g.get_voxels_in_frustum(Kf, cfg.width, cfg.height, T3, min_count=1, depth_max=3.0, depth_min=0.05)
    allp = np.concatenate(pts_all)
    bbox = np.concatenate([np.quantile(allp, 0.2, axis=0), np.quantile(allp, 0.75, axis=0)])
    bp, bc = g.get_voxels_in_bb(bbox, min_count=1)
    g2 = oracle.RefGrid(cfg.voxel_size, 8)
    for p_, c_ in zip(pts_all, col_all):
        g2.integrate(p_, c_)
    # carve against a depth image in which the scene has receded by 0.3 m (valid pixels only): every
    # stored voxel seen by this camera now floats in front of the observed surface
    d3 = np.where(d3 > 0, d3 + np.float32(0.3), d3).astype(np.float32)
    g2.carve(Kf, cfg.width, cfg.height, T3, d3, depth_threshold=0.05, depth_max=3.0, depth_min=0.05)
    carved = sort_dump(g2.dump_blocks())
    rd = sort_dump(g.dump_blocks())
    vp, vc = g.get_voxels(min_count=2)
    order = np.lexsort((vp[:, 2], vp[:, 1], vp[:, 0]))
    np.savez_compressed(os.path.join(GOLDEN, "refgrid_T0.npz"), voxel_size=cfg.voxel_size,
                        points=np.concatenate(pts_all), colors=np.concatenate(col_all),
                        frame_counts=np.array(counts), keys=rd["keys"], hashes=rd["hashes"],
                        count=rd["count"], pos_sum=rd["pos_sum"], col_sum=rd["col_sum"],
                        voxels_min2_points=vp[order], voxels_min2_colors=vc[order],
                        query_K=Kf, query_Tcw=T3, query_depth=d3, query_bbox=bbox,
                        frustum_points=fp[np.lexsort((fp[:, 2], fp[:, 1], fp[:, 0]))],
                        bbox_points=bp[np.lexsort((bp[:, 2], bp[:, 1], bp[:, 0]))],
                        carved_count=carved["count"])
    print("queries:", len(fp), "in frustum,", len(bp), "in bbox,",
          int((rd["count"] > 0).sum() - (carved["count"] > 0).sum()), "voxels carved")
    print("refgrid_T0:", len(rd["keys"]), "blocks,", int((rd["count"] > 0).sum()), "voxels,",
          len(vp), "voxels with count >= 2")


if __name__ == "__main__":
    main()
