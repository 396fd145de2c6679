"""GPU: parity of the CUDA TSDF path (through the C ABI) against the CPU twin (oracle/tsdf_oracle.c) and the
golden fixtures.  Bars: block keys, hashes, touched sets, weights, tsdf, rgb, triangle topology,
canonical edge ids and float64 vertex positions / colours are BIT-EXACT against the twin (both sides execute the
same IEEE operations in Open3D's order).  The comparison with the literal Open3D-order restatement
(oracle/open3d_order.c) is tests/test_gpu_open3d.py."""

import os

import numpy as np
import pytest

import oracle
from pyslam_b200 import B200TsdfVolume
from pyslam_b200 import synthetic as S
from tests._util import GOLDEN, blocks_checksum, sort_dump, sorted_keys

pytestmark = pytest.mark.gpu


def _pair(cfg, capacity=1 << 15, stride=4, unit=16):
    vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=capacity,
                         depth_sampling_stride=stride, volume_unit_resolution=unit)
    orc = oracle.TsdfOracle(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, stride=stride, unit_resolution=unit)
    return vol, orc


def _assert_same_volume(vol, orc):
    a = sort_dump(vol.dump_blocks())
    b = sort_dump(orc.dump_blocks())
    assert np.array_equal(a["keys"], b["keys"])
    assert np.array_equal(a["hashes"], b["hashes"])
    assert np.array_equal(a["vox"][:, 1], b["vox"][:, 1])            # weights
    assert np.array_equal(a["vox"], b["vox"])                        # tsdf, rgb: bit-exact
    return a


def test_golden_fixture_bit_exact():
    g = np.load(os.path.join(GOLDEN, "tsdf_T0.npz"))
    vol = B200TsdfVolume(float(g["voxel_size"]), float(g["sdf_trunc"]), float(g["depth_trunc"]),
                         capacity_blocks=4096)
    for i in range(int(g["n_frames"])):
        vol.integrate(g["depth"][i], g["color"][i], g["K"], g["Tcw"][i])
        assert np.array_equal(sorted_keys(vol.last_touched_keys()), g[f"touched_{i}"])
        touched, _ = vol.last_frame_stats()
        assert touched == len(g[f"touched_{i}"])
    d = sort_dump(vol.dump_blocks())
    assert np.array_equal(d["keys"], g["keys"])
    assert np.array_equal(d["hashes"], g["hashes"])
    assert np.array_equal(d["vox"], g["vox"])
    m = vol.extract_mesh()
    cm = oracle.canonical_mesh(m.vertices, m.vertex_colors,
                               m.edge_ids, m.triangles)
    assert np.array_equal(cm["edges"], g["mesh_edges"])
    assert np.array_equal(cm["triangles"], g["mesh_triangles"])
    assert np.array_equal(cm["vertices"], g["mesh_vertices"])
    assert np.array_equal(cm["colors"], g["mesh_colors"])
    assert m.vertex_normals.shape == (0, 3)


@pytest.mark.parametrize("cfg_name,frames,stride,unit", [
    ("T0", list(range(6)), 4, 16),
    ("T0", list(range(6)), 4, 8),      # decision D1: float32 pyslam key range, 8^3 units
    ("T0", [0, 3], 1, 16),             # stride-1 superset mode
    ("C1", [0, 1, 2, 50], 4, 16),      # config 1 shape: 320x240, 1 cm
    ("C1", [0, 1, 2, 50], 4, 8),
    ("C4", [0, 1], 4, 16),             # ScanNet shape, 4 mm
    ("C3", [0], 4, 16),                # Replica shape 1200x680, 5 mm
    ("C5", [0, 40], 4, 16),            # KITTI shape 1241x376 (W % 16 != 0: plain-load path), 10 cm, tau 0.4
    ("C5", [0, 40], 4, 8),
])
def test_integrate_matches_oracle(cfg_name, frames, stride, unit):
    cfg = S.CONFIGS[cfg_name]
    vol, orc = _pair(cfg, capacity=1 << 16, stride=stride, unit=unit)
    total_new = 0
    for i in frames:
        d, c, T = S.render_frame(cfg, i)
        vol.integrate(d, c, cfg.K, T)
        n = orc.integrate(d, c, cfg.K, T)
        assert np.array_equal(sorted_keys(vol.last_touched_keys()), sorted_keys(orc.last_touched()))
        touched, new = vol.last_frame_stats()
        assert touched == n
        total_new += new
    assert total_new == orc.num_blocks() == vol.num_blocks()
    _assert_same_volume(vol, orc)
    updates, launches = vol.counters()
    assert launches >= 2 * len(frames) and updates > 0


def test_full_size_tum_frames_match_oracle_and_properties():
    """BASELINE config 2 at full size (640x480, 5 mm): checksum-of-checksums vs the oracle plus
    size-independent properties (integer weights, bounded tsdf / colour, unique keys)."""
    cfg = S.CONFIGS["C2"]
    vol, orc = _pair(cfg, capacity=1 << 16)
    for i in (0, 1, 2):
        d, c, T = S.render_frame(cfg, i)
        vol.integrate(d, c, cfg.K, T)
        orc.integrate(d, c, cfg.K, T)
    a = _assert_same_volume(vol, orc)
    assert np.array_equal(blocks_checksum(a), blocks_checksum(sort_dump(orc.dump_blocks())))
    assert len(np.unique(a["keys"], axis=0)) == len(a["keys"])
    w = a["vox"][:, 1]
    assert np.array_equal(w, np.round(w)) and w.max() == 3.0
    assert a["vox"][:, 0].min() >= -1.0 and a["vox"][:, 0].max() <= 1.0
    assert a["vox"][:, 2:].min() >= 0.0 and a["vox"][:, 2:].max() <= 255.0
    # reported hash is the reference's BlockKeyHash (sign-extending u64 arithmetic)
    k = a["keys"].astype(np.int64).astype(np.uint64)
    assert np.array_equal(a["hashes"], k[:, 0] ^ (k[:, 1] << np.uint64(1)) ^ (k[:, 2] << np.uint64(2)))


def test_mesh_matches_oracle_in_float64():
    cfg = S.CONFIGS["C1"]
    vol, orc = _pair(cfg, capacity=1 << 15)
    for i in (0, 1, 2, 3):
        d, c, T = S.render_frame(cfg, i)
        vol.integrate(d, c, cfg.K, T)
        orc.integrate(d, c, cfg.K, T)
    m = vol.extract_mesh()
    ref = orc.extract_mesh()
    assert len(m.vertices) == len(ref["vertices"]) and len(m.triangles) == len(ref["triangles"]) > 1000
    a = oracle.canonical_mesh(m.vertices, m.vertex_colors,
                              m.edge_ids, m.triangles)
    b = oracle.canonical_mesh(ref["vertices"], ref["colors"], ref["edges"], ref["triangles"])
    assert np.array_equal(a["edges"], b["edges"])
    assert np.array_equal(a["triangles"], b["triangles"])
    assert np.array_equal(a["vertices"], b["vertices"])
    assert np.array_equal(a["colors"], b["colors"])
    assert m.vertices.dtype == np.float64 and m.vertex_colors.dtype == np.float64   # like Open3D's TriangleMesh
    assert a["colors"].min() >= 0.0 and a["colors"].max() <= 1.0 + 1e-12
    # a second extraction of the same volume is identical (deterministic count -> scan -> emit)
    m2 = vol.extract_mesh()
    assert np.array_equal(m.triangles, m2.triangles) and np.array_equal(m.vertices, m2.vertices)
    # the extraction narrows its work through the blocks' sign summaries without losing a tile (the equalities above)
    st = vol.last_mesh_stats()
    assert st["blocks"] == vol.num_blocks()
    assert 0 < st["vertex_blocks"] <= st["tiles_with_both_signs"] <= st["candidate_tiles"] <= st["blocks"]
    assert 0 < st["triangle_blocks"] <= st["tiles_with_both_signs"]


def test_upload_dump_round_trip_and_sphere_mesh():
    """dump -> reset -> upload -> dump is the identity; an analytic sphere meshes closed."""
    vs, tau, r = 0.02, 0.08, 0.5
    vol = B200TsdfVolume(vs, tau, 4.0, capacity_blocks=4096)
    nb = int(np.ceil((r + 3 * tau) / (8 * vs)))
    l = np.arange(512)
    lx, ly, lz = l % 8, (l // 8) % 8, l // 64
    keys, vox = [], []
    for bx in range(-nb, nb):
        for by in range(-nb, nb):
            for bz in range(-nb, nb):
                c = np.stack([(bx * 8 + lx + 0.5) * vs, (by * 8 + ly + 0.5) * vs, (bz * 8 + lz + 0.5) * vs], 1)
                v = np.zeros((5, 512), np.float32)
                v[0] = np.clip((np.linalg.norm(c, axis=1) - r) / tau, -1, 1)
                v[1] = 1.0
                v[2:] = np.array([[200.0], [100.0], [50.0]])
                keys.append((bx, by, bz))
                vox.append(v)
    keys, vox = np.array(keys, np.int32), np.stack(vox)
    vol.upload_blocks(keys, vox)
    d = sort_dump(vol.dump_blocks())
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    assert np.array_equal(d["keys"], keys[order]) and np.array_equal(d["vox"], vox[order])
    m = vol.extract_mesh()
    T = m.triangles
    e = np.concatenate([T[:, [0, 1]], T[:, [1, 2]], T[:, [2, 0]]])
    _, cnt = np.unique(np.sort(e, axis=1), axis=0, return_counts=True)
    assert np.all(cnt == 2)
    assert len(m.vertices) - len(cnt) + len(T) == 2
    assert np.max(np.abs(np.linalg.norm(m.vertices, axis=1) - r)) < 0.2 * vs
    # same mesh as the oracle on the same volume
    orc = oracle.TsdfOracle(vs, tau, 4.0)
    for k, v in zip(keys, vox):
        orc.set_block(k, v)
    ref = orc.extract_mesh()
    a = oracle.canonical_mesh(m.vertices, m.vertex_colors,
                              m.edge_ids, m.triangles)
    b = oracle.canonical_mesh(ref["vertices"], ref["colors"], ref["edges"], ref["triangles"])
    for name in ("edges", "triangles", "vertices", "colors"):
        assert np.array_equal(a[name], b[name]), name


def test_reset_empty_and_ragged_inputs():
    cfg = S.CONFIGS["T0"]
    vol, orc = _pair(cfg, capacity=2048)
    assert vol.num_blocks() == 0
    m = vol.extract_mesh()
    assert m.vertices.shape == (0, 3) and m.triangles.shape == (0, 3)
    d, c, T = S.render_frame(cfg, 0)
    # all-invalid depth (zeros, negatives, NaN, beyond depth_trunc) touches nothing
    bad = np.zeros_like(d)
    bad[::2] = -1.0
    bad[1::3] = np.nan
    bad[5] = cfg.depth_trunc + 1.0
    vol.integrate(bad, c, cfg.K, T)
    assert vol.last_frame_stats() == (0, 0) and vol.num_blocks() == 0
    # ragged size (not a multiple of the stride or of the allocation tile)
    dr, cr = np.ascontiguousarray(d[:61, :83]), np.ascontiguousarray(c[:61, :83])
    vol.integrate(dr, cr, cfg.K, T)
    orc.integrate(dr, cr, cfg.K, T)
    _assert_same_volume(vol, orc)
    vol.reset()
    assert vol.num_blocks() == 0
    orc.reset()
    vol.integrate(d, c, cfg.K, T)
    orc.integrate(d, c, cfg.K, T)
    _assert_same_volume(vol, orc)


def test_error_behaviour():
    cfg = S.CONFIGS["T0"]
    d, c, T = S.render_frame(cfg, 0)
    with pytest.raises(RuntimeError):
        B200TsdfVolume(0.0, 0.04)                      # invalid voxel size
    with pytest.raises(RuntimeError):
        B200TsdfVolume(0.01, 0.04, block_size=16)      # only the reference default block size 8
    vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=64)
    with pytest.raises(RuntimeError):
        vol.integrate(d, c[:, :, :2], cfg.K, T)        # colour must be [H,W,3]
    with pytest.raises(RuntimeError):
        vol.integrate(d, c.astype(np.float32), cfg.K, T)
    with pytest.raises(RuntimeError):
        vol.integrate(d[None], c, cfg.K, T)
    # pool overflow is reported, not silently dropped
    vol.integrate(d, c, cfg.K, T)
    with pytest.raises(RuntimeError, match="capacity"):
        vol.synchronize()


def test_device_pointer_inputs_and_batch_equal_host_path():
    import torch
    cfg = S.CONFIGS["T0"]
    frames = [S.render_frame(cfg, i) for i in range(4)]
    host, _ = _pair(cfg)
    dev, _ = _pair(cfg)
    bat, _ = _pair(cfg)
    for d, c, T in frames:
        host.integrate(d, c, cfg.K, T)
        dev.integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), cfg.K, T)
    torch.cuda.synchronize()
    bat.integrate_batch(np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames]), cfg.K,
                        np.stack([f[2] for f in frames]))
    a, b, c_ = (sort_dump(v.dump_blocks()) for v in (host, dev, bat))
    for name in ("keys", "hashes", "vox"):
        assert np.array_equal(a[name], b[name]) and np.array_equal(a[name], c_[name])


@pytest.mark.parametrize("n", [4, 3, 8])   # power-of-two rank counts take a mask, the others the 64-bit modulo
def test_sharded_volumes_partition_the_blocks(n):
    """Hash-bucket sharding: shard r owns BlockKeyHash % n == r; the union of the
    shards equals the unsharded volume bit for bit and no block is owned twice."""
    cfg = S.CONFIGS["T0"]
    frames = [S.render_frame(cfg, i) for i in range(3)]
    full, _ = _pair(cfg)
    shards = [B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=4096,
                             shard_rank=r, shard_count=n) for r in range(n)]
    for d, c, T in frames:
        full.integrate(d, c, cfg.K, T)
        for s in shards:
            s.integrate(d, c, cfg.K, T)
    ref = sort_dump(full.dump_blocks())
    parts = [s.dump_blocks() for s in shards]
    for r, p in enumerate(parts):
        assert np.all(p["hashes"] % np.uint64(n) == r)
    merged = sort_dump({k: np.concatenate([p[k] for p in parts]) for k in ("keys", "hashes", "vox")})
    for name in ("keys", "hashes", "vox"):
        assert np.array_equal(merged[name], ref[name])


def test_point_cloud_extraction_matches_definition():
    cfg = S.CONFIGS["T0"]
    vol, orc = _pair(cfg)
    for i in range(3):
        d, c, T = S.render_frame(cfg, i)
        vol.integrate(d, c, cfg.K, T)
    pc = vol.extract_point_cloud()
    dump = vol.dump_blocks()
    # count zero crossings on the host straight from the dump (A.4 ExtractPointCloud definition)
    idx = {tuple(k): i for i, k in enumerate(dump["keys"])}
    expect = 0
    for bi, key in enumerate(dump["keys"]):
        f = dump["vox"][bi, 0].reshape(8, 8, 8)   # [z, y, x]
        w = dump["vox"][bi, 1].reshape(8, 8, 8)
        for axis, dk in ((2, (1, 0, 0)), (1, (0, 1, 0)), (0, (0, 0, 1))):
            nk = (key[0] + dk[0], key[1] + dk[1], key[2] + dk[2])
            if nk in idx:
                fn = dump["vox"][idx[nk], 0].reshape(8, 8, 8)
                wn = dump["vox"][idx[nk], 1].reshape(8, 8, 8)
            else:
                fn, wn = np.zeros((8, 8, 8), np.float32), np.zeros((8, 8, 8), np.float32)
            f1 = np.concatenate([np.take(f, range(1, 8), axis), np.take(fn, [0], axis)], axis)
            w1 = np.concatenate([np.take(w, range(1, 8), axis), np.take(wn, [0], axis)], axis)
            ok0 = (w != 0) & (f < 0.98) & (f >= -0.98)
            ok1 = (w1 != 0) & (f1 < 0.98) & (f1 >= -0.98)
            expect += int((ok0 & ok1 & (f * f1 < 0)).sum())
    assert len(pc.points) == expect > 100
    assert pc.colors.min() >= 0.0 and pc.colors.max() <= 1.0 + 1e-6


@pytest.mark.parametrize("env", [{"B2V_TMA": "0"}, {"B2V_OVERLAP": "1"}, {"B2V_OVERLAP": "1", "B2V_INT_CTAS_PER_SM": "6"}])
def test_execution_variants_are_bit_identical(env, monkeypatch):
    """TMA tile staging vs plain loads, and allocate/integrate stream overlap, change scheduling only:
    the resulting volume must be bit-identical to the default path's (and hence to the oracle's)."""
    cfg = S.CONFIGS["C1"]
    frames = [S.render_frame(cfg, i) for i in (0, 1, 2, 3, 4, 5)]
    base, orc = _pair(cfg)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    var, _ = _pair(cfg)
    for d, c, T in frames:
        base.integrate(d, c, cfg.K, T)
        var.integrate(d, c, cfg.K, T)
        orc.integrate(d, c, cfg.K, T)
    a, b = sort_dump(base.dump_blocks()), sort_dump(var.dump_blocks())
    for name in ("keys", "hashes", "vox"):
        assert np.array_equal(a[name], b[name]), name
    _assert_same_volume(var, orc)


@pytest.mark.parametrize("group", [8, 3, 16, 32])
def test_fused_batch_equals_frame_by_frame_and_oracle(group):
    """integrate_batch fuses groups of `group` frames per block visit (8 by default: the unrolled kernel; larger
    groups take the constant-indexed loop); 19 or 75 frames exercise full and partial groups and the rotation of
    the group buffers.  Bit-identical to frame-by-frame and the oracle."""
    cfg = S.CONFIGS["C1"]
    n = 19 if group <= 8 else 75
    frames = [S.render_frame(cfg, i) for i in range(n)]
    D, Cc, T = (np.stack([f[k] for f in frames]) for k in range(3))
    fused, orc = _pair(cfg, capacity=1 << 16)
    plain, _ = _pair(cfg, capacity=1 << 16)
    plain.set_fusion(False)
    fused.set_group_size(group)
    with pytest.raises(RuntimeError):
        fused.set_group_size(33)
    fused.integrate_batch(D, Cc, cfg.K, T)
    plain.integrate_batch(D, Cc, cfg.K, T)
    for d, c, t in frames:
        orc.integrate(d, c, cfg.K, t)
    a, b = sort_dump(fused.dump_blocks()), sort_dump(plain.dump_blocks())
    for name in ("keys", "hashes", "vox"):
        assert np.array_equal(a[name], b[name]), name
    _assert_same_volume(fused, orc)
    upd_f, _ = fused.counters()
    upd_p, _ = plain.counters()
    assert upd_f == upd_p                               # every (block, frame) update is still applied
    assert fused.block_visits() < 0.5 * plain.block_visits() == 0.5 * upd_p   # ... with far fewer block visits
    touched, _ = fused.last_frame_stats()
    assert touched == len(orc.last_touched())
    # a second batch into the same volume, then single frames again (mode switches share the table)
    fused.integrate_batch(D[:5], Cc[:5], cfg.K, T[:5])
    fused.integrate(frames[5][0], frames[5][1], cfg.K, frames[5][2])
    for d, c, t in frames[:6]:
        orc.integrate(d, c, cfg.K, t)
    _assert_same_volume(fused, orc)
    assert np.array_equal(sorted_keys(fused.last_touched_keys()), sorted_keys(orc.last_touched()))


def test_mixed_call_patterns_stay_consistent_with_the_oracle():
    """Interleave every entry path (single frames from host / device memory, fused and un-fused batches of
    odd lengths, resets, mesh extraction in between) on one volume: the stream / event choreography must
    never change the result."""
    import torch
    cfg = S.CONFIGS["T0"]
    rng = np.random.default_rng(11)
    frames = [S.render_frame(cfg, i) for i in range(24)]
    vol, orc = _pair(cfg, capacity=8192)
    pos = 0

    def take(n):
        nonlocal pos
        idx = [(pos + k) % len(frames) for k in range(n)]
        pos += n
        return idx

    for round_ in range(3):
        for step in range(10):
            mode = int(rng.integers(0, 5))
            if mode == 0:                                   # single frame, host memory
                (i,) = take(1)
                vol.integrate(*frames[i][:2], cfg.K, frames[i][2])
            elif mode == 1:                                 # single frame, device memory
                (i,) = take(1)
                vol.integrate(torch.from_numpy(frames[i][0]).cuda(), torch.from_numpy(frames[i][1]).cuda(),
                              cfg.K, frames[i][2])
            elif mode in (2, 3):                            # batch (fused unless mode 3) of odd length
                idx = take(int(rng.integers(2, 20)))
                vol.set_fusion(mode == 2)
                D = np.stack([frames[i][0] for i in idx])
                Cc = np.stack([frames[i][1] for i in idx])
                T = np.stack([frames[i][2] for i in idx])
                if rng.integers(0, 2):
                    vol.integrate_batch(torch.from_numpy(D).cuda(), torch.from_numpy(Cc).cuda(), cfg.K, T)
                else:
                    vol.integrate_batch(D, Cc, cfg.K, T)
                for i in idx:
                    orc.integrate(*frames[i][:2], cfg.K, frames[i][2])
                continue
            else:                                           # an extraction in the middle of the stream
                vol.extract_mesh()
                continue
            orc.integrate(*frames[i][:2], cfg.K, frames[i][2])
        _assert_same_volume(vol, orc)
        if round_ == 1:
            vol.reset()
            orc.reset()
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------------
# rectification row: GPU remap == cv2.remap, bit for bit (tests/golden/remap_T0.npz)
# ---------------------------------------------------------------------------------------------------
def test_remap_equals_opencv_golden():
    from pyslam_b200 import remap
    g = np.load(os.path.join(GOLDEN, "remap_T0.npz"))
    col = remap(g["bgr"], g["map1"], g["map2"], "linear")
    assert np.array_equal(col, g["color_u"])
    assert np.array_equal(remap(g["bgr"], g["map1"], g["map2"], "linear", swap_rb=True), g["rgb_u"])
    dep = remap(g["depth"], g["map1"], g["map2"], "nearest")
    assert np.array_equal(dep, g["depth_u"])
    lab = remap(g["labels"], g["map1"], g["map2"], "nearest")
    assert np.array_equal(lab, g["labels_u"])
    assert (g["depth_u"] == 0).sum() > 50          # the zero border is exercised
    with pytest.raises(RuntimeError):
        remap(g["depth"], g["map1"], g["map2"], "linear")
    with pytest.raises(RuntimeError):
        remap(g["bgr"], g["map1"][:10], g["map2"][:10], "linear")


def test_volume_with_rectification_equals_prerectified_input():
    """set_rectification + raw frames == cv2-rectified frames fed directly, on the per-frame path and on the
    fused batch path (19 frames: two full groups + a ragged one)."""
    g = np.load(os.path.join(GOLDEN, "remap_T0.npz"))
    cfg = S.CONFIGS["T0"]
    K = (float(g["new_K"][0, 0]), float(g["new_K"][1, 1]), float(g["new_K"][0, 2]), float(g["new_K"][1, 2]))
    n = 19
    frames = [S.render_frame(cfg, i) for i in range(n)]
    raw_d = np.stack([f[0] for f in frames])
    raw_bgr = np.stack([np.ascontiguousarray(f[1][..., ::-1]) for f in frames])
    Ts = np.stack([f[2] for f in frames])
    from pyslam_b200 import remap
    rect_d = np.stack([remap(d, g["map1"], g["map2"], "nearest") for d in raw_d])
    rect_rgb = np.stack([remap(c, g["map1"], g["map2"], "linear", swap_rb=True) for c in raw_bgr])
    assert np.array_equal(rect_d[2], g["depth_u"]) and np.array_equal(rect_rgb[2], g["rgb_u"])

    def run(batch, rectify):
        vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=1 << 13)
        if rectify:
            vol.set_rectification(g["map1"], g["map2"], swap_rb=True)
        d, c = (raw_d, raw_bgr) if rectify else (rect_d, rect_rgb)
        if batch:
            vol.integrate_batch(d, c, K, Ts)
        else:
            for i in range(n):
                vol.integrate(d[i], c[i], K, Ts[i])
        out = sort_dump(vol.dump_blocks())
        vol.close()
        return out

    ref = run(False, False)
    assert len(ref["keys"]) > 50
    for batch in (False, True):
        got = run(batch, True)
        for k in ("keys", "vox"):
            assert np.array_equal(got[k], ref[k]), (batch, k)
    # removing the maps restores the plain path; a wrong image size is an argument error
    vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=1 << 13)
    vol.set_rectification(g["map1"][:-8], g["map2"][:-8])
    with pytest.raises(RuntimeError):
        vol.integrate(raw_d[0], raw_bgr[0], K, Ts[0])
    vol.set_rectification(None, None)
    vol.integrate(rect_d[0], rect_rgb[0], K, Ts[0])
    assert vol.num_blocks() > 0
    vol.close()


def test_device_block_export_import_round_trip():
    """b2v_export_blocks_device / b2v_import_blocks_device (the multi-GPU mesh gather's device path): a volume
    rebuilt from another volume's device-resident blocks has the same blocks and the same mesh."""
    import torch
    cfg = S.CONFIGS["T0"]
    a = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=4096)
    for i in range(3):
        d, c, T = S.render_frame(cfg, i)
        a.integrate(d, c, cfg.K, T)
    keys, vox = a.export_blocks_torch()
    assert keys.is_cuda and keys.shape == (a.num_blocks(), 4) and vox.shape == (a.num_blocks(), 5, 512)
    b = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=4096)
    perm = torch.randperm(keys.shape[0], device=keys.device)        # block order must not matter
    b.import_blocks_torch(keys[perm].contiguous(), vox[perm].contiguous())
    da, db = sort_dump(a.dump_blocks()), sort_dump(b.dump_blocks())
    assert np.array_equal(da["keys"], db["keys"]) and np.array_equal(da["vox"], db["vox"])
    ma, mb = a.extract_mesh(), b.extract_mesh()
    ca = oracle.canonical_mesh(ma.vertices, ma.vertex_colors, ma.edge_ids,
                               ma.triangles)
    cb = oracle.canonical_mesh(mb.vertices, mb.vertex_colors, mb.edge_ids,
                               mb.triangles)
    for n in ("edges", "triangles", "vertices", "colors"):
        assert np.array_equal(ca[n], cb[n]), n
    empty = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=64)
    k0, v0 = empty.export_blocks_torch()
    assert k0.shape[0] == 0 and v0.shape[0] == 0
    empty.import_blocks_torch(k0, v0)
    for v in (a, b, empty):
        v.close()


def test_sharded_fused_batches_equal_the_unsharded_volume():
    """4-way hash-sharded volumes fed through the fused batch path (what a rank of `bench.py --gpus 4` runs): the
    union of the shards must equal the unsharded volume and the oracle bit for bit."""
    cfg = S.CONFIGS["C1"]
    n = 19
    frames = [S.render_frame(cfg, i) for i in range(n)]
    D, Cc, T = (np.stack([f[k] for f in frames]) for k in range(3))
    full, orc = _pair(cfg, capacity=1 << 16)
    full.integrate_batch(D, Cc, cfg.K, T)
    for d, c, t in frames:
        orc.integrate(d, c, cfg.K, t)
    _assert_same_volume(full, orc)
    parts = []
    for r in range(4):
        s = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=1 << 15, shard_rank=r,
                           shard_count=4)
        s.integrate_batch(D, Cc, cfg.K, T)
        parts.append(s.dump_blocks())
        s.close()
    merged = sort_dump({k: np.concatenate([p[k] for p in parts]) for k in ("keys", "hashes", "vox")})
    ref = sort_dump(full.dump_blocks())
    for name in ("keys", "hashes", "vox"):
        assert np.array_equal(merged[name], ref[name]), name


def test_raw_uint16_depth_equals_host_converted_float_depth():
    """b2v_integrate_u16 / b2v_integrate_batch_u16: raw 16-bit depth widened on the GPU == the reference's host
    conversion depth.astype(float32) * depth_factor (volumetric_integrator_base.py:1008-1015) fed as float32,
    bit for bit, on the fused batch path, the frame-by-frame path and with device-resident input."""
    import torch
    cfg = S.CONFIGS["C1"]
    n = 19
    frames = [S.render_frame(cfg, i) for i in range(n)]
    D, Cc, T = (np.stack([f[k] for f in frames]) for k in range(3))
    raw = np.round(D * 5000.0).astype(np.uint16)                 # TUM convention: 5000 units per metre
    scale = 1.0 / 5000.0
    Df = raw.astype(np.float32) * np.float32(scale)              # what the reference computes on the host
    ref, _ = _pair(cfg, capacity=1 << 16)
    ref.integrate_batch(Df, Cc, cfg.K, T)
    want = sort_dump(ref.dump_blocks())

    def check(vol):
        got = sort_dump(vol.dump_blocks())
        for name in ("keys", "hashes", "vox"):
            assert np.array_equal(got[name], want[name]), name
        vol.close()

    a, _ = _pair(cfg, capacity=1 << 16)
    a.integrate_batch(raw, Cc, cfg.K, T, depth_scale=scale)      # fused groups, host staging
    check(a)
    b, _ = _pair(cfg, capacity=1 << 16)
    for i in range(n):                                           # frame by frame
        b.integrate(raw[i], Cc[i], cfg.K, T[i], depth_scale=scale)
    check(b)
    c, _ = _pair(cfg, capacity=1 << 16)
    c.set_fusion(False)
    c.integrate_batch(raw, Cc, cfg.K, T, depth_scale=scale)      # un-fused batch
    check(c)
    d, _ = _pair(cfg, capacity=1 << 16)
    raw_dev = torch.from_numpy(raw.view(np.int16)).cuda()        # device-resident raw depth (same 16 bits)
    col_dev = torch.from_numpy(Cc).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d.integrate_batch(raw_dev, col_dev, cfg.K, T, stream=st.cuda_stream, depth_scale=scale)
    torch.cuda.synchronize()
    check(d)
    with pytest.raises(RuntimeError):
        ref.integrate(Df[0], Cc[0], cfg.K, T[0], depth_scale=scale)   # a scale goes with uint16 input only
    with pytest.raises(RuntimeError):
        ref.integrate_batch(raw, Cc, cfg.K, T, depth_scale=0.0)       # non-positive scale
    ref.close()


def test_fast_division_is_ieee():
    """The update kernels divide with a shared correctly rounded reciprocal + two residual corrections instead of the
    compiler's div.rn expansion.  All 2^23 significands of the reciprocal and 2^30 operand pairs of the quotient
    (depth-like and integer-weight denominators, plus random exponents) must equal __frcp_rn / __fdiv_rn bit for bit."""
    import ctypes as C
    from pyslam_b200 import _lib
    L = _lib.load()
    bad_r, bad_q = C.c_uint64(1), C.c_uint64(1)
    assert L.b2v_selftest_division(0, 1 << 30, C.byref(bad_r), C.byref(bad_q)) == 0
    assert bad_r.value == 0 and bad_q.value == 0, (bad_r.value, bad_q.value)


def test_frame_ingest_single_gpu_equals_integrate_batch():
    """FrameIngest without a process group: chunked, buffered uploads on a side stream + device-pointer batches;
    float32 and raw uint16 depth.  Same volume as one integrate_batch call."""
    import torch
    from pyslam_b200.sharding import FrameIngest
    cfg = S.CONFIGS["C1"]
    n = 21
    frames = [S.render_frame(cfg, i) for i in range(n)]
    D, Cc, T = (np.stack([f[k] for f in frames]) for k in range(3))
    ref, _ = _pair(cfg)
    ref.integrate_batch(D, Cc, cfg.K, T)
    vol, _ = _pair(cfg)
    ing = FrameIngest(vol, chunk_frames=8, buffers=2)
    ing.integrate_batch(torch.from_numpy(D).pin_memory(), torch.from_numpy(Cc).pin_memory(), cfg.K, T)
    ing.synchronize()
    a, b = sort_dump(ref.dump_blocks()), sort_dump(vol.dump_blocks())
    for name in ("keys", "hashes", "vox"):
        assert np.array_equal(a[name], b[name]), name
    assert ing.h2d_bytes == n * cfg.height * cfg.width * 7 and ing.gather_bytes == 0
    raw = np.round(D * 5000.0).astype(np.uint16)
    ref16, _ = _pair(cfg)
    ref16.integrate_batch(raw, Cc, cfg.K, T, depth_scale=np.float32(1 / 5000.0))
    vol16, _ = _pair(cfg)
    ing16 = FrameIngest(vol16, chunk_frames=8)
    ing16.integrate_batch(raw, Cc, cfg.K, T, depth_scale=np.float32(1 / 5000.0))   # pageable numpy input
    ing16.synchronize()
    a, b = sort_dump(ref16.dump_blocks()), sort_dump(vol16.dump_blocks())
    for name in ("keys", "hashes", "vox"):
        assert np.array_equal(a[name], b[name]), name


def test_degenerate_pose_takes_the_exact_division_path():
    """A (non-rigid) world->camera matrix whose depth row is ~1e-33 puts every voxel within 2^-100 of the camera
    plane: the update kernels leave their division fast path for __fdiv_rn.  Result == twin (which always divides
    exactly), and a following regular frame is unaffected."""
    cfg = S.CONFIGS["T0"]
    vol, orc = _pair(cfg)
    d, c, T = S.render_frame(cfg, 0)
    Tdeg = T.copy()
    Tdeg[2, :3] = 0.0
    Tdeg[2, 3] = 1.0e-33
    for pose in (T, Tdeg, S.render_frame(cfg, 1)[2]):
        vol.integrate(d, c, cfg.K, pose)
        orc.integrate(d, c, cfg.K, pose)
    _assert_same_volume(vol, orc)
