"""GPU (needs >= 2 devices; skipped otherwise): one process per GPU over NCCL.  Each rank integrates
every frame into its hash-bucket shard; rank 0 gathers the shards and meshes the union; volume and
mesh must equal the single-volume oracle bit for bit."""

import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import oracle
    from pyslam_b200 import B200TsdfVolume, sharding
    from pyslam_b200 import synthetic as S

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = S.CONFIGS["T0"]
        vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=4096,
                             device=rank, shard_rank=rank, shard_count=world)
        frames = [S.render_frame(cfg, i) for i in range(3)]
        for d, c, T in frames:
            vol.integrate(d, c, cfg.K, T)
        dump = vol.dump_blocks()
        assert np.all(sharding.owner_of(dump["keys"], world) == rank)
        mesh = sharding.extract_mesh_distributed(vol, dst=0, device=f"cuda:{rank}")
        gk, gv = sharding.gather_blocks(dump["keys"], dump["vox"], dst=0, device=f"cuda:{rank}")
        if rank == 0:
            orc = oracle.TsdfOracle(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc)
            for d, c, T in frames:
                orc.integrate(d, c, cfg.K, T)
            ref = orc.dump_blocks()
            oa = np.lexsort((gk[:, 2], gk[:, 1], gk[:, 0]))
            ob = np.lexsort((ref["keys"][:, 2], ref["keys"][:, 1], ref["keys"][:, 0]))
            ok = np.array_equal(gk[oa], ref["keys"][ob]) and np.array_equal(gv[oa], ref["vox"][ob])
            rm = orc.extract_mesh()
            a = oracle.canonical_mesh(mesh.vertices, mesh.vertex_colors,
                                      mesh.edge_ids, mesh.triangles)
            b = oracle.canonical_mesh(rm["vertices"], rm["colors"], rm["edges"], rm["triangles"])
            ok = ok and all(np.array_equal(a[n], b[n]) for n in ("edges", "triangles", "vertices", "colors"))
            q.put("ok" if ok else "mismatch")
        else:
            assert mesh is None
            q.put("worker")
    finally:
        dist.destroy_process_group()


def test_two_gpu_shards_and_distributed_mesh():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert "ok" in res and "mismatch" not in res


def _ingest_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import oracle
    from pyslam_b200 import B200TsdfVolume, sharding
    from pyslam_b200 import synthetic as S

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cfg = S.CONFIGS["C1"]
        n = 21                                             # 2 chunks of 8 + a ragged one of 5 (3 + 2 frames)
        frames = [S.render_frame(cfg, i) for i in range(n)]
        D = torch.from_numpy(np.stack([f[0] for f in frames])).pin_memory()
        Cc = torch.from_numpy(np.stack([f[1] for f in frames])).pin_memory()
        T = np.stack([f[2] for f in frames])
        vol = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=1 << 15,
                             device=rank, shard_rank=rank, shard_count=world)
        ing = sharding.FrameIngest(vol, chunk_frames=8, buffers=2)
        ing.integrate_batch(D, Cc, cfg.K, T)                                  # float32 depth
        raw = torch.from_numpy(np.round(D.numpy() * 5000.0).astype(np.uint16)).pin_memory()
        vol16 = B200TsdfVolume(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc, capacity_blocks=1 << 15,
                               device=rank, shard_rank=rank, shard_count=world)
        ing16 = sharding.FrameIngest(vol16, chunk_frames=8, buffers=2)
        ing16.integrate_batch(raw, Cc, cfg.K, T, depth_scale=np.float32(1.0 / 5000.0))   # raw 16-bit depth
        ing.synchronize()
        ing16.synchronize()
        up = torch.tensor([float(ing.h2d_bytes)], dtype=torch.float64, device="cuda")
        dist.all_reduce(up)
        dump, dump16 = vol.dump_blocks(), vol16.dump_blocks()
        gk, gv = sharding.gather_blocks(dump["keys"], dump["vox"], dst=0, device=f"cuda:{rank}")
        gk16, gv16 = sharding.gather_blocks(dump16["keys"], dump16["vox"], dst=0, device=f"cuda:{rank}")
        if rank == 0:
            ok = int(up.item()) == n * cfg.height * cfg.width * 7          # every frame crossed PCIe exactly once
            for keys, vox, scale in ((gk, gv, None), (gk16, gv16, np.float32(1.0 / 5000.0))):
                orc = oracle.TsdfOracle(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc)
                for i, (d, c, t) in enumerate(frames):
                    dd = d if scale is None else raw[i].numpy().astype(np.float32) * scale
                    orc.integrate(dd, c, cfg.K, t)
                ref = orc.dump_blocks()
                oa = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
                ob = np.lexsort((ref["keys"][:, 2], ref["keys"][:, 1], ref["keys"][:, 0]))
                ok = ok and np.array_equal(keys[oa], ref["keys"][ob]) and np.array_equal(vox[oa], ref["vox"][ob])
            q.put("ok" if ok else "mismatch")
        else:
            q.put("worker")
    finally:
        dist.destroy_process_group()


def test_two_gpu_frame_split_ingest_equals_the_oracle():
    """FrameIngest: rank r uploads half of every chunk, an NCCL all-gather completes it on both GPUs, each rank
    integrates its hash shard; the union equals the single-volume oracle, and every frame crossed PCIe once."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ingest_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert "ok" in res and "mismatch" not in res
