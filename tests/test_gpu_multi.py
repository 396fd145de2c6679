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
