"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path.

The sharded integrate needs no collective; what runs across ranks is ownership (hash % N), the
max-over-ranks timing reduction of bench.py, and the gather of shards for meshing."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from pyslam_b200 import sharding
from pyslam_b200 import synthetic as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank sees every frame (as in bench.py) and keeps only the blocks it owns
        cfg = S.CONFIGS["T0"]
        orc = oracle.TsdfOracle(cfg.voxel_size, cfg.sdf_trunc, cfg.depth_trunc)
        for i in range(2):
            d, c, T = S.render_frame(cfg, i)
            orc.integrate(d, c, cfg.K, T)
        full = orc.dump_blocks()
        mine = sharding.owner_of(full["keys"], world) == rank
        keys, vox = full["keys"][mine], full["vox"][mine]
        # bench.py's timing reduction: max over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == 10.0 + world - 1
        gk, gv = sharding.gather_blocks(keys, vox, dst=0)
        if rank == 0:
            order_a = np.lexsort((gk[:, 2], gk[:, 1], gk[:, 0]))
            order_b = np.lexsort((full["keys"][:, 2], full["keys"][:, 1], full["keys"][:, 0]))
            ok = np.array_equal(gk[order_a], full["keys"][order_b]) and \
                np.array_equal(gv[order_a], full["vox"][order_b])
            q.put(("ok" if ok else "mismatch", int(mine.sum()), len(full["keys"])))
        else:
            assert gk is None
            q.put(("worker", int(mine.sum()), len(full["keys"])))
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_and_gather_over_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert any(r[0] == "ok" for r in res) and not any(r[0] == "mismatch" for r in res)
    owned = sum(r[1] for r in res)
    assert owned == res[0][2]                      # the shards partition the block set
    assert all(0.3 < r[1] / r[2] < 0.7 for r in res)   # and are balanced


def test_hash_and_owner_match_the_reference_hash():
    rng = np.random.default_rng(3)
    keys = rng.integers(-5000, 5000, (200, 3)).astype(np.int32)
    h = sharding.block_key_hash(keys)
    L = oracle.oracle._tsdf()
    for k, hv in zip(keys, h):
        assert int(hv) == L.tsdf_oracle_block_key_hash(int(k[0]), int(k[1]), int(k[2]))
    if oracle.have_ref():
        for k, hv in zip(keys[:50], h[:50]):
            assert int(hv) == oracle.ref_block_key_hash(*k)
    own = sharding.owner_of(keys, 8)
    assert own.min() >= 0 and own.max() <= 7 and len(np.unique(own)) == 8
    merged = sharding.merge_dumps([
        dict(keys=keys[:100], hashes=h[:100], vox=np.zeros((100, 5, 512), np.float32)),
        dict(keys=keys[100:], hashes=h[100:], vox=np.ones((100, 5, 512), np.float32))])
    assert len(merged["keys"]) == 200 and np.array_equal(sharding.block_key_hash(merged["keys"]), merged["hashes"])


class _RecordingVolume:
    """Stand-in for B200TsdfVolume: records what FrameIngest hands to integrate_batch."""
    device = 0

    def __init__(self):
        self.calls = []

    def integrate_batch(self, depths, colors, K, poses, stream=None, depth_scale=None):
        self.calls.append((depths.clone(), colors.clone(), np.array(poses), depth_scale))

    def synchronize(self):
        pass


def _ingest_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)   # same batch on every rank
        n, H, W = 23, 6, 8               # ragged: 23 = 2 chunks of 10 + one of 3 (3 frames over 2 ranks: 2 + 1)
        D = rng.random((n, H, W)).astype(np.float32)
        Cc = rng.integers(0, 255, (n, H, W, 3), dtype=np.uint8)
        T = rng.random((n, 4, 4))
        vol = _RecordingVolume()
        ing = sharding.FrameIngest(vol, chunk_frames=10, buffers=2, device="cpu")
        ing.integrate_batch(D, Cc, (1.0, 1.0, 0.0, 0.0), T)
        ing.integrate_batch(np.round(D * 1000).astype(np.uint16), Cc, (1.0, 1.0, 0.0, 0.0), T, depth_scale=1e-3)
        got_d = np.concatenate([c[0].numpy() for c in vol.calls[:3]])
        got_c = np.concatenate([c[1].numpy() for c in vol.calls[:3]])
        got_t = np.concatenate([c[2] for c in vol.calls[:3]])
        got16 = np.concatenate([c[0].numpy() for c in vol.calls[3:]])
        ok = (np.array_equal(got_d, D) and np.array_equal(got_c, Cc) and np.array_equal(got_t, T)
              and [len(c[0]) for c in vol.calls] == [10, 10, 3, 10, 10, 3]
              and np.array_equal(got16, np.round(D * 1000).astype(np.uint16)) and vol.calls[3][3] == 1e-3)
        # every frame was uploaded by exactly one rank
        t = torch.tensor([float(ing.h2d_bytes)], dtype=torch.float64)
        dist.all_reduce(t)
        ok = ok and int(t.item()) == n * H * W * (4 + 3) + n * H * W * (2 + 3)
        q.put("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


def test_frame_ingest_splits_uploads_and_gathers_frames_in_order():
    assert sharding.chunk_plan(23, 2, 10) == [(0, 10, 5), (10, 10, 5), (20, 3, 2)]
    assert sharding.chunk_plan(8, 8, 64) == [(0, 8, 1)]
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ingest_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert res == ["ok", "ok"]
