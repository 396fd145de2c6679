"""Multi-GPU partitioning of the voxel-block hash space.

One process per GPU.  A block belongs to rank `BlockKeyHash(key) % world` — the reference's own hash
(`cpp/volumetric/voxel_hashing.h:106-113`), so ownership is reproducible from the keys alone.  Every rank
integrates every frame into the blocks it owns (`b2v_config.shard_rank / shard_count`).

* Ingest (`FrameIngest`): a frame crosses PCIe ONCE in the whole job - rank r uploads 1/world of every chunk of
  frames over its own link and the chunk is completed GPU <-> GPU by an NCCL all-gather over NVLink / NVSwitch on a
  side stream, overlapped with the kernels of the previous chunk.
* Mesh extraction needs the +1-voxel halos of blocks that may live on another rank; `gather_blocks_device` collects
  all shards on one rank GPU-to-GPU over NCCL (`gather_blocks` is the host-array variant used with gloo in the CPU
  tests), which then meshes the union.
"""

from __future__ import annotations

import numpy as np


def block_key_hash(keys) -> np.ndarray:
    """`BlockKeyHash` of int32 keys [n,3] as uint64 (sign-extending, like libstdc++'s identity hash)."""
    k = np.asarray(keys, dtype=np.int32).reshape(-1, 3).astype(np.int64).astype(np.uint64)
    return k[:, 0] ^ (k[:, 1] << np.uint64(1)) ^ (k[:, 2] << np.uint64(2))


def owner_of(keys, world: int) -> np.ndarray:
    """Rank owning each block key."""
    return (block_key_hash(keys) % np.uint64(world)).astype(np.int64)


def merge_dumps(dumps):
    """Union of per-rank block dumps (dicts with keys/hashes/vox), sorted by key."""
    keys = np.concatenate([d["keys"] for d in dumps])
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return {name: np.concatenate([d[name] for d in dumps])[order] for name in dumps[0]}


def chunk_plan(n_frames: int, world: int, chunk_frames: int):
    """Split a batch into chunks of <= chunk_frames frames.  -> [(first frame, frame count, q)]: rank r uploads the
    frames [first + r*q, first + min((r+1)*q, count)) of the chunk (q = ceil(count / world)) into segment r of the
    chunk buffer, so after the all-gather of the world q-frame segments the chunk's frames are contiguous and in
    frame order (the last segments of a ragged chunk may be partly or wholly padding)."""
    plan = []
    c0 = 0
    while c0 < n_frames:
        cnt = min(chunk_frames, n_frames - c0)
        plan.append((c0, cnt, -(-cnt // world)))
        c0 += cnt
    return plan


class FrameIngest:
    """Frame-split ingest of a hash-sharded volume: `integrate_batch(depths, colors, K, poses)` with HOST frames
    (pinned numpy arrays or torch tensors; every rank passes the same batch) is, per chunk of `chunk_frames`:

        upload stream   H2D of this rank's 1/world share of the chunk (its own PCIe link)
        gather stream   all-gather of the shares (NCCL over NVLink; in place in the chunk buffer)
        compute stream  volume.integrate_batch(chunk buffer, device pointers)   [allocate + fused update kernels]

    with `buffers` chunk buffers in rotation: the upload of chunk c+2, the all-gather of chunk c+1 and the kernels of
    chunk c run concurrently (three streams, events between them).
    Frame order is unchanged, so the result is bit-identical to a single-GPU `integrate_batch` of the batch.
    world = 1 (or no process group) degenerates to a chunked, double-buffered upload.  `depth_scale`: the depths are
    raw uint16 (2 bytes per pixel over PCIe AND NVLink), widened on the GPU."""

    def __init__(self, volume, group=None, chunk_frames: int = 64, buffers: int = 4, device=None):
        import torch
        import torch.distributed as dist
        self.volume = volume
        self.group = group
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.chunk_frames = int(chunk_frames)
        self.device = torch.device(device) if device is not None else torch.device("cuda", volume.device)
        self.cuda = self.device.type == "cuda"
        self.n_buffers = int(buffers)
        self._bufs = None
        self._shape = None
        if self.cuda:
            self.s_up = torch.cuda.Stream(self.device)
            self.s_gather = torch.cuda.Stream(self.device)
            self.s_int = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0       # bytes this rank uploaded (accounting for bench.py)
        self.gather_bytes = 0    # bytes this rank received from its peers

    def _ensure(self, H, W, ddtype):
        import torch
        shape = (H, W, ddtype)
        if self._shape == shape:
            return
        if self.cuda and self._bufs is not None:
            torch.cuda.synchronize(self.device)
        cap = -(-self.chunk_frames // self.world) * self.world   # room for world equal segments
        self._bufs = []
        for _ in range(self.n_buffers):
            b = dict(depth=torch.empty((cap, H, W), dtype=ddtype, device=self.device),
                     color=torch.empty((cap, H, W, 3), dtype=torch.uint8, device=self.device))
            if self.cuda:
                b["free"] = torch.cuda.Event()
                b["uploaded"] = torch.cuda.Event()
                b["ready"] = torch.cuda.Event()
            self._bufs.append(b)
        self._shape = shape

    def integrate_batch(self, depths, colors, K, poses, depth_scale=None):
        import contextlib
        import torch
        D = depths if torch.is_tensor(depths) else torch.from_numpy(np.ascontiguousarray(depths))
        Cc = colors if torch.is_tensor(colors) else torch.from_numpy(np.ascontiguousarray(colors))
        T = np.ascontiguousarray(np.asarray(poses, np.float64).reshape(-1, 4, 4))
        n, H, W = int(D.shape[0]), int(D.shape[1]), int(D.shape[2])
        if tuple(Cc.shape) != (n, H, W, 3) or T.shape[0] != n:
            raise RuntimeError("depths must be [n,H,W], colors [n,H,W,3], poses [n,4,4]")
        if depth_scale is None and D.dtype != torch.float32:
            raise RuntimeError("depths must be float32 (or uint16 with depth_scale)")
        self._ensure(H, W, D.dtype)
        world, r = self.world, self.rank
        up = torch.cuda.stream(self.s_up) if self.cuda else contextlib.nullcontext()
        ga = torch.cuda.stream(self.s_gather) if self.cuda else contextlib.nullcontext()
        for i, (c0, cnt, q) in enumerate(chunk_plan(n, world, self.chunk_frames)):
            b = self._bufs[i % self.n_buffers]
            lo, hi = min(r * q, cnt), min((r + 1) * q, cnt)
            with up:
                if self.cuda:
                    self.s_up.wait_event(b["free"])      # the kernels that read this buffer last are done
                if hi > lo:
                    b["depth"][lo:hi].copy_(D[c0 + lo:c0 + hi], non_blocking=True)
                    b["color"][lo:hi].copy_(Cc[c0 + lo:c0 + hi], non_blocking=True)
                    self.h2d_bytes += (hi - lo) * H * W * (D.element_size() + 3)
                if self.cuda:
                    b["uploaded"].record(self.s_up)
            with ga:
                if self.cuda:
                    self.s_gather.wait_event(b["uploaded"])
                if world > 1:
                    # in place: segment r of the buffer is this rank's contribution
                    for t in (b["depth"], b["color"]):
                        if t.dtype == torch.uint16:      # no uint16 in NCCL / gloo: gather the same bytes as uint8
                            t = t.view(torch.uint8)
                        self.dist.all_gather_into_tensor(t[:world * q], t[r * q:(r + 1) * q], group=self.group)
                    self.gather_bytes += (cnt - (hi - lo)) * H * W * (D.element_size() + 3)
                if self.cuda:
                    b["ready"].record(self.s_gather)
            if self.cuda:
                if hasattr(self.volume, "set_input_event"):
                    # the allocate kernels of this chunk wait for the upload / all-gather only, not for the update
                    # kernels of the previous chunk that are queued on the compute stream
                    self.volume.set_input_event(b["ready"].cuda_event)
                else:
                    self.s_int.wait_event(b["ready"])
                self.volume.integrate_batch(b["depth"][:cnt], b["color"][:cnt], K, T[c0:c0 + cnt],
                                            stream=self.s_int.cuda_stream, depth_scale=depth_scale)
                b["free"].record(self.s_int)
            else:   # CPU / gloo (host-logic tests with a stand-in volume)
                self.volume.integrate_batch(b["depth"][:cnt], b["color"][:cnt], K, T[c0:c0 + cnt],
                                            depth_scale=depth_scale)

    def synchronize(self):
        import torch
        if self.cuda:
            self.s_up.synchronize()
            self.s_gather.synchronize()
            self.s_int.synchronize()
        self.volume.synchronize()


def gather_blocks(keys, vox, dst: int = 0, group=None, device=None):
    """Gather every rank's blocks on `dst` with torch.distributed (padded all_gather of sizes, then
    gather of the payloads).  keys int32 [n,3], vox float32 [n,5,512] numpy arrays.
    Returns (keys, vox) on dst, (None, None) elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    n = torch.tensor([len(keys)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(max(sizes), 1)
    k = torch.zeros((nmax, 3), dtype=torch.int32, device=dev)
    v = torch.zeros((nmax,) + tuple(np.asarray(vox).shape[1:]), dtype=torch.float32, device=dev)
    if len(keys):
        k[:len(keys)] = torch.from_numpy(np.ascontiguousarray(keys, np.int32)).to(dev)
        v[:len(keys)] = torch.from_numpy(np.ascontiguousarray(vox, np.float32)).to(dev)
    if rank == dst:
        ks = [torch.zeros_like(k) for _ in range(world)]
        vs = [torch.zeros_like(v) for _ in range(world)]
    else:
        ks = vs = None
    dist.gather(k, ks, dst=dst, group=group)
    dist.gather(v, vs, dst=dst, group=group)
    if rank != dst:
        return None, None
    out_k = np.concatenate([ks[r][:sizes[r]].cpu().numpy() for r in range(world)])
    out_v = np.concatenate([vs[r][:sizes[r]].cpu().numpy() for r in range(world)])
    return out_k, out_v


def gather_blocks_device(volume, dst: int = 0, group=None):
    """Device-resident gather over NCCL: every rank exports its blocks device-to-device
    (`b2v_export_blocks_device`), the payloads travel GPU to GPU (NVLink), nothing touches host memory.
    Returns (keys int32 [n,4], vox float32 [n,5,512]) CUDA tensors on dst, (None, None) elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    keys, vox = volume.export_blocks_torch()
    n = torch.tensor([keys.shape[0]], dtype=torch.int64, device=keys.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(max(sizes), 1)
    k = torch.zeros((nmax, 4), dtype=torch.int32, device=keys.device)
    v = torch.zeros((nmax,) + tuple(vox.shape[1:]), dtype=torch.float32, device=keys.device)
    k[:keys.shape[0]] = keys
    v[:vox.shape[0]] = vox
    ks = [torch.zeros_like(k) for _ in range(world)] if rank == dst else None
    vs = [torch.zeros_like(v) for _ in range(world)] if rank == dst else None
    dist.gather(k, ks, dst=dst, group=group)
    dist.gather(v, vs, dst=dst, group=group)
    if rank != dst:
        return None, None
    return (torch.cat([ks[r][:sizes[r]] for r in range(world)]).contiguous(),
            torch.cat([vs[r][:sizes[r]] for r in range(world)]).contiguous())


def extract_mesh_distributed(volume, dst: int = 0, group=None, device=None, capacity_blocks=None):
    """Mesh of a sharded volume: gather all shards on `dst`, load them into a scratch single-GPU
    volume there and run the marching-cubes kernels on the union.  Returns a TriangleMesh on dst,
    None elsewhere.  With the NCCL backend the blocks stay on the GPUs (`gather_blocks_device`); any other
    backend (gloo in the CPU tests) goes through host arrays (`gather_blocks`)."""
    import torch.distributed as dist
    from .volume import B200TsdfVolume

    on_device = dist.get_backend(group) == "nccl"
    if on_device:
        keys, vox = gather_blocks_device(volume, dst=dst, group=group)
    else:
        d = volume.dump_blocks()
        keys, vox = gather_blocks(d["keys"], d["vox"], dst=dst, group=group, device=device)
    if dist.get_rank(group) != dst:
        return None
    cap = capacity_blocks or max(2 * len(keys), 1024)
    scratch = getattr(volume, "_mesh_scratch", None)   # kept between extractions (one per output tick)
    if scratch is None or scratch.capacity_blocks < len(keys) + 1:
        if scratch is not None:
            scratch.close()
        scratch = B200TsdfVolume(volume.voxel_length, volume.sdf_trunc, volume.depth_trunc,
                                 capacity_blocks=cap, device=volume.device,
                                 volume_unit_resolution=volume.volume_unit_resolution)
        volume._mesh_scratch = scratch
    else:
        scratch.reset()
    if on_device:
        scratch.import_blocks_torch(keys, vox)
    else:
        scratch.upload_blocks(keys, vox)
    return scratch.extract_mesh()
