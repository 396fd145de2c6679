"""Multi-GPU partitioning of the voxel-block hash space.

One process per GPU.  A block belongs to rank `BlockKeyHash(key) % world` — the reference's own hash
(`cpp/volumetric/voxel_hashing.h:106-113`), so ownership is reproducible from the keys alone.  The
integrate path needs no collective: every rank sees every frame and keeps the keys it owns
(`b2v_config.shard_rank / shard_count`).  Mesh extraction needs the +1-voxel halos of blocks that
may live on another rank; `gather_blocks_device` collects all shards on one rank GPU-to-GPU over NCCL
(`gather_blocks` is the host-array variant used with gloo in the CPU tests), which then meshes the union.
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
    scratch = B200TsdfVolume(volume.voxel_length, volume.sdf_trunc, volume.depth_trunc,
                             capacity_blocks=cap, device=volume.device,
                             volume_unit_resolution=volume.volume_unit_resolution)
    if on_device:
        scratch.import_blocks_torch(keys, vox)
    else:
        scratch.upload_blocks(keys, vox)
    mesh = scratch.extract_mesh()
    scratch.close()
    return mesh
