"""Host-side mirror of the reference's volume objects, over the C ABI (include/b2v.h).

Two duck types are provided, the two `self.volume` shapes pySLAM's dense front-end drives
(SURVEY.md §8b):

* `B200TsdfVolume` — the north_star API `integrate(depth, color, K, pose)` / `extract_mesh()`
  plus the Open3D-style methods the TSDF backend calls
  (`pyslam/dense/volumetric_integrator_tsdf.py:156,223,239,246,260,267`):
  `integrate(rgbd, intrinsic, extrinsic)`, `extract_triangle_mesh()`, `extract_point_cloud()`,
  `reset()`.
* `VoxelBlockGrid` — pySLAM's own `volumetric.VoxelBlockGrid` surface
  (`cpp/volumetric/volumetric_grid_module.h:732-935`): `integrate(points, colors)`,
  `get_voxels(min_count)`, `get_points()`, `get_colors()`, `clear()`, `reset()`, `size()`,
  `empty()`, `num_blocks()`, `get_block_size()`, `get_total_voxel_count()`,
  `remove_low_count_voxels(n)`.

Everything numeric happens in the CUDA library; these classes only validate arguments (same
error behaviour as the reference: `RuntimeError` on shape / dtype mismatches,
`volumetric_grid_module.h:140-258`) and move pointers.  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import B2VConfig, BLOCK_SIZE, BLOCK_VOXELS, VOXEL_PLANES


def _as_K4(K) -> np.ndarray:
    """Accept (fx, fy, cx, cy), a 3x3 matrix, or an object with fx/fy/cx/cy (camera / o3d-like)."""
    if hasattr(K, "fx") and hasattr(K, "cx"):
        return np.array([K.fx, K.fy, K.cx, K.cy], dtype=np.float64)
    if hasattr(K, "intrinsic_matrix"):
        K = np.asarray(K.intrinsic_matrix)
    K = np.asarray(K, dtype=np.float64)
    if K.shape == (3, 3):
        return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)
    if K.size == 4:
        return np.ascontiguousarray(K.reshape(4))
    raise RuntimeError("K must be (fx, fy, cx, cy) or a 3x3 intrinsic matrix")


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and bool(x.is_cuda)


class TriangleMesh:
    """Arrays shaped like `VolumetricIntegrationMesh` (volumetric_integrator_base.py:213-226)."""

    def __init__(self, vertices, triangles, vertex_colors, edge_ids=None):
        self.vertices = vertices                 # [V,3] float64
        self.triangles = triangles               # [T,3] int32
        self.vertex_colors = vertex_colors       # [V,3] float64 in [0,1]
        self.vertex_normals = np.zeros((0, 3), dtype=np.float64)  # Open3D leaves them empty too
        self.edge_ids = edge_ids                 # [V,4] int32 canonical weld key (parity hook)


class PointCloud:
    """Arrays shaped like `VolumetricIntegrationPointCloud` (volumetric_integrator_base.py:159-210)."""

    def __init__(self, points, colors):
        self.points = points
        self.colors = colors


class B200TsdfVolume:
    """B200-native TSDF + colour volume on 8^3 voxel blocks in a GPU hash table.

    Parameters mirror `o3d.pipelines.integration.ScalableTSDFVolume(voxel_length, sdf_trunc, ...)`
    as constructed at volumetric_integrator_tsdf.py:104-108, plus the depth truncation the reference
    applies while building the RGBD image (tsdf.py:215-221).
    """

    def __init__(self, voxel_length: float, sdf_trunc: float, depth_trunc: float = 4.0,
                 capacity_blocks: int = 1 << 18, device: int = 0, depth_sampling_stride: int = 4,
                 block_size: int = BLOCK_SIZE, shard_rank: int = 0, shard_count: int = 1,
                 volume_unit_resolution: int = 16):
        """`volume_unit_resolution`: Open3D's parameter of that name.  16 (the reference's value) allocates every
        8^3 block of each 16^3 unit `ScalableTSDFVolume::LocateVolumeUnit` touches - the same voxels Open3D updates;
        8 is SURVEY decision D1 (the float32 pyslam key range of the +-sdf_trunc box, ~9 % fewer blocks)."""
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.voxel_length = float(voxel_length)
        self.sdf_trunc = float(sdf_trunc)
        self.depth_trunc = float(depth_trunc)
        self.block_size = int(block_size)
        self.capacity_blocks = int(capacity_blocks)
        self.device = int(device)
        self.shard_rank, self.shard_count = int(shard_rank), int(shard_count)
        self.volume_unit_resolution = int(volume_unit_resolution)
        cfg = B2VConfig(voxel_length, block_size, sdf_trunc, depth_trunc, depth_sampling_stride,
                        capacity_blocks, device, shard_rank, shard_count, int(volume_unit_resolution),
                        float(voxel_length), float(sdf_trunc))
        rc = self._L.b2v_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.B2V_OK:
            msg = self._L.b2v_last_error(self._h).decode() if self._h else "invalid configuration"
            if self._h:
                self._L.b2v_destroy(self._h)
                self._h = C.c_void_p()
            raise RuntimeError(f"b2v_create failed (status {rc}): {msg}")
        self._keepalive = []  # host arrays of frames still in flight

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None):
            self._L.b2v_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != _lib.B2V_OK:
            raise RuntimeError(f"{what} failed (status {rc}): {self._L.b2v_last_error(self._h).decode()}")

    # ---- integrate ----
    def integrate(self, depth, color=None, K=None, pose=None, stream=None, depth_scale=None):
        """north_star: `integrate(depth, color, K, pose)` with depth float32 [H,W] metres, colour
        uint8 RGB [H,W,3], K = (fx,fy,cx,cy) | 3x3, pose = Tcw 4x4 float64.
        Open3D style (tsdf.py:223): `integrate(rgbd, intrinsic, extrinsic)` where `rgbd` has
        `.color` / `.depth`.  Inputs may be numpy arrays or CUDA torch tensors.  Asynchronous.
        `stream`: optional cudaStream_t handle (int) to launch on; device inputs only.
        `depth_scale`: given with a RAW uint16 depth image (numpy) - it is uploaded as 16-bit and widened on the
        GPU to float32(depth) * float32(depth_scale), like `depth.astype(np.float32) * depth_factor`."""
        if hasattr(depth, "depth") and hasattr(depth, "color"):
            rgbd, K, pose = depth, color, K
            depth, color = rgbd.depth, rgbd.color
        if K is None or pose is None or color is None:
            raise RuntimeError("integrate(depth, color, K, pose): missing argument")
        K4 = _as_K4(K)
        T = np.ascontiguousarray(np.asarray(pose, dtype=np.float64).reshape(4, 4)).reshape(16)
        if _is_torch_cuda(depth):
            if not _is_torch_cuda(color):
                raise RuntimeError("depth and color must both be CUDA tensors or both host arrays")
            if str(depth.dtype) != "torch.float32" or str(color.dtype) != "torch.uint8":
                raise RuntimeError("depth must be float32 and color uint8")
            if not depth.is_contiguous() or not color.is_contiguous():
                raise RuntimeError("depth and color must be contiguous")
            H, W = int(depth.shape[0]), int(depth.shape[1])
            if depth.dim() != 2 or tuple(color.shape) != (H, W, 3):
                raise RuntimeError("depth must be [H,W] and color [H,W,3]")
            dp, cp = depth.data_ptr(), color.data_ptr()
            self._keepalive.append((depth, color))
        else:
            d = np.asarray(depth)
            c = np.asarray(color)
            if d.ndim != 2:
                raise RuntimeError("depth must have 2 dimensions [H,W]")
            if c.ndim != 3 or c.shape[2] != 3 or c.shape[:2] != d.shape:
                raise RuntimeError("color must be [H,W,3] with the depth image's size")
            if c.dtype != np.uint8:
                raise RuntimeError("color must be uint8 RGB")
            raw16 = depth_scale is not None and d.dtype == np.uint16
            if depth_scale is not None and not raw16:
                raise RuntimeError("depth_scale goes with a uint16 depth image")
            # reference: depth.astype(float32), base.py:1008-1017 (raw uint16 is widened on the GPU instead)
            d = np.ascontiguousarray(d) if raw16 else np.ascontiguousarray(d, dtype=np.float32)
            c = np.ascontiguousarray(c)
            H, W = d.shape
            dp, cp = d.ctypes.data, c.ctypes.data
            self._keepalive.append((d, c))
            if raw16:
                rc = self._L.b2v_integrate_u16(self._h, dp, float(depth_scale), cp, H, W, K4.ctypes.data,
                                               T.ctypes.data, None)
                self._check(rc, "b2v_integrate_u16")
                if len(self._keepalive) > 8:
                    del self._keepalive[:-8]
                return
        if depth_scale is not None:
            raise RuntimeError("depth_scale is supported for host (numpy) uint16 depth images")
        rc = self._L.b2v_integrate(self._h, dp, cp, H, W, K4.ctypes.data, T.ctypes.data,
                                   C.c_void_p(stream) if stream else None)
        self._check(rc, "b2v_integrate")
        if len(self._keepalive) > 8:
            # staging ring is 4 deep: anything older has been consumed by the copy engine
            del self._keepalive[:-8]

    def integrate_batch(self, depths, colors, K, poses, stream=None, depth_scale=None):
        """n frames back to back (the rebuild(map) bulk path, base.py:1242-1318): depths [n,H,W] f32,
        colors [n,H,W,3] u8, poses [n,4,4] Tcw.  One C call enqueues every frame.  With `depth_scale`, depths
        is RAW uint16 [n,H,W] (numpy, or a CUDA uint16 / int16 tensor) widened on the GPU."""
        K4 = _as_K4(K)
        T = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 16))
        n = T.shape[0]
        raw16 = depth_scale is not None
        if _is_torch_cuda(depths):
            if not (depths.is_contiguous() and colors.is_contiguous()):
                raise RuntimeError("depths and colors must be contiguous")
            if raw16 and depths.element_size() != 2:
                raise RuntimeError("depth_scale goes with 16-bit depth images")
            H, W = int(depths.shape[1]), int(depths.shape[2])
            dp, cp = depths.data_ptr(), colors.data_ptr()
            hold = (depths, colors)
        else:
            if raw16 and np.asarray(depths).dtype != np.uint16:
                raise RuntimeError("depth_scale goes with uint16 depth images")
            d = np.ascontiguousarray(depths) if raw16 else np.ascontiguousarray(depths, dtype=np.float32)
            c = np.ascontiguousarray(colors, dtype=np.uint8)
            if d.ndim != 3 or c.shape != d.shape + (3,) or d.shape[0] != n:
                raise RuntimeError("depths must be [n,H,W], colors [n,H,W,3], poses [n,4,4]")
            H, W = d.shape[1:]
            dp, cp = d.ctypes.data, c.ctypes.data
            hold = (d, c)
        if raw16:
            rc = self._L.b2v_integrate_batch_u16(self._h, n, dp, float(depth_scale), cp, H, W, K4.ctypes.data,
                                                 T.ctypes.data, C.c_void_p(stream) if stream else None)
        else:
            rc = self._L.b2v_integrate_batch(self._h, n, dp, cp, H, W, K4.ctypes.data, T.ctypes.data,
                                             C.c_void_p(stream) if stream else None)
        self._check(rc, "b2v_integrate_batch")
        self._keepalive = [hold]

    def synchronize(self):
        self._check(self._L.b2v_synchronize(self._h), "b2v_synchronize")
        self._keepalive.clear()

    def reset(self):
        """`self.volume.reset()` (tsdf.py:156; base.py:642)."""
        self._check(self._L.b2v_reset(self._h), "b2v_reset")
        self._keepalive.clear()

    # ---- inspection ----
    def num_blocks(self) -> int:
        n = self._L.b2v_num_blocks(self._h)
        if n < 0:
            raise RuntimeError(self._L.b2v_last_error(self._h).decode())
        return int(n)

    def last_mesh_stats(self) -> dict:
        """How the most recent mesh / point extraction narrowed its work (b2v_last_mesh_stats)."""
        st = (C.c_int64 * 5)()
        self._check(self._L.b2v_last_mesh_stats(self._h, st), "b2v_last_mesh_stats")
        return dict(zip(("blocks", "candidate_tiles", "tiles_with_both_signs", "vertex_blocks", "triangle_blocks"),
                        (int(x) for x in st)))

    def last_frame_stats(self):
        t, n = C.c_int64(0), C.c_int64(0)
        self._check(self._L.b2v_last_frame_stats(self._h, C.byref(t), C.byref(n)), "b2v_last_frame_stats")
        return int(t.value), int(n.value)

    def counters(self):
        """(total (block, frame) updates, kernel launches) since create/reset."""
        u, k, b = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self._L.b2v_counters(self._h, C.byref(u), C.byref(k), C.byref(b)), "b2v_counters")
        return int(u.value), int(k.value)

    def block_visits(self) -> int:
        """Blocks read + written since create/reset (== updates frame by frame; fewer when fused)."""
        u, k, b = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self._L.b2v_counters(self._h, C.byref(u), C.byref(k), C.byref(b)), "b2v_counters")
        return int(b.value)

    def set_fusion(self, enable: bool):
        """integrate_batch: fuse groups of frames per block visit (default, see set_group_size) or go frame by frame."""
        self._check(self._L.b2v_set_fusion(self._h, 1 if enable else 0), "b2v_set_fusion")

    def set_input_event(self, cuda_event):
        """The next integrate_batch call's device frames are ready when `cuda_event` (a raw cudaEvent_t handle, e.g.
        `torch.cuda.Event.cuda_event`) fires; see b2v_set_input_event."""
        self._check(self._L.b2v_set_input_event(self._h, C.c_void_p(int(cuda_event))), "b2v_set_input_event")

    def set_group_size(self, frames: int):
        """Frames per fused group of integrate_batch (1..32, default 16); results do not depend on it."""
        self._check(self._L.b2v_set_group_size(self._h, int(frames)), "b2v_set_group_size")

    def set_rectification(self, map_x, map_y, swap_rb: bool = False):
        """Install the undistortion maps of `cv2.initUndistortRectifyMap(K, D, None, new_K, (w, h), CV_32FC1)`
        (volumetric_integrator_base.py:766-778): integrate() then takes the RAW images and rectifies them on the
        GPU exactly like `cv2.remap` (colour bilinear, depth nearest; base.py:1034-1039).  `swap_rb=True` also
        converts BGR input to RGB (base.py:1054).  `None` maps remove the stage."""
        if map_x is None or map_y is None:
            self._check(self._L.b2v_set_rectification(self._h, None, None, 0, 0, 0), "b2v_set_rectification")
            return
        mx = np.ascontiguousarray(map_x, np.float32)
        my = np.ascontiguousarray(map_y, np.float32)
        if mx.ndim != 2 or mx.shape != my.shape:
            raise RuntimeError("map_x and map_y must be float32 [H,W] arrays of the same shape")
        self._check(self._L.b2v_set_rectification(self._h, mx.ctypes.data, my.ctypes.data, mx.shape[0], mx.shape[1],
                                                  1 if swap_rb else 0), "b2v_set_rectification")

    def set_overlap(self, enable: bool):
        """Run allocate(f+1) concurrently with integrate(f) (default) or serialise them."""
        self._check(self._L.b2v_set_overlap(self._h, 1 if enable else 0), "b2v_set_overlap")

    def profile_enable(self, enable: bool = True):
        self._check(self._L.b2v_profile_enable(self._h, 1 if enable else 0), "b2v_profile_enable")

    def profile_read(self):
        """(allocate_ms, integrate_ms, frames, integrate launches) summed since the last read."""
        a, b, n, l = C.c_double(0), C.c_double(0), C.c_int64(0), C.c_int64(0)
        self._check(self._L.b2v_profile_read(self._h, C.byref(a), C.byref(b), C.byref(n), C.byref(l)),
                    "b2v_profile_read")
        return a.value, b.value, int(n.value), int(l.value)

    def last_touched_keys(self) -> np.ndarray:
        n = self._L.b2v_last_touched_keys(self._h, None, 0)
        keys = np.zeros((max(int(n), 0), 3), np.int32)
        if n > 0:
            self._L.b2v_last_touched_keys(self._h, keys.ctypes.data, int(n))
        return keys

    def dump_blocks(self):
        """Parity hook: keys int32 [nb,3], hashes uint64 [nb] (reference BlockKeyHash),
        vox float32 [nb,5,512] planes (tsdf, weight, r, g, b)."""
        nb = self.num_blocks()
        keys = np.zeros((nb, 3), np.int32)
        hashes = np.zeros(nb, np.uint64)
        vox = np.zeros((nb, VOXEL_PLANES, BLOCK_VOXELS), np.float32)
        n = self._L.b2v_dump_blocks(self._h, keys.ctypes.data, hashes.ctypes.data, vox.ctypes.data)
        if n != nb:
            raise RuntimeError(f"b2v_dump_blocks returned {n}, expected {nb}")
        return dict(keys=keys, hashes=hashes, vox=vox)

    def upload_blocks(self, keys, vox):
        """Restore / seed blocks: keys int32 [n,3] (unique), vox float32 [n,5,512]."""
        k = np.ascontiguousarray(keys, np.int32).reshape(-1, 3)
        x = np.ascontiguousarray(vox, np.float32).reshape(k.shape[0], VOXEL_PLANES, BLOCK_VOXELS)
        self._check(self._L.b2v_upload_blocks(self._h, k.shape[0], k.ctypes.data, x.ctypes.data),
                    "b2v_upload_blocks")

    def export_blocks_torch(self):
        """(keys int32 [n,4] = {x,y,z,0}, vox float32 [n,5,512]) as torch CUDA tensors on this volume's device:
        device-to-device copies of the block keys and the block pool (multi-GPU mesh gather)."""
        import torch
        n = self._L.b2v_export_blocks_device(self._h, None, None, 0)
        if n < 0:
            raise RuntimeError(self._L.b2v_last_error(self._h).decode())
        dev = torch.device("cuda", self.device)
        keys = torch.empty((n, 4), dtype=torch.int32, device=dev)
        vox = torch.empty((n, VOXEL_PLANES, BLOCK_VOXELS), dtype=torch.float32, device=dev)
        if n:
            got = self._L.b2v_export_blocks_device(self._h, keys.data_ptr(), vox.data_ptr(), n)
            if got != n:
                raise RuntimeError(self._L.b2v_last_error(self._h).decode())
        return keys, vox

    def import_blocks_torch(self, keys, vox):
        """Inverse of export_blocks_torch: contiguous torch CUDA tensors on this volume's device."""
        if keys.shape[0] == 0:
            return
        if not (keys.is_cuda and vox.is_cuda and keys.is_contiguous() and vox.is_contiguous()):
            raise RuntimeError("keys and vox must be contiguous CUDA tensors")
        self._check(self._L.b2v_import_blocks_device(self._h, int(keys.shape[0]), keys.data_ptr(), vox.data_ptr()),
                    "b2v_import_blocks_device")

    # ---- outputs ----
    def extract_mesh(self) -> TriangleMesh:
        """north_star `extract_mesh()` == Open3D `extract_triangle_mesh()` (tsdf.py:239,260)."""
        nv, nt = C.c_int64(0), C.c_int64(0)
        self._check(self._L.b2v_extract_mesh(self._h, C.byref(nv), C.byref(nt)), "b2v_extract_mesh")
        V = np.zeros((nv.value, 3), np.float64)   # float64 like Open3D's TriangleMesh, computed in float64
        Cc = np.zeros((nv.value, 3), np.float64)
        E = np.zeros((nv.value, 4), np.int32)
        T = np.zeros((nt.value, 3), np.int32)
        self._check(self._L.b2v_copy_mesh(self._h, V.ctypes.data, Cc.ctypes.data, E.ctypes.data,
                                          T.ctypes.data), "b2v_copy_mesh")
        return TriangleMesh(V, T, Cc, E)

    extract_triangle_mesh = extract_mesh

    def extract_point_cloud(self) -> PointCloud:
        n = C.c_int64(0)
        self._check(self._L.b2v_extract_points(self._h, C.byref(n)), "b2v_extract_points")
        P = np.zeros((n.value, 3), np.float64)
        Cc = np.zeros((n.value, 3), np.float64)
        self._check(self._L.b2v_copy_points(self._h, P.ctypes.data, Cc.ctypes.data), "b2v_copy_points")
        return PointCloud(P, Cc)


def filter_shadow_points(depth, delta_depth=None, delta_x=2, delta_y=2, fill_value=-1, device=0):
    """GPU version of `pyslam.utilities.depth.filter_shadow_points` (depth.py:103-146) for the default
    `delta_depth=None` (median-based threshold).  depth float32 [H,W] -> filtered copy."""
    if delta_depth is not None:
        raise NotImplementedError("only the median-based threshold (delta_depth=None) is implemented")
    d = np.ascontiguousarray(depth, dtype=np.float32)
    if d.ndim != 2:
        raise RuntimeError("depth must be a 2D array")
    out = np.empty_like(d)
    rc = _lib.load().b2v_filter_shadow_points(d.ctypes.data, d.shape[0], d.shape[1], int(delta_x), int(delta_y),
                                             float(fill_value), out.ctypes.data, int(device))
    if rc != _lib.B2V_OK:
        raise RuntimeError(f"b2v_filter_shadow_points failed (status {rc})")
    return out


def remap(src, map_x, map_y, interpolation="linear", swap_rb=False, device=0):
    """GPU `cv2.remap(src, map_x, map_y, interpolation)` for the two cases the dense front-end uses
    (volumetric_integrator_base.py:1017-1047): uint8 [H,W,3] with INTER_LINEAR, and float32 / int32 [H,W] with
    INTER_NEAREST; constant zero border.  Bit-exact with OpenCV's fixed-point arithmetic."""
    a = np.ascontiguousarray(src)
    mx = np.ascontiguousarray(map_x, np.float32)
    my = np.ascontiguousarray(map_y, np.float32)
    if a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3 and interpolation == "linear":
        kind = 0
    elif a.dtype in (np.float32, np.int32) and a.ndim == 2 and interpolation == "nearest":
        kind = 1
    else:
        raise RuntimeError("remap supports uint8 [H,W,3] + 'linear' and float32/int32 [H,W] + 'nearest'")
    if mx.shape != a.shape[:2] or my.shape != a.shape[:2]:
        raise RuntimeError("maps must have the image's height and width")
    out = np.empty_like(a)
    rc = _lib.load().b2v_remap(a.ctypes.data, kind, a.shape[0], a.shape[1], mx.ctypes.data, my.ctypes.data,
                              out.ctypes.data, 1 if swap_rb else 0, int(device))
    if rc != _lib.B2V_OK:
        raise RuntimeError(f"b2v_remap failed (status {rc})")
    return out


class CameraFrustrum:
    """Mirror of `volumetric.CameraFrustrum(fx, fy, cx, cy, width, height, T_cw, depth_max, depth_min)`
    (cpp/volumetric/camera_frustrum.h:36-48): the arguments of carve / frustum queries."""

    def __init__(self, fx, fy, cx, cy, width, height, T_cw=None, depth_max=10.0, depth_min=1e-2):
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.width, self.height = int(width), int(height)
        self.depth_max, self.depth_min = float(depth_max), float(depth_min)
        self.T_cw = np.eye(4) if T_cw is None else np.asarray(T_cw, np.float64).reshape(4, 4)

    def set_T_cw(self, T_cw):
        self.T_cw = np.asarray(T_cw, np.float64).reshape(4, 4)

    def get_width(self):
        return self.width

    def get_height(self):
        return self.height

    def _args(self):
        K = np.array([self.fx, self.fy, self.cx, self.cy], np.float32)
        T = np.ascontiguousarray(self.T_cw, np.float64).reshape(16)
        return K, T


class BoundingBox3D:
    """Mirror of `volumetric.BoundingBox3D(min_x, min_y, min_z, max_x, max_y, max_z)`
    (cpp/volumetric/bounding_boxes_3d.h:40-80)."""

    def __init__(self, min_x, min_y, min_z, max_x, max_y, max_z):
        self.bounds = np.array([min_x, min_y, min_z, max_x, max_y, max_z], np.float64)


class TBBUtils:
    """`volumetric.TBBUtils` of the reference module (cpp/volumetric/volumetric_module.cpp:43-50): the integrators call
    `TBBUtils.set_max_threads(n)` to size the CPU thread pool of the voxel grids.  The B200 grids have no CPU pool; the
    value is kept so that `get_max_threads()` answers what was set."""
    _max_threads = 0

    @staticmethod
    def set_max_threads(num_threads: int) -> None:
        TBBUtils._max_threads = int(num_threads)

    @staticmethod
    def get_max_threads() -> int:
        import os
        return TBBUtils._max_threads if TBBUtils._max_threads > 0 else (os.cpu_count() or 1)


class VoxelGridData:
    """`VoxelGridDataT` (cpp/volumetric/voxel_grid_data.h:36-50): points / colors SoA."""

    def __init__(self, points, colors):
        self.points = points
        self.colors = colors
        self.class_ids = None
        self.object_ids = None
        self.confidences = None


class VoxelBlockGrid:
    """GPU drop-in for pySLAM's `volumetric.VoxelBlockGrid(voxel_size, block_size=8)`."""

    def __init__(self, voxel_size: float, block_size: int = 8, capacity_blocks: int = 1 << 17,
                 device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.voxel_size = float(voxel_size)
        self._block_size = int(block_size)
        rc = self._L.b2v_grid_create(voxel_size, block_size, capacity_blocks, device, C.byref(self._h))
        if rc != _lib.B2V_OK:
            msg = self._L.b2v_grid_last_error(self._h).decode() if self._h else "invalid configuration"
            if self._h:
                self._L.b2v_grid_destroy(self._h)
                self._h = C.c_void_p()
            raise RuntimeError(f"b2v_grid_create failed (status {rc}): {msg}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.b2v_grid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != _lib.B2V_OK:
            raise RuntimeError(f"{what} failed (status {rc}): {self._L.b2v_grid_last_error(self._h).decode()}")

    def integrate(self, points, colors=None, class_ids=None, instance_ids=None, depths=None):
        """integrate(points [N,3] f32|f64, colors [N,3] u8|f32 | None)
        (volumetric_grid_module.h:131-467).  float64 points take the reference's float64 overload (:737-749): voxel
        keys from the float64 coordinates, sums accumulate float32(x); uint8 colours are scaled on the device by the
        float32 constant 1/255 exactly as voxel_data.h:82-85 does (b2v_grid_integrate_ex)."""
        if class_ids is not None or instance_ids is not None or depths is not None:
            raise NotImplementedError("semantic integration is a SURVEY.md §8(f) 'next' row")
        pts = np.asarray(points)
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise RuntimeError("points must be a 2D array with shape (N, 3)")
        if pts.dtype not in (np.float32, np.float64):
            raise RuntimeError("points must be float32 or float64")
        f64 = pts.dtype == np.float64
        pts = np.ascontiguousarray(pts)
        cp, cu8 = None, 0
        cols = None
        if colors is not None and np.asarray(colors).size > 0:
            cols = np.asarray(colors)
            if cols.ndim != 2 or cols.shape[1] != 3:
                raise RuntimeError("colors must be a 2D array with shape (N, 3)")
            if cols.shape[0] != pts.shape[0]:
                raise RuntimeError("points and colors must have the same number of rows")
            if cols.dtype == np.uint8:
                cu8 = 1
                cols = np.ascontiguousarray(cols)
            elif cols.dtype in (np.float32, np.float64):
                cols = np.ascontiguousarray(cols, dtype=np.float32)
            else:
                raise RuntimeError("colors must be uint8 or float32")
            cp = cols.ctypes.data
        self._check(self._L.b2v_grid_integrate_ex(self._h, pts.ctypes.data, 1 if f64 else 0, cp, cu8, pts.shape[0]),
                    "b2v_grid_integrate")
        self._check(self._L.b2v_grid_synchronize(self._h), "b2v_grid_synchronize")

    def integrate_rgbd(self, depth, color, K, Twc, max_depth=np.inf, min_depth=0.0, filter_shadow_points=False):
        """Fused front-end of `VolumetricIntegratorVoxelGrid.volume_integration`
        (volumetric_integrator_voxel_grid.py:247-300): `depth2pointcloud(depth, color, fx, fy, cx, cy,
        max_depth)` + `Twc` transform + `integrate(points, colors)` in one GPU call.  depth float32 [H,W]
        metres, color uint8 RGB [H,W,3], Twc = inv_T(pose) 4x4 float64.  `filter_shadow_points=True` applies
        the reference's shadow-point filter first (kVolumetricIntegrationVoxelGridShadowPointsFilter,
        voxel_grid.py:236-245)."""
        d = np.ascontiguousarray(depth, dtype=np.float32)
        c = np.ascontiguousarray(color)
        if d.ndim != 2 or c.shape != d.shape + (3,) or c.dtype != np.uint8:
            raise RuntimeError("depth must be float32 [H,W] and color uint8 [H,W,3]")
        K4 = _as_K4(K)
        T = np.ascontiguousarray(np.asarray(Twc, np.float64).reshape(4, 4)).reshape(16)
        mx = float(np.finfo(np.float32).max) if not np.isfinite(max_depth) else float(max_depth)
        self._check(self._L.b2v_grid_integrate_rgbd(self._h, d.ctypes.data, c.ctypes.data, d.shape[0], d.shape[1],
                                                    K4.ctypes.data, T.ctypes.data, mx, float(min_depth),
                                                    1 if filter_shadow_points else 0),
                    "b2v_grid_integrate_rgbd")
        self._check(self._L.b2v_grid_synchronize(self._h), "b2v_grid_synchronize")

    def get_voxels(self, min_count: int = 1, min_confidence: float = 0.0) -> VoxelGridData:
        """get_voxels(min_count, min_confidence): min_confidence is ignored for the non-semantic grid,
        as in the reference (voxel_block_grid.hpp:750-752)."""
        n = self._L.b2v_grid_get_voxels(self._h, int(min_count))
        if n < 0:
            raise RuntimeError(self._L.b2v_grid_last_error(self._h).decode())
        P = np.zeros((n, 3), np.float32)
        Cc = np.zeros((n, 3), np.float32)
        self._check(self._L.b2v_grid_copy_voxels(self._h, P.ctypes.data, Cc.ctypes.data),
                    "b2v_grid_copy_voxels")
        return VoxelGridData(P, Cc)

    def get_points(self):
        return self.get_voxels(1).points

    def get_colors(self):
        return self.get_voxels(1).colors

    def clear(self):
        self._check(self._L.b2v_grid_clear(self._h), "b2v_grid_clear")

    reset = clear

    def num_blocks(self) -> int:
        return int(self._L.b2v_grid_num_blocks(self._h))

    def size(self) -> int:
        return int(self._L.b2v_grid_size(self._h))

    def get_total_voxel_count(self) -> int:
        return self.size()

    def empty(self) -> bool:
        return self.num_blocks() == 0

    def get_block_size(self) -> int:
        return self._block_size

    def remove_low_count_voxels(self, min_count: int):
        self._check(self._L.b2v_grid_remove_low_count_voxels(self._h, int(min_count)),
                    "b2v_grid_remove_low_count_voxels")

    def remove_low_confidence_voxels(self, min_confidence: float):
        # no-op for the non-semantic grid, as in the reference (voxel_block_grid.hpp:650-676)
        return None

    def dump_blocks(self):
        nb = self.num_blocks()
        keys = np.zeros((nb, 3), np.int32)
        hashes = np.zeros(nb, np.uint64)
        count = np.zeros((nb, BLOCK_VOXELS), np.int32)
        pos = np.zeros((nb, BLOCK_VOXELS, 3), np.float32)
        col = np.zeros((nb, BLOCK_VOXELS, 3), np.float32)
        n = self._L.b2v_grid_dump_blocks(self._h, keys.ctypes.data, hashes.ctypes.data,
                                         count.ctypes.data, pos.ctypes.data, col.ctypes.data)
        if n != nb:
            raise RuntimeError(f"b2v_grid_dump_blocks returned {n}, expected {nb}")
        return dict(keys=keys, hashes=hashes, count=count, pos_sum=pos, col_sum=col)

    # ---- spatial queries and carving ----
    def _collect(self, n):
        if n < 0:
            raise RuntimeError(self._L.b2v_grid_last_error(self._h).decode())
        P = np.zeros((n, 3), np.float32)
        Cc = np.zeros((n, 3), np.float32)
        self._check(self._L.b2v_grid_copy_voxels(self._h, P.ctypes.data, Cc.ctypes.data), "b2v_grid_copy_voxels")
        return VoxelGridData(P, Cc)

    def carve(self, camera_frustrum, depth_image, depth_threshold: float = 1e-2):
        """carve(camera_frustrum, depth_image, depth_threshold) (voxel_block_grid.hpp:616-622): reset voxels
        in the frustum that lie in front of the observed depth by more than the threshold.  Like the
        reference (voxel_grid_carving.h:51-58) an empty or wrongly sized image is a soft failure."""
        d = np.asarray(depth_image)
        if d.size == 0 or d.shape != (camera_frustrum.height, camera_frustrum.width):
            print("volumetric::carve: depth image is empty or has the wrong size")
            return
        d = np.ascontiguousarray(d, dtype=np.float32)
        K, T = camera_frustrum._args()
        self._check(self._L.b2v_grid_carve(self._h, K.ctypes.data, camera_frustrum.width, camera_frustrum.height,
                                           T.ctypes.data, camera_frustrum.depth_max, camera_frustrum.depth_min,
                                           d.ctypes.data, float(depth_threshold)), "b2v_grid_carve")

    def get_voxels_in_camera_frustrum(self, camera_frustrum, min_count: int = 1, min_confidence: float = 0.0):
        K, T = camera_frustrum._args()
        return self._collect(self._L.b2v_grid_get_voxels_in_frustum(
            self._h, K.ctypes.data, camera_frustrum.width, camera_frustrum.height, T.ctypes.data,
            camera_frustrum.depth_max, camera_frustrum.depth_min, int(min_count)))

    def get_voxels_in_bb(self, bbox, min_count: int = 1, min_confidence: float = 0.0):
        bb = np.ascontiguousarray(getattr(bbox, "bounds", bbox), np.float64).reshape(6)
        return self._collect(self._L.b2v_grid_get_voxels_in_bb(self._h, bb.ctypes.data, int(min_count)))


class OrientedBoundingBox3D:
    """`volumetric.OrientedBoundingBox3D` fields: center [3], size [3], orientation (unit quaternion w, x, y, z of the
    box -> world rotation).  `compute_from_points` is the reference's PCA method (bounding_boxes_3d.cpp:373-556):
    running centroid / covariance (Welford) in float64, eigenvectors sorted by descending eigenvalue, right-handed,
    extents from the min / max of the points in that frame."""

    def __init__(self, center=None, size=None, rotation=None):
        self.center = np.zeros(3) if center is None else np.asarray(center, np.float64)
        self.size = np.zeros(3) if size is None else np.asarray(size, np.float64)
        self.R = np.eye(3) if rotation is None else np.asarray(rotation, np.float64)

    @property
    def orientation(self):
        return _quat_wxyz(self.R)

    def get_matrix(self):
        M = np.eye(4)
        M[:3, :3], M[:3, 3] = self.R, self.center
        return M

    def get_corners(self):
        h = self.size / 2.0
        sg = np.array([[1, 1, -1], [-1, 1, -1], [-1, -1, -1], [1, -1, -1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1], [1, -1, 1]], float)
        return self.center + (sg * h) @ self.R.T

    @staticmethod
    def compute_from_points(points):
        P = np.asarray(points, np.float64).reshape(-1, 3)
        n = len(P)
        if n == 0:
            return OrientedBoundingBox3D()
        if n == 1:
            return OrientedBoundingBox3D(P[0], np.zeros(3), np.eye(3))
        if n == 2:
            c, diff = 0.5 * (P[0] + P[1]), P[1] - P[0]
            dn = np.linalg.norm(diff)
            if dn < 1e-10:
                return OrientedBoundingBox3D(c, np.zeros(3), np.eye(3))
            a1 = diff / dn
            ref = np.array([1.0, 0, 0]) if abs(a1[0]) < 0.9 else np.array([0, 1.0, 0])
            a2 = np.cross(ref, a1)
            a2 /= np.linalg.norm(a2)
            a3 = np.cross(a1, a2)
            a3 /= np.linalg.norm(a3)
            R = np.stack([a1, a2, a3], axis=1)
            if np.linalg.det(R) < 0:
                R[:, 2] = -R[:, 2]
            centroid = c
        else:
            # the reference's one-pass Welford update equals the two-pass centroid / covariance up to rounding
            centroid = P.mean(axis=0)
            d = P - centroid
            cov = d.T @ d / n
            w, V = np.linalg.eigh(cov)
            R = V[:, np.argsort(-w, kind="stable")]
            if np.dot(np.cross(R[:, 0], R[:, 1]), R[:, 2]) < 0:
                R[:, 2] = -R[:, 2]
        local = (P - centroid) @ R
        lo, hi = local.min(axis=0), local.max(axis=0)
        return OrientedBoundingBox3D(centroid + R @ (0.5 * (hi + lo)), hi - lo, R)


def _quat_wxyz(R):
    """Rotation matrix -> unit quaternion (w, x, y, z), Eigen's branch order."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = [0.0] * 4
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return np.array(q)


class ObjectData:
    """`volumetric.ObjectData` (voxel_grid_data.h:64-79)."""

    def __init__(self, object_id, class_id, points, colors, confidence_min, confidence_max):
        self.object_id, self.class_id = int(object_id), int(class_id)
        self.points, self.colors = points, colors
        self.confidence_min, self.confidence_max = confidence_min, confidence_max
        self.oriented_bounding_box = OrientedBoundingBox3D.compute_from_points(points)


class ObjectDataGroup:
    """`volumetric.ObjectDataGroup` (voxel_grid_data.h:87-96)."""

    def __init__(self, objects):
        self.object_vector = objects
        self.class_ids = [o.class_id for o in objects]
        self.object_ids = [o.object_id for o in objects]


class ClassData:
    """`volumetric.ClassData` (voxel_grid_data.h:110-125)."""

    def __init__(self, class_id, points, colors, confidence_min, confidence_max):
        self.class_id = int(class_id)
        self.points, self.colors = points, colors
        self.confidence_min, self.confidence_max = confidence_min, confidence_max


class ClassDataGroup:
    """`volumetric.ClassDataGroup` (voxel_grid_data.h:127-140)."""

    def __init__(self, classes):
        self.class_vector = classes
        self.class_ids = [c.class_id for c in classes]


class VoxelBlockSemanticGrid:
    """GPU drop-in for `volumetric.VoxelBlockSemanticGrid(voxel_size, block_size=8)` — per-voxel label
    *voting* (cpp/volumetric/voxel_block_semantic_grid.h:59-118; voxel_data_semantic.h:106-199).

    integrate(points, colors, class_ids, instance_ids, depths) -> get_voxels(min_count, min_confidence) with
    `class_ids / object_ids / confidences`.  Observations reach a voxel in input order (the reference's
    deterministic build), so labels, counters, float64 position sums and float32 colour sums are bit-identical."""

    KIND = _lib.B2V_SEM_VOTING

    def __init__(self, voxel_size: float = 0.05, block_size: int = 8, capacity_blocks: int = 1 << 14,
                 device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.voxel_size = float(voxel_size)
        self._block_size = int(block_size)
        rc = self._L.b2v_sgrid_create(float(voxel_size), int(block_size), int(capacity_blocks), self.KIND,
                                      int(device), C.byref(self._h))
        if rc != _lib.B2V_OK:
            msg = self._L.b2v_sgrid_last_error(self._h).decode() if self._h else "invalid configuration"
            if self._h:
                self._L.b2v_sgrid_destroy(self._h)
                self._h = C.c_void_p()
            raise RuntimeError(f"b2v_sgrid_create failed (status {rc}): {msg}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.b2v_sgrid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != _lib.B2V_OK:
            raise RuntimeError(f"{what} failed (status {rc}): {self._L.b2v_sgrid_last_error(self._h).decode()}")

    # ---- parameters (class-static in the reference, per grid here) ----
    def set_depth_threshold(self, depth_threshold: float):
        self._check(self._L.b2v_sgrid_set_depth_threshold(self._h, float(depth_threshold)), "set_depth_threshold")

    def set_depth_decay_rate(self, depth_decay_rate: float):
        self._check(self._L.b2v_sgrid_set_depth_decay_rate(self._h, float(depth_decay_rate)), "set_depth_decay_rate")

    # ---- integrate (volumetric_grid_module.h: integrate(points, colors, class_ids, instance_ids, depths)) ----
    def integrate(self, points, colors=None, class_ids=None, instance_ids=None, depths=None):
        pts = np.asarray(points)
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise RuntimeError("points must be a 2D array with shape (N, 3)")
        if pts.dtype not in (np.float32, np.float64):
            raise RuntimeError("points must be float32 or float64")
        pts = np.ascontiguousarray(pts)
        n = pts.shape[0]
        cp, cu8, cols = None, 0, None
        if colors is not None and np.asarray(colors).size > 0:
            cols = np.asarray(colors)
            if cols.ndim != 2 or cols.shape[1] != 3 or cols.shape[0] != n:
                raise RuntimeError("points and colors must have the same size")
            if cols.dtype == np.uint8:
                cu8 = 1
                cols = np.ascontiguousarray(cols)
            elif cols.dtype in (np.float32, np.float64):
                cols = np.ascontiguousarray(cols, dtype=np.float32)
            else:
                raise RuntimeError("colors must be uint8 or float32")
            cp = cols.ctypes.data
        hold = []

        def opt(a, dt, name):
            if a is None or np.asarray(a).size == 0:
                return None
            b = np.ascontiguousarray(a, dtype=dt).reshape(-1)
            if b.shape[0] != n:
                raise RuntimeError(f"points and {name} must have the same size")
            hold.append(b)
            return b.ctypes.data

        ci = opt(class_ids, np.int32, "class_ids")
        ii = opt(instance_ids, np.int32, "instance_ids")
        di = opt(depths, np.float32, "depths")
        if ii is not None and ci is None:
            raise RuntimeError("instance_ids but no class_ids is not supported")  # voxel_block_grid.hpp:43-46
        self._check(self._L.b2v_sgrid_integrate(self._h, n, pts.ctypes.data, 1 if pts.dtype == np.float64 else 0,
                                                cp, cu8, ci, ii, di), "b2v_sgrid_integrate")

    def integrate_segment(self, points, colors, class_id: int, object_id: int):
        """`integrate_segment(points, colors, class_id, object_id)` (volumetric_grid_module.h:97-125, 564-590 ->
        voxel_block_semantic_grid.hpp:52-99): every point carries the same (class, object) label; a negative id
        skips the whole segment."""
        pts = np.asarray(points)
        cols = np.asarray(colors)
        if pts.ndim != 2 or pts.shape[1] != 3:
            raise RuntimeError("points must be a contiguous Nx3 array")
        if cols.ndim != 2 or cols.shape[1] != 3:
            raise RuntimeError("colors must be a contiguous Nx3 array")
        if cols.shape[0] != pts.shape[0]:
            raise RuntimeError("points and colors must have the same size")
        if int(object_id) < 0 or int(class_id) < 0 or pts.shape[0] == 0:
            return
        n = pts.shape[0]
        self.integrate(pts, cols, np.full(n, int(class_id), np.int32), np.full(n, int(object_id), np.int32))

    # ---- segments (voxel_block_semantic_grid.hpp:204-316) ----
    def _segments(self, by_class: bool, min_count: int, min_confidence: float):
        # the reference keeps voxels with count > min_count (strict, unlike get_voxels' >=) and confidence >=
        # min_confidence: the GPU read-out (count -> scan -> emit) does the scan and the compaction ...
        v = self.get_voxels(int(min_count) + 1, float(min_confidence))
        ids = np.asarray(v.class_ids if by_class else v.object_ids)
        keep = ids >= 0                                   # negative = uninitialised label
        pts, cols = np.asarray(v.points)[keep], np.asarray(v.colors)[keep]
        cls, conf, ids = np.asarray(v.class_ids)[keep], np.asarray(v.confidences)[keep], ids[keep]
        # ... and the host groups the survivors by id (stable: block order within a segment is kept)
        order = np.argsort(ids, kind="stable")
        uniq, start = np.unique(ids[order], return_index=True)
        bounds = list(start) + [len(order)]
        out = []
        for k, seg_id in enumerate(uniq):
            sel = order[bounds[k]:bounds[k + 1]]
            out.append((int(seg_id), pts[sel], cols[sel], int(cls[sel[0]]), float(conf[sel].min()),
                        float(conf[sel].max())))
        return out

    def get_object_segments(self, min_count: int = 1, min_confidence: float = 0.0):
        """`get_object_segments(min_count, min_confidence)` -> ObjectDataGroup (voxel_grid_data.h:64-96): voxels grouped
        by object id (ids < 0 dropped), each with its points / colours, the class id of its first voxel, the
        confidence range and a PCA oriented bounding box (bounding_boxes_3d.cpp:373-556, the reference's default
        OBBComputationMethod::PCA)."""
        objs = [ObjectData(i, c, p, col, cmin, cmax) for i, p, col, c, cmin, cmax in
                self._segments(False, min_count, min_confidence)]
        return ObjectDataGroup(objs)

    def get_class_segments(self, min_count: int = 1, min_confidence: float = 0.0):
        """`get_class_segments(min_count, min_confidence)` -> ClassDataGroup (voxel_grid_data.h:110-140)."""
        return ClassDataGroup([ClassData(i, p, col, cmin, cmax) for i, p, col, _c, cmin, cmax in
                               self._segments(True, min_count, min_confidence)])

    def integrate_rgbd(self, depth, color, K, Twc, class_image=None, object_image=None, max_depth=np.inf,
                       min_depth=0.0, use_depths=True, filter_shadow_points=False):
        """The reference integrator's per-frame front-end fused on the GPU
        (volumetric_integrator_voxel_semantic_grid.py:332-461): optional `filter_shadow_points`, `depth2pointcloud`
        with the class / object-id images, world transform by `Twc` (camera -> world), `integrate`.  `use_depths`
        mirrors kVolumetricSemanticProbabilisticIntegrationUseDepth.  color is RGB uint8."""
        d = np.ascontiguousarray(depth, np.float32)
        c = np.ascontiguousarray(color, np.uint8)
        if d.ndim != 2 or c.shape != d.shape + (3,):
            raise RuntimeError("depth must be [H,W] float32 and color [H,W,3] uint8")
        hold = [d, c]

        def img(a, name):
            if a is None or np.asarray(a).size == 0:
                return None
            b = np.ascontiguousarray(a, np.int32)
            if b.shape != d.shape:
                raise RuntimeError(f"{name} must have the depth image's shape")
            hold.append(b)
            return b.ctypes.data

        ci, oi = img(class_image, "class_image"), img(object_image, "object_image")
        K4 = _as_K4(K)
        T = np.ascontiguousarray(np.asarray(Twc, np.float64).reshape(16))
        md = float(np.finfo(np.float32).max) if not np.isfinite(max_depth) else float(max_depth)
        self._check(self._L.b2v_sgrid_integrate_rgbd(self._h, d.ctypes.data, c.ctypes.data, ci, oi, d.shape[0],
                                                     d.shape[1], K4.ctypes.data, T.ctypes.data, md, float(min_depth),
                                                     1 if use_depths else 0, 1 if filter_shadow_points else 0),
                    "b2v_sgrid_integrate_rgbd")

    # ---- read-outs ----
    def get_voxels(self, min_count: int = 1, min_confidence: float = 0.0) -> VoxelGridData:
        return self._collect(self._L.b2v_sgrid_get_voxels(self._h, int(min_count), float(min_confidence)))

    def get_voxels_in_bb(self, bbox, min_count: int = 1, min_confidence: float = 0.0) -> VoxelGridData:
        bb = np.ascontiguousarray(getattr(bbox, "bounds", bbox), np.float64).reshape(6)
        return self._collect(self._L.b2v_sgrid_get_voxels_in_bb(self._h, bb.ctypes.data, int(min_count),
                                                                float(min_confidence)))

    def get_voxels_in_camera_frustrum(self, camera_frustrum, min_count: int = 1,
                                      min_confidence: float = 0.0) -> VoxelGridData:
        K, T = camera_frustrum._args()
        return self._collect(self._L.b2v_sgrid_get_voxels_in_frustum(
            self._h, K.ctypes.data, camera_frustrum.width, camera_frustrum.height, T.ctypes.data,
            camera_frustrum.depth_max, camera_frustrum.depth_min, int(min_count), float(min_confidence)))

    def _collect(self, n) -> VoxelGridData:
        if n < 0:
            raise RuntimeError(self._L.b2v_sgrid_last_error(self._h).decode())
        out = VoxelGridData(np.zeros((n, 3), np.float64), np.zeros((n, 3), np.float32))
        out.class_ids = np.zeros(n, np.int32)
        out.object_ids = np.zeros(n, np.int32)
        out.confidences = np.zeros(n, np.float32)
        if n:
            self._check(self._L.b2v_sgrid_copy_voxels(self._h, out.points.ctypes.data, out.colors.ctypes.data,
                                                      out.class_ids.ctypes.data, out.object_ids.ctypes.data,
                                                      out.confidences.ctypes.data), "b2v_sgrid_copy_voxels")
        return out

    def get_points(self):
        return self.get_voxels(1, 0.0).points

    def get_colors(self):
        return self.get_voxels(1, 0.0).colors

    def get_ids(self):
        """(class_ids, object_ids) of every non-empty voxel (voxel_block_semantic_grid.hpp:185-202)."""
        v = self.get_voxels(1, -np.inf)
        return v.class_ids, v.object_ids

    def num_blocks(self) -> int:
        n = self._L.b2v_sgrid_num_blocks(self._h)
        if n < 0:
            raise RuntimeError("b2v_sgrid_num_blocks failed")
        return int(n)

    def get_block_size(self) -> int:
        return self._block_size

    def size(self) -> int:
        return len(self.get_voxels(1, -np.inf).points)

    def empty(self) -> bool:
        return self.num_blocks() == 0

    def clear(self):
        self._check(self._L.b2v_sgrid_clear(self._h), "b2v_sgrid_clear")

    reset = clear

    def remove_low_count_voxels(self, min_count: int):
        self._check(self._L.b2v_sgrid_remove_low_count_voxels(self._h, int(min_count)), "remove_low_count_voxels")

    def remove_low_confidence_segments(self, min_confidence: int):
        """The reference takes an `int` threshold (voxel_block_semantic_grid.h:103)."""
        self._check(self._L.b2v_sgrid_remove_low_confidence_segments(self._h, int(min_confidence)),
                    "remove_low_confidence_segments")

    def merge_segments(self, instance_id1: int, instance_id2: int):
        self._check(self._L.b2v_sgrid_merge_segments(self._h, int(instance_id1), int(instance_id2)), "merge_segments")

    def remove_segment(self, object_id: int):
        self._check(self._L.b2v_sgrid_remove_segment(self._h, int(object_id)), "remove_segment")

    def carve(self, camera_frustrum, depth_image, depth_threshold: float = 1e-2):
        """carve(camera_frustrum, depth_image, depth_threshold) (voxel_block_grid.hpp:616-622); a wrongly sized
        image is a soft failure like in the reference (voxel_grid_carving.h:51-58)."""
        d = np.asarray(depth_image)
        if d.size == 0 or d.shape != (camera_frustrum.height, camera_frustrum.width):
            print("volumetric::carve: depth image is empty or has the wrong size")
            return
        d = np.ascontiguousarray(d, dtype=np.float32)
        K, T = camera_frustrum._args()
        self._check(self._L.b2v_sgrid_carve(self._h, K.ctypes.data, camera_frustrum.width, camera_frustrum.height,
                                            T.ctypes.data, camera_frustrum.depth_max, camera_frustrum.depth_min,
                                            d.ctypes.data, float(depth_threshold)), "b2v_sgrid_carve")

    def assign_object_ids_to_instance_ids(self, camera_frustrum, class_ids_image, semantic_instances_image,
                                          depth_image=None, depth_threshold: float = 0.1, do_carving: bool = False,
                                          min_vote_ratio: float = 0.5, min_votes: int = 3) -> dict:
        """`MapInstanceIdToObjectId` of the frame (voxel_block_semantic_grid.h:67-71;
        voxel_semantic_data_association.h:69-373): 2-D instance id -> 3-D object id (-1: no confident match).
        Soft failures (empty / wrongly sized label images) return an empty map like the reference (:80-103)."""
        hw = (camera_frustrum.height, camera_frustrum.width)
        ci, ii = np.asarray(class_ids_image), np.asarray(semantic_instances_image)
        if ci.size == 0 or ii.size == 0 or ci.shape != hw or ii.shape != hw:
            print("volumetric::assign_object_ids_to_instance_ids: label images are empty or have the wrong size")
            return {}
        ci = np.ascontiguousarray(ci, np.int32)
        ii = np.ascontiguousarray(ii, np.int32)
        dp, d = None, None
        if depth_image is not None and np.asarray(depth_image).size and np.asarray(depth_image).shape == hw:
            d = np.ascontiguousarray(depth_image, np.float32)
            dp = d.ctypes.data
        K, T = camera_frustrum._args()
        n = self._L.b2v_sgrid_assign_object_ids_to_instance_ids(
            self._h, K.ctypes.data, camera_frustrum.width, camera_frustrum.height, T.ctypes.data,
            camera_frustrum.depth_max, camera_frustrum.depth_min, ci.ctypes.data, ii.ctypes.data, dp,
            float(depth_threshold), 1 if do_carving else 0, float(min_vote_ratio), int(min_votes))
        if n < 0:
            raise RuntimeError(self._L.b2v_sgrid_last_error(self._h).decode())
        ids, objs = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._check(self._L.b2v_sgrid_copy_instance_map(self._h, ids.ctypes.data, objs.ctypes.data),
                    "b2v_sgrid_copy_instance_map")
        return {int(i): int(o) for i, o in zip(ids, objs)}

    def set_next_object_id(self, next_object_id: int):
        self._check(self._L.b2v_sgrid_set_next_object_id(self._h, int(next_object_id)), "set_next_object_id")

    def get_next_object_id(self) -> int:
        return int(self._L.b2v_sgrid_get_next_object_id(self._h))

    def label_overflows(self) -> int:
        out = C.c_uint64(0)
        self._check(self._L.b2v_sgrid_label_overflows(self._h, C.byref(out)), "b2v_sgrid_label_overflows")
        return int(out.value)

    def dump_blocks(self, K: int = 8):
        """Parity hook: per-block arrays [nb,512,...] incl. labels (see include/b2v.h)."""
        nb, nv = self.num_blocks(), BLOCK_VOXELS
        d = dict(keys=np.zeros((nb, 3), np.int32), hashes=np.zeros(nb, np.uint64),
                 count=np.zeros((nb, nv), np.int32), pos_sum=np.zeros((nb, nv, 3), np.float64),
                 col_sum=np.zeros((nb, nv, 3), np.float32), object_id=np.zeros((nb, nv), np.int32),
                 class_id=np.zeros((nb, nv), np.int32), confidence=np.zeros((nb, nv), np.float32),
                 aux=np.zeros((nb, nv), np.int32), lab_obj=np.full((nb, nv, K), -1, np.int32),
                 lab_cls=np.full((nb, nv, K), -1, np.int32), lab_logp=np.full((nb, nv, K), -np.inf, np.float32))
        if nb:
            n = self._L.b2v_sgrid_dump_blocks(
                self._h, d["keys"].ctypes.data, d["hashes"].ctypes.data, d["count"].ctypes.data,
                d["pos_sum"].ctypes.data, d["col_sum"].ctypes.data, d["object_id"].ctypes.data,
                d["class_id"].ctypes.data, d["confidence"].ctypes.data, d["aux"].ctypes.data, int(K),
                d["lab_obj"].ctypes.data, d["lab_cls"].ctypes.data, d["lab_logp"].ctypes.data)
            if n != nb:
                raise RuntimeError(f"b2v_sgrid_dump_blocks returned {n}, expected {nb}")
        return d


class VoxelBlockSemanticProbabilisticGrid(VoxelBlockSemanticGrid):
    """GPU drop-in for `volumetric.VoxelBlockSemanticProbabilisticGrid` — Bayesian label fusion in log space
    over joint (object, class) pairs with depth-decayed evidence (voxel_data_semantic.h:249-672)."""

    KIND = _lib.B2V_SEM_PROBABILISTIC


def remap_instance_ids(instance_ids, instance_to_object: dict, invalid_instance_id: int = -1):
    """`volumetric.remap_instance_ids(instance_ids_image, map)` (cpp/volumetric/image_utils.h:69-163): every pixel's
    instance id is replaced by its object id; ids missing from the map - and everything when the map is empty -
    become `invalid_instance_id`."""
    img = np.ascontiguousarray(instance_ids, np.int32)
    if img.size == 0:
        return img
    out = np.full(img.shape, invalid_instance_id, np.int32)
    if instance_to_object:
        keys = np.fromiter(instance_to_object.keys(), np.int64, len(instance_to_object))
        vals = np.fromiter(instance_to_object.values(), np.int64, len(instance_to_object))
        order = np.argsort(keys)
        keys, vals = keys[order], vals[order]
        pos = np.clip(np.searchsorted(keys, img), 0, len(keys) - 1)
        hit = keys[pos] == img
        out[hit] = vals[pos][hit]
    return out


# The direct (non-block) grids of the reference's known-answer tests (cpp/test_volumetric_voxel_semantic.py) hold
# the same voxel records; only the container differs, so the block grids serve as their drop-in as well.
VoxelSemanticGrid = VoxelBlockSemanticGrid
VoxelSemanticGridProbabilistic = VoxelBlockSemanticProbabilisticGrid
