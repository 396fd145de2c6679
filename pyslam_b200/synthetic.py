"""Seeded synthetic RGBD sequences for the five BASELINE.json config shapes.

The reference has no RGBD fixture for this path (`cpp/test_volumetric.py:438,462-463` is a
3-point link test), so the sequences of SURVEY.md §8(d) are generated here:
an analytic scene (axis-aligned room + spheres + one tilted finite plane), exact ray-cast
z-depth (float32 metres), procedural colour (uint8 RGB), ~2 % invalid (zero-depth) pixels and
closed-form camera trajectories.  Everything is float64 numpy until the final casts and is a
pure function of (config name, frame index), so tests, golden fixtures and `bench.py` agree.

Conventions follow the reference front-end: pose is **Tcw** (world->camera) 4x4 float64
(`pyslam/dense/volumetric_integrator_base.py:116`), depth is metric float32
(`base.py:713`), colour is RGB uint8 when it reaches the volume (`base.py:1054`).
World and camera axes: x right, y down, z forward.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass(frozen=True)
class Scene:
    """Axis-aligned room (camera inside) + spheres + one finite tilted plane (a disc)."""

    room_min: tuple = (-2.0, -1.25, -2.0)
    room_max: tuple = (2.0, 1.25, 2.0)
    spheres: tuple = (
        ((0.9, 0.75, 1.2), 0.5),
        ((-1.1, 0.55, 0.6), 0.35),
        ((0.2, 0.85, -1.3), 0.4),
    )
    # disc: centre, unit normal, radius
    disc_center: tuple = (-0.6, 0.2, -0.9)
    disc_normal: tuple = (0.35, -0.8, 0.48)
    disc_radius: float = 0.7
    checker_scale: float = 4.0  # colour cells per metre


@dataclass(frozen=True)
class SequenceConfig:
    name: str
    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float
    voxel_size: float
    sdf_trunc: float
    depth_trunc: float
    n_frames: int
    seed: int
    scene: Scene = field(default_factory=Scene)
    trajectory: str = "orbit"  # "orbit" | "street"
    orbit_radius: float = 0.5
    invalid_fraction: float = 0.02
    n_classes: int = 0  # >0: also render an int32 class-id image (C3)

    @property
    def K(self) -> np.ndarray:
        """(fx, fy, cx, cy) float64 — the north_star `K` argument."""
        return np.array([self.fx, self.fy, self.cx, self.cy], dtype=np.float64)


_STREET = Scene(
    room_min=(-6.0, -28.0, -10.0),
    room_max=(6.0, 1.65, 2010.0),
    spheres=(((2.5, 0.9, 30.0), 0.9), ((-2.8, 0.8, 75.0), 1.0), ((3.0, 0.7, 140.0), 1.1)),
    disc_center=(0.0, 1.2, 55.0),
    disc_normal=(0.0, -0.96, 0.28),
    disc_radius=2.0,
    checker_scale=0.5,
)

#: SURVEY.md §8(d) / BASELINE.md §3 configs.  Intrinsics: settings/TUM1.yaml:27-30,
#: settings/REPLICA.yaml:27-39, settings/SCANNET.yaml:34-46, settings/KITTI00-02.yaml:23-34
#: (distortion zeroed).
CONFIGS = {
    "C1": SequenceConfig("C1", 320, 240, 262.5, 262.5, 159.5, 119.5, 0.01, 0.04, 4.0, 100, 1),
    "C2": SequenceConfig("C2", 640, 480, 517.306408, 516.469215, 318.643040, 255.313989,
                         0.005, 0.04, 4.0, 300, 2),
    "C3": SequenceConfig("C3", 1200, 680, 600.0, 600.0, 599.5, 339.5, 0.005, 0.04, 4.0, 300, 3,
                         n_classes=40),
    "C4": SequenceConfig("C4", 640, 480, 577.87056, 580.25846, 319.8765, 239.876,
                         0.004, 0.04, 4.0, 1000, 4),
    "C5": SequenceConfig("C5", 1241, 376, 718.856, 718.856, 607.1928, 185.2157,
                         0.10, 0.4, 10.0, 4541, 5, scene=_STREET, trajectory="street"),
    # tiny case for unit tests / smoke (not a BASELINE config)
    "T0": SequenceConfig("T0", 96, 72, 80.0, 80.0, 47.5, 35.5, 0.02, 0.08, 4.0, 24, 7),
}


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def _rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def pose_Twc(cfg: SequenceConfig, i: int) -> np.ndarray:
    """Camera-to-world pose of frame i (closed form)."""
    T = np.eye(4)
    if cfg.trajectory == "street":
        yaw = 0.04 * np.sin(0.013 * i)
        T[:3, :3] = _rot_y(yaw) @ _rot_x(0.02 * np.sin(0.021 * i))
        T[:3, 3] = (0.3 * np.sin(0.005 * i), 0.0, 0.44 * i)
        return T
    n = max(cfg.n_frames, 1)
    th = 2.0 * np.pi * i / n
    pitch = 0.25 * np.sin(3.0 * th) + 0.1
    T[:3, :3] = _rot_y(th) @ _rot_x(pitch)
    r = cfg.orbit_radius
    T[:3, 3] = (r * np.sin(th), 0.15 * np.sin(2.0 * th), r * np.cos(th))
    return T


def inv_T(T: np.ndarray) -> np.ndarray:
    """Rigid inverse (same convention as the reference's `inv_T`, voxel_grid.py:207)."""
    R, t = T[:3, :3], T[:3, 3]
    out = np.eye(4)
    out[:3, :3] = R.T
    out[:3, 3] = -R.T @ t
    return out


def pose_Tcw(cfg: SequenceConfig, i: int) -> np.ndarray:
    return np.ascontiguousarray(inv_T(pose_Twc(cfg, i)), dtype=np.float64)


def _rotate(dc: np.ndarray, R: np.ndarray) -> np.ndarray:
    """dc [H,W,3] -> R @ dc per pixel, written out element-wise (no BLAS: bit-reproducible)."""
    return np.stack([dc[..., 0] * R[a, 0] + dc[..., 1] * R[a, 1] + dc[..., 2] * R[a, 2]
                     for a in range(3)], axis=-1)


def _raycast(scene: Scene, o: np.ndarray, d: np.ndarray):
    """o [3], d [H,W,3] with d_cam.z == 1 so that t is z-depth.  Returns (t, object id)."""
    H, W, _ = d.shape
    t_best = np.full((H, W), np.inf)
    obj = np.zeros((H, W), dtype=np.int32)
    # room: exit distance of a ray that starts inside the box
    lo, hi = np.array(scene.room_min), np.array(scene.room_max)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_axis = np.where(d > 0, (hi - o) / d, np.where(d < 0, (lo - o) / d, np.inf))
    face = np.argmin(t_axis, axis=-1)
    t_room = np.min(t_axis, axis=-1)
    t_best = t_room
    obj = 1 + face.astype(np.int32) * 2 + (np.take_along_axis(d, face[..., None], -1)[..., 0] > 0)
    # spheres
    for k, (c, r) in enumerate(scene.spheres):
        oc = o - np.array(c)
        a = np.sum(d * d, axis=-1)
        b = 2.0 * np.sum(d * oc, axis=-1)
        cc = float(oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2]) - r * r
        disc = b * b - 4 * a * cc
        with np.errstate(invalid="ignore"):
            ts = (-b - np.sqrt(disc)) / (2 * a)
        hit = (disc > 0) & (ts > 1e-6) & (ts < t_best)
        t_best = np.where(hit, ts, t_best)
        obj = np.where(hit, 10 + k, obj)
    # tilted disc
    n = np.array(scene.disc_normal, dtype=np.float64)
    n = n / np.linalg.norm(n)
    c = np.array(scene.disc_center)
    denom = d[..., 0] * n[0] + d[..., 1] * n[1] + d[..., 2] * n[2]
    co = c - o
    with np.errstate(divide="ignore", invalid="ignore"):
        tp = (co[0] * n[0] + co[1] * n[1] + co[2] * n[2]) / denom
    p = o + tp[..., None] * d
    hit = (np.abs(denom) > 1e-9) & (tp > 1e-6) & (tp < t_best) & \
          (np.sum((p - c) ** 2, axis=-1) < scene.disc_radius ** 2)
    t_best = np.where(hit, tp, t_best)
    obj = np.where(hit, 20, obj)
    return t_best, obj


def _procedural_rgb(scene: Scene, p: np.ndarray, obj: np.ndarray) -> np.ndarray:
    s = scene.checker_scale
    cell = np.floor(p * s + 1e-9).astype(np.int64)
    parity = (cell[..., 0] + cell[..., 1] + cell[..., 2]) & 1
    lo, hi = np.array(scene.room_min), np.array(scene.room_max)
    g = np.clip((p - lo) / (hi - lo), 0.0, 1.0)
    base = 0.25 + 0.6 * g
    tint = ((obj[..., None] * np.array([37, 91, 53])) % 64) / 255.0
    col = base * (0.55 + 0.45 * parity[..., None]) + tint
    return np.clip(np.round(col * 255.0), 0, 255).astype(np.uint8)


def render_frame(cfg: SequenceConfig, i: int, noise_sigma: float = 0.0):
    """Frame i of a config -> (depth f32 [H,W], rgb u8 [H,W,3], Tcw f64 [4,4])."""
    Twc = pose_Twc(cfg, i)
    R, o = Twc[:3, :3], Twc[:3, 3]
    u = (np.arange(cfg.width, dtype=np.float64) - cfg.cx) / cfg.fx
    v = (np.arange(cfg.height, dtype=np.float64) - cfg.cy) / cfg.fy
    dc = np.stack(np.broadcast_arrays(u[None, :], v[:, None], 1.0), axis=-1)
    d = _rotate(dc, R)
    t, obj = _raycast(cfg.scene, o, d)
    p = o + t[..., None] * d
    rgb = _procedural_rgb(cfg.scene, p, obj)
    rng = np.random.default_rng([cfg.seed, i])
    depth = t.copy()
    if noise_sigma > 0:
        depth = depth + rng.normal(0.0, noise_sigma, size=depth.shape)
    depth[~np.isfinite(depth)] = 0.0
    depth[rng.random(depth.shape) < cfg.invalid_fraction] = 0.0
    return (np.ascontiguousarray(depth, dtype=np.float32), np.ascontiguousarray(rgb),
            np.ascontiguousarray(inv_T(Twc), dtype=np.float64))


def render_class_ids(cfg: SequenceConfig, i: int) -> np.ndarray:
    """int32 class-id image (piecewise constant by object), for the C3 semantics row."""
    Twc = pose_Twc(cfg, i)
    R, o = Twc[:3, :3], Twc[:3, 3]
    u = (np.arange(cfg.width, dtype=np.float64) - cfg.cx) / cfg.fx
    v = (np.arange(cfg.height, dtype=np.float64) - cfg.cy) / cfg.fy
    dc = np.stack(np.broadcast_arrays(u[None, :], v[:, None], 1.0), axis=-1)
    _, obj = _raycast(cfg.scene, o, _rotate(dc, R))
    k = max(cfg.n_classes, 1)
    return np.ascontiguousarray((obj * 7) % k, dtype=np.int32)


def sequence(cfg: SequenceConfig, n_frames: int | None = None, start: int = 0, step: int = 1):
    """Yield (depth, rgb, Tcw) for frames start, start+step, ..."""
    n = cfg.n_frames if n_frames is None else n_frames
    for k in range(n):
        yield render_frame(cfg, start + k * step)
