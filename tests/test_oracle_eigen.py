"""Pins the two evaluation-order assumptions of oracle/open3d_order.c against REAL Eigen (the copy vendored in the
reference tree, compiled by oracle/Makefile into oracle/_ref/libeigen_ops.so for the x86-64 baseline of Open3D's
wheels).  CPU only."""
import numpy as np
import pytest

import oracle
from pyslam_b200 import synthetic as S

pytestmark = pytest.mark.skipif(not oracle.have_eigen_ops(), reason="oracle/_ref/libeigen_ops.so not built")


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def test_eigen_is_the_3_4_line():
    assert oracle.EigenOps().version // 100 == 304   # 3.4.x, the line Open3D 0.19 builds against


def test_matrix4f_times_vector4f_accumulates_columns_left_to_right():
    """UniformTSDFVolume: pt_camera = extrinsic_f * pt_3d_homo.  open3d_order.c and the kernels compute
    ((E0*h0 + E1*h1) + E2*h2) + E3*1: the bit patterns must equal Eigen's."""
    E = oracle.EigenOps()
    rng = np.random.default_rng(0)
    cfg = S.CONFIGS["C2"]
    mats = [S.pose_Tcw(cfg, i).astype(np.float32) for i in range(0, 300, 7)]   # extrinsic.cast<float>() of real poses
    mats += [rng.standard_normal((4, 4)).astype(np.float32) for _ in range(200)]
    for M in mats:
        for _ in range(50):
            v = rng.uniform(-4.0, 4.0, 4).astype(np.float32)
            v[3] = 1.0
            ref = ((M[:, 0] * v[0] + M[:, 1] * v[1]) + M[:, 2] * v[2]) + M[:, 3] * v[3]
            assert np.array_equal(_bits(E.mat4f_times_vec4f(M, v)), _bits(ref.astype(np.float32)))


def test_matrix4d_times_vector4d_accumulates_columns_left_to_right():
    """PointCloudFactory: point = camera_pose * Vector4d(x, y, z, 1) (the allocation samples)."""
    E = oracle.EigenOps()
    rng = np.random.default_rng(1)
    for _ in range(2000):
        M = rng.standard_normal((4, 4))
        v = rng.uniform(-4.0, 4.0, 4)
        v[3] = 1.0
        ref = ((M[:, 0] * v[0] + M[:, 1] * v[1]) + M[:, 2] * v[2]) + M[:, 3] * v[3]
        assert np.array_equal(_bits(E.mat4d_times_vec4d(M, v)), _bits(ref))


@pytest.mark.parametrize("cfg_name", ["T0", "C2", "C5"])
def test_cofactor_inverse_allocates_the_same_units_as_eigens_inverse(cfg_name):
    """camera_pose = extrinsic.inverse(): open3d_order.c uses a float64 cofactor inverse, Eigen its own 4x4 kernel.
    They differ by a few ulp of float64; what matters is LocateVolumeUnit of every allocation sample, so the unit
    ranges floor((p -+ tau) / L) of all stride-4 samples of rendered frames are compared under both inverses (and under
    the rigid inverse [R^T | -R^T t] the CUDA kernels use)."""
    E = oracle.EigenOps()
    cfg = S.CONFIGS[cfg_name]
    L = cfg.voxel_size * 16
    tau = cfg.sdf_trunc
    worst = 0.0
    for i in range(0, min(cfg.n_frames, 40), 5):
        d, _, T = S.render_frame(cfg, i)
        inv_e = E.mat4d_inverse(T)
        inv_c = oracle.open3d_order_inverse4(T)
        inv_r = S.inv_T(T)
        worst = max(worst, float(np.abs(inv_e - inv_c).max()), float(np.abs(inv_e - inv_r).max()))
        dd = d[::4, ::4].astype(np.float64)
        rows, cols = np.nonzero((dd > 0) & (dd < cfg.depth_trunc))
        z = dd[rows, cols]
        x = (cols * 4.0 - cfg.cx) * z / cfg.fx
        y = (rows * 4.0 - cfg.cy) * z / cfg.fy
        units = []
        for P in (inv_e, inv_c, inv_r):
            p = np.stack([((P[r, 0] * x + P[r, 1] * y) + P[r, 2] * z) + P[r, 3] for r in range(3)], axis=1)
            units.append((np.floor((p - tau) / L).astype(np.int64), np.floor((p + tau) / L).astype(np.int64)))
        assert len(z) > 300
        for lo, hi in units[1:]:
            assert np.array_equal(lo, units[0][0]) and np.array_equal(hi, units[0][1])
    assert worst < 1e-12   # a few ulp of float64 on metre-scale poses
