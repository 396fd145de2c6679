// TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product.
//
// eigen_ops.cpp - the three Eigen expressions of Open3D's TSDF path whose evaluation order the published source does
// not spell out, evaluated by REAL Eigen (the copy vendored in the reference tree, thirdparty/lietorch/eigen; Open3D
// 0.19 builds against Eigen 3.4), so that oracle/open3d_order.c's restatement of them can be pinned:
//   UniformTSDFVolume.cpp   Eigen::Vector4f pt_camera = extrinsic_f * pt_3d_homo;            eig_mat4f_times_vec4f
//   PointCloudFactory.cpp   Eigen::Matrix4d camera_pose = extrinsic.inverse();               eig_mat4d_inverse
//                           Eigen::Vector4d point = camera_pose * Eigen::Vector4d(x,y,z,1);  eig_mat4d_times_vec4d
// Built by oracle/Makefile for the x86-64 baseline (SSE2, no FMA), the instruction set of Open3D's published wheels.
// Matrices cross the C ABI row-major.
#include <Eigen/Core>
#include <Eigen/LU>

extern "C" {

void eig_mat4f_times_vec4f(const float *m_rowmajor, const float *v, float *out) {
    Eigen::Matrix4f M;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) M(i, j) = m_rowmajor[4 * i + j];
    const Eigen::Vector4f x(v[0], v[1], v[2], v[3]);
    const Eigen::Vector4f r = M * x;
    for (int i = 0; i < 4; ++i) out[i] = r(i);
}

void eig_mat4d_times_vec4d(const double *m_rowmajor, const double *v, double *out) {
    Eigen::Matrix4d M;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) M(i, j) = m_rowmajor[4 * i + j];
    const Eigen::Vector4d x(v[0], v[1], v[2], v[3]);
    const Eigen::Vector4d r = M * x;
    for (int i = 0; i < 4; ++i) out[i] = r(i);
}

void eig_mat4d_inverse(const double *m_rowmajor, double *out_rowmajor) {
    Eigen::Matrix4d M;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) M(i, j) = m_rowmajor[4 * i + j];
    const Eigen::Matrix4d inv = M.inverse();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out_rowmajor[4 * i + j] = inv(i, j);
}

int eig_version(void) { return EIGEN_WORLD_VERSION * 10000 + EIGEN_MAJOR_VERSION * 100 + EIGEN_MINOR_VERSION; }
}
