// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product; never linked or
// loaded by pyslam_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library built from this file.
//
// Thin extern "C" harness around the UNMODIFIED reference implementation
//   /root/reference/cpp/volumetric/voxel_block_grid.h(.hpp)   (VoxelBlockGridT)
//   /root/reference/cpp/volumetric/voxel_hashing.h            (keys, hash, floor_div)
//   /root/reference/cpp/volumetric/voxel_data.h               (VoxelData running sums)
// compiled from where the sources lie (nothing is copied into this repo) by
// oracle/Makefile into oracle/_ref/libref_grid.so.  TBB is absent, so the
// reference takes its deterministic sequential branch
// (voxel_block_grid.hpp:457-461 -> integrate_raw_baseline :220-288).
#include "voxel_block_grid.h"

#include <chrono>
#include <cstdint>
#include <cstring>

namespace {

using volumetric::BlockKey;
using volumetric::BlockKeyHash;
using volumetric::LocalVoxelKey;
using volumetric::VoxelBlockGrid;
using volumetric::VoxelKey;

// blocks_ is protected (voxel_block_grid.h:224-233): a subclass may iterate it.
class DumpableGrid : public VoxelBlockGrid {
  public:
    using VoxelBlockGrid::VoxelBlockGrid;
    const auto &blocks() const { return blocks_; }
    float inv_voxel_size() const { return inv_voxel_size_; }
};

} // namespace

static volumetric::CameraFrustrum make_frustum(float fx, float fy, float cx, float cy, int width, int height,
                                               const double *Tcw, float depth_max, float depth_min) {
    Eigen::Matrix4d T;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) T(r, c) = Tcw[4 * r + c];
    return volumetric::CameraFrustrum(fx, fy, cx, cy, width, height, T, depth_max, depth_min);
}

template <typename Data> static int64_t copy_out(const Data &out, float *points, float *colors) {
    const int64_t n = static_cast<int64_t>(out.points.size());
    if (points) std::memcpy(points, out.points.data(), sizeof(float) * 3 * n);
    if (colors) std::memcpy(colors, out.colors.data(), sizeof(float) * 3 * n);
    return n;
}

extern "C" {

void *refgrid_create(float voxel_size, int block_size) {
    return new DumpableGrid(voxel_size, block_size);
}

void refgrid_destroy(void *h) { delete static_cast<DumpableGrid *>(h); }

void refgrid_clear(void *h) { static_cast<DumpableGrid *>(h)->clear(); }

// reference: VoxelBlockGridT::integrate_raw<float,float> (voxel_block_grid.hpp:115-136)
// cols may be NULL (points only).  Returns elapsed seconds of the reference call.
double refgrid_integrate(void *h, const float *pts, const float *cols, int64_t n) {
    auto *g = static_cast<DumpableGrid *>(h);
    const auto t0 = std::chrono::steady_clock::now();
    if (cols != nullptr) {
        g->integrate_raw<float, float>(pts, static_cast<size_t>(n), cols);
    } else {
        g->integrate_raw<float>(pts, static_cast<size_t>(n));
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// reference: the float64-points overload, VoxelBlockGridT::integrate_raw<double,float> (volumetric_grid_module.h:737-749)
double refgrid_integrate_f64(void *h, const double *pts, const float *cols, int64_t n) {
    auto *g = static_cast<DumpableGrid *>(h);
    const auto t0 = std::chrono::steady_clock::now();
    if (cols != nullptr) {
        g->integrate_raw<double, float>(pts, static_cast<size_t>(n), cols);
    } else {
        g->integrate_raw<double>(pts, static_cast<size_t>(n));
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

int64_t refgrid_num_blocks(void *h) {
    return static_cast<int64_t>(static_cast<DumpableGrid *>(h)->num_blocks());
}

int refgrid_block_size(void *h) { return static_cast<DumpableGrid *>(h)->get_block_size(); }

float refgrid_inv_voxel_size(void *h) { return static_cast<DumpableGrid *>(h)->inv_voxel_size(); }

// Dump every block: key int32[nb,3], hash u64[nb] (BlockKeyHash, voxel_hashing.h:106-113),
// count int32[nb,B^3], pos_sum f32[nb,B^3,3], col_sum f32[nb,B^3,3]; voxel order is the
// block's own flat index lx + ly*B + lz*B^2 (voxel_block.h:67-70).  Any output may be NULL.
int64_t refgrid_dump_blocks(void *h, int32_t *keys, uint64_t *hashes, int32_t *count,
                            float *pos_sum, float *col_sum) {
    auto *g = static_cast<DumpableGrid *>(h);
    const int B = g->get_block_size();
    const size_t nv = static_cast<size_t>(B) * B * B;
    int64_t b = 0;
    BlockKeyHash hasher;
    for (const auto &[key, block] : g->blocks()) {
        if (keys) {
            keys[3 * b + 0] = key.x;
            keys[3 * b + 1] = key.y;
            keys[3 * b + 2] = key.z;
        }
        if (hashes) hashes[b] = static_cast<uint64_t>(hasher(key));
        for (size_t i = 0; i < nv; ++i) {
            const auto &v = block.data[i];
            if (count) count[b * nv + i] = v.count;
            if (pos_sum) {
                pos_sum[(b * nv + i) * 3 + 0] = v.position_sum[0];
                pos_sum[(b * nv + i) * 3 + 1] = v.position_sum[1];
                pos_sum[(b * nv + i) * 3 + 2] = v.position_sum[2];
            }
            if (col_sum) {
                col_sum[(b * nv + i) * 3 + 0] = v.color_sum[0];
                col_sum[(b * nv + i) * 3 + 1] = v.color_sum[1];
                col_sum[(b * nv + i) * 3 + 2] = v.color_sum[2];
            }
        }
        ++b;
    }
    return b;
}

// reference: VoxelBlockGridT::get_voxels (voxel_block_grid.hpp:717-819).
// Two-call pattern: call with points == NULL to get the count.
int64_t refgrid_get_voxels(void *h, int min_count, float *points, float *colors,
                           double *elapsed_s) {
    auto *g = static_cast<DumpableGrid *>(h);
    const auto t0 = std::chrono::steady_clock::now();
    const auto out = g->get_voxels(min_count, 0.0f);
    const auto t1 = std::chrono::steady_clock::now();
    if (elapsed_s) *elapsed_s = std::chrono::duration<double>(t1 - t0).count();
    const int64_t n = static_cast<int64_t>(out.points.size());
    if (points) std::memcpy(points, out.points.data(), sizeof(float) * 3 * n);
    if (colors) std::memcpy(colors, out.colors.data(), sizeof(float) * 3 * n);
    return n;
}

void refgrid_remove_low_count_voxels(void *h, int min_count) {
    static_cast<DumpableGrid *>(h)->remove_low_count_voxels(min_count);
}

// reference: VoxelBlockGridT::carve (voxel_block_grid.hpp:616-622; voxel_grid_carving.h:47-80;
// camera_frustrum.cpp:174-196).  depth is a row-major f32 HxW image; Tcw row-major 4x4 f64.
void refgrid_carve(void *h, float fx, float fy, float cx, float cy, int width, int height,
                   const double *Tcw, float depth_max, float depth_min, const float *depth,
                   float depth_threshold) {
    auto *g = static_cast<DumpableGrid *>(h);
    Eigen::Matrix4d T;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) T(r, c) = Tcw[4 * r + c];
    volumetric::CameraFrustrum frustum(fx, fy, cx, cy, width, height, T, depth_max, depth_min);
    cv::Mat img(height, width, CV_32FC1, const_cast<float *>(depth));
    g->carve(frustum, img, depth_threshold);
}

// reference: VoxelBlockGridT::get_voxels_in_camera_frustrum (voxel_block_grid.hpp:1019-1195); two-call pattern
int64_t refgrid_get_voxels_in_frustum(void *h, float fx, float fy, float cx, float cy, int width, int height,
                                      const double *Tcw, float depth_max, float depth_min, int min_count,
                                      float *points, float *colors) {
    auto *g = static_cast<DumpableGrid *>(h);
    const auto fr = make_frustum(fx, fy, cx, cy, width, height, Tcw, depth_max, depth_min);
    return copy_out(g->get_voxels_in_camera_frustrum(fr, min_count, 0.0f), points, colors);
}

// reference: VoxelBlockGridT::get_voxels_in_bb (voxel_block_grid.hpp:822-1016); bbox = min xyz, max xyz
int64_t refgrid_get_voxels_in_bb(void *h, const double *bb, int min_count, float *points, float *colors) {
    auto *g = static_cast<DumpableGrid *>(h);
    const volumetric::BoundingBox3D box(bb[0], bb[1], bb[2], bb[3], bb[4], bb[5]);
    return copy_out(g->get_voxels_in_bb(box, min_count, 0.0f), points, colors);
}

// ---- leaf helpers straight from voxel_hashing.h, for known-answer tests ----

// get_voxel_key_inv<float,float> (voxel_hashing.h:69-75)
void ref_voxel_key_inv(float x, float y, float z, float inv_voxel_size, int32_t *out3) {
    const VoxelKey k = volumetric::get_voxel_key_inv<float, float>(x, y, z, inv_voxel_size);
    out3[0] = k.x;
    out3[1] = k.y;
    out3[2] = k.z;
}

// floor_div (voxel_hashing.h:139-142)
int64_t ref_floor_div(int64_t a, int64_t b) { return volumetric::floor_div(a, b); }

// get_block_key + get_local_voxel_key (voxel_hashing.h:145-161)
void ref_block_and_local_key(const int32_t *voxel3, int block_size, int32_t *block3,
                             int32_t *local3) {
    const VoxelKey vk(voxel3[0], voxel3[1], voxel3[2]);
    const BlockKey bk = volumetric::get_block_key(vk, static_cast<size_t>(block_size));
    const LocalVoxelKey lk = volumetric::get_local_voxel_key(vk, bk, block_size);
    block3[0] = bk.x;
    block3[1] = bk.y;
    block3[2] = bk.z;
    local3[0] = lk.x;
    local3[1] = lk.y;
    local3[2] = lk.z;
}

// BlockKeyHash (voxel_hashing.h:106-113)
uint64_t ref_block_key_hash(int32_t x, int32_t y, int32_t z) {
    return static_cast<uint64_t>(BlockKeyHash{}(BlockKey(x, y, z)));
}

int ref_sizeof_voxel_data(void) { return static_cast<int>(sizeof(volumetric::VoxelData)); }

} // extern "C"
