// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product; never linked or
// loaded by pyslam_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library built from this file.
//
// Thin extern "C" harness around the UNMODIFIED reference semantic voxel-block grids
//   /root/reference/cpp/volumetric/voxel_block_semantic_grid.h(.hpp)  (VoxelBlockSemanticGridT)
//   /root/reference/cpp/volumetric/voxel_data_semantic.h              (voting :106-199, Bayesian :249-672)
// compiled from where the sources lie (nothing is copied into this repo) by oracle/Makefile into
// oracle/_ref/libref_semantic.so.  TBB is absent, so integrate takes the reference's sequential
// branch (voxel_block_grid.hpp:220-288): observations reach a voxel in input order.
#include "voxel_block_semantic_grid.h"

#include <cstdint>
#include <cstring>
#include <limits>

namespace {

using volumetric::BlockKeyHash;

template <typename Grid> class Dumpable : public Grid {
  public:
    using Grid::Grid;
    const auto &blocks() const { return this->blocks_; }
};

struct ISem {
    virtual ~ISem() = default;
    virtual void clear() = 0;
    virtual void integrate(const double *pts, int64_t n, const float *cols, const int *cls, const int *inst,
                           const float *depths) = 0;
    virtual void integrate_f32(const float *pts, int64_t n, const float *cols, const int *cls, const int *inst,
                               const float *depths) = 0;
    virtual int64_t num_blocks() const = 0;
    virtual int64_t dump(int32_t *keys, uint64_t *hashes, int32_t *count, double *pos, float *col, int32_t *obj,
                         int32_t *cls, float *conf, int32_t *aux, int K, int32_t *lab_obj, int32_t *lab_cls,
                         float *lab_logp) const = 0;
    virtual int64_t get_voxels(int min_count, float min_conf, double *pts, float *cols, int32_t *cls, int32_t *obj,
                               float *conf) const = 0;
    virtual int64_t assign(const volumetric::CameraFrustrum &fr, const cv::Mat &cls, const cv::Mat &inst,
                           const cv::Mat &depth, float thr, bool carve, float ratio, int min_votes, int32_t *ids,
                           int32_t *objs, int64_t cap) = 0;
    virtual void carve(const volumetric::CameraFrustrum &fr, const cv::Mat &depth, float thr) = 0;
    virtual int64_t query(const volumetric::CameraFrustrum *fr, const double *bb, int min_count, float min_conf,
                          double *pts, float *cols, int32_t *cls, int32_t *obj, float *conf) const = 0;
    virtual void integrate_segment(const double *pts, int64_t n, const float *cols, int cls, int obj) = 0;
    // flattened get_object_segments (by_class = 0) / get_class_segments (1): per segment {id, class id, n points,
    // conf min, conf max, OBB centre 3 + size 3 + quaternion wxyz (objects only)}; points / colours concatenated
    virtual int64_t segments(int by_class, int min_count, float min_conf, int32_t *ids, int32_t *cls, int64_t *npts,
                             float *conf_min, float *conf_max, double *obb, double *pts, float *cols,
                             int64_t *total_points) const = 0;
    virtual void remove_low_count(int min_count) = 0;
    virtual void remove_low_confidence(int min_confidence) = 0;
    virtual void merge_segments(int a, int b) = 0;
    virtual void remove_segment(int id) = 0;
};

template <typename V> constexpr bool is_prob = std::is_same_v<V, volumetric::VoxelSemanticDataProbabilistic>;

template <typename Grid, typename V> struct Sem final : ISem {
    Dumpable<Grid> g;
    Sem(double vs, int bs) : g(vs, bs) {}
    void clear() override { g.clear(); }
    template <typename Tp>
    void integrate_t(const Tp *pts, int64_t n, const float *cols, const int *cls, const int *inst,
                     const float *depths) {
        const size_t m = static_cast<size_t>(n);
        // the pybind entry (voxel_block_grid.hpp:12-112) dispatches on which arrays were given
        if (cls && inst && depths)
            g.template integrate_raw<Tp, float, int, int, float>(pts, m, cols, cls, inst, depths);
        else if (cls && inst)
            g.template integrate_raw<Tp, float, int, int>(pts, m, cols, cls, inst);
        else if (cls && depths)
            g.template integrate_raw<Tp, float, std::nullptr_t, int, float>(pts, m, cols, cls, nullptr, depths);
        else if (cls)
            g.template integrate_raw<Tp, float, std::nullptr_t, int>(pts, m, cols, cls);
        else
            g.template integrate_raw<Tp, float>(pts, m, cols);
    }
    void integrate(const double *pts, int64_t n, const float *cols, const int *cls, const int *inst,
                   const float *depths) override {
        integrate_t<double>(pts, n, cols, cls, inst, depths);
    }
    void integrate_f32(const float *pts, int64_t n, const float *cols, const int *cls, const int *inst,
                       const float *depths) override {
        integrate_t<float>(pts, n, cols, cls, inst, depths);
    }
    int64_t num_blocks() const override { return static_cast<int64_t>(g.num_blocks()); }
    int64_t dump(int32_t *keys, uint64_t *hashes, int32_t *count, double *pos, float *col, int32_t *obj,
                 int32_t *cls, float *conf, int32_t *aux, int K, int32_t *lab_obj, int32_t *lab_cls,
                 float *lab_logp) const override {
        const size_t nv = 512;
        int64_t b = 0;
        BlockKeyHash hasher;
        for (const auto &[key, block] : g.blocks()) {
            if (keys) {
                keys[3 * b + 0] = key.x;
                keys[3 * b + 1] = key.y;
                keys[3 * b + 2] = key.z;
            }
            if (hashes) hashes[b] = static_cast<uint64_t>(hasher(key));
            for (size_t i = 0; i < nv; ++i) {
                const auto &v = block.data[i];
                const size_t o = b * nv + i;
                if (count) count[o] = v.count;
                for (int a = 0; a < 3; ++a) {
                    if (pos) pos[3 * o + a] = v.position_sum[a];
                    if (col) col[3 * o + a] = v.color_sum[a];
                }
                if (obj) obj[o] = v.count ? v.get_object_id() : -1;
                if (cls) cls[o] = v.count ? v.get_class_id() : -1;
                if (conf) conf[o] = v.get_confidence();
                if constexpr (is_prob<V>) {
                    if (aux) aux[o] = static_cast<int32_t>(v.log_probabilities.size());
                    int k = 0;
                    for (const auto &[pair, lp] : v.log_probabilities) {  // std::map order = (object, class)
                        if (k >= K) break;
                        if (lab_obj) lab_obj[o * K + k] = pair.first;
                        if (lab_cls) lab_cls[o * K + k] = pair.second;
                        if (lab_logp) lab_logp[o * K + k] = lp;
                        ++k;
                    }
                    for (; k < K; ++k) {
                        if (lab_obj) lab_obj[o * K + k] = -1;
                        if (lab_cls) lab_cls[o * K + k] = -1;
                        if (lab_logp) lab_logp[o * K + k] = -std::numeric_limits<float>::infinity();
                    }
                } else {
                    if (aux) aux[o] = v.get_confidence_counter();
                }
            }
            ++b;
        }
        return b;
    }
    int64_t get_voxels(int min_count, float min_conf, double *pts, float *cols, int32_t *cls, int32_t *obj,
                       float *conf) const override {
        const auto out = g.get_voxels(min_count, min_conf);
        const int64_t n = static_cast<int64_t>(out.points.size());
        if (pts) std::memcpy(pts, out.points.data(), sizeof(double) * 3 * n);
        if (cols) std::memcpy(cols, out.colors.data(), sizeof(float) * 3 * n);
        if (cls) std::memcpy(cls, out.class_ids.data(), sizeof(int) * n);
        if (obj) std::memcpy(obj, out.object_ids.data(), sizeof(int) * n);
        if (conf) std::memcpy(conf, out.confidences.data(), sizeof(float) * n);
        return n;
    }
    int64_t assign(const volumetric::CameraFrustrum &fr, const cv::Mat &cls, const cv::Mat &inst, const cv::Mat &depth,
                   float thr, bool carve, float ratio, int min_votes, int32_t *ids, int32_t *objs,
                   int64_t cap) override {
        const auto m = g.assign_object_ids_to_instance_ids(fr, cls, inst, depth, thr, carve, ratio, min_votes);
        int64_t k = 0;
        for (const auto &[i, o] : m) {
            if (k < cap) {
                ids[k] = i;
                objs[k] = o;
            }
            ++k;
        }
        return k;
    }
    void carve(const volumetric::CameraFrustrum &fr, const cv::Mat &depth, float thr) override {
        g.carve(fr, depth, thr);
    }
    int64_t query(const volumetric::CameraFrustrum *fr, const double *bb, int min_count, float min_conf, double *pts,
                  float *cols, int32_t *cls, int32_t *obj, float *conf) const override {
        // IncludeSemantics = true: the variant that also returns labels (voxel_block_grid.h:125-137)
        const auto out =
            fr ? g.template get_voxels_in_camera_frustrum<true>(*fr, min_count, min_conf)
               : g.template get_voxels_in_bb<true>(volumetric::BoundingBox3D(bb[0], bb[1], bb[2], bb[3], bb[4], bb[5]),
                                                   min_count, min_conf);
        const int64_t n = static_cast<int64_t>(out.points.size());
        if (pts) std::memcpy(pts, out.points.data(), sizeof(double) * 3 * n);
        if (cols) std::memcpy(cols, out.colors.data(), sizeof(float) * 3 * n);
        if (cls) std::memcpy(cls, out.class_ids.data(), sizeof(int) * n);
        if (obj) std::memcpy(obj, out.object_ids.data(), sizeof(int) * n);
        if (conf) std::memcpy(conf, out.confidences.data(), sizeof(float) * n);
        return n;
    }
    void integrate_segment(const double *pts, int64_t n, const float *cols, int cls, int obj) override {
        g.template integrate_segment_raw<double, float, int, int>(pts, static_cast<size_t>(n), cols, cls, obj);
    }
    int64_t segments(int by_class, int min_count, float min_conf, int32_t *ids, int32_t *cls, int64_t *npts,
                     float *conf_min, float *conf_max, double *obb, double *pts, float *cols,
                     int64_t *total_points) const override {
        int64_t k = 0, off = 0;
        auto emit = [&](int id, int class_id, const auto &d, const volumetric::OrientedBoundingBox3D *box) {
            const int64_t n = static_cast<int64_t>(d.points.size());
            if (ids) ids[k] = id;
            if (cls) cls[k] = class_id;
            if (npts) npts[k] = n;
            if (conf_min) conf_min[k] = d.confidence_min;
            if (conf_max) conf_max[k] = d.confidence_max;
            if (obb && box) {
                for (int a = 0; a < 3; ++a) {
                    obb[10 * k + a] = box->center[a];
                    obb[10 * k + 3 + a] = box->size[a];
                }
                obb[10 * k + 6] = box->orientation.w();
                obb[10 * k + 7] = box->orientation.x();
                obb[10 * k + 8] = box->orientation.y();
                obb[10 * k + 9] = box->orientation.z();
            }
            if (pts) std::memcpy(pts + 3 * off, d.points.data(), sizeof(double) * 3 * n);
            if (cols) std::memcpy(cols + 3 * off, d.colors.data(), sizeof(float) * 3 * n);
            off += n;
            ++k;
        };
        if (by_class) {
            const auto grp = g.get_class_segments(min_count, min_conf);
            for (const auto &c : grp->class_vector) emit(c->class_id, c->class_id, *c, nullptr);
        } else {
            const auto grp = g.get_object_segments(min_count, min_conf);
            for (const auto &o : grp->object_vector) emit(o->object_id, o->class_id, *o, &o->oriented_bounding_box);
        }
        if (total_points) *total_points = off;
        return k;
    }
    void remove_low_count(int min_count) override { g.remove_low_count_voxels(min_count); }
    void remove_low_confidence(int min_confidence) override { g.remove_low_confidence_segments(min_confidence); }
    void merge_segments(int a, int b) override { g.merge_segments(a, b); }
    void remove_segment(int id) override { g.remove_segment(id); }
};

} // namespace

extern "C" {

// kind 0: VoxelBlockSemanticGrid (voting), 1: VoxelBlockSemanticProbabilisticGrid (voxel_block_semantic_grid.h:118-121)
void *refsem_create(int kind, double voxel_size, int block_size) {
    if (kind == 0)
        return new Sem<volumetric::VoxelBlockSemanticGrid, volumetric::VoxelSemanticData>(voxel_size, block_size);
    return new Sem<volumetric::VoxelBlockSemanticProbabilisticGrid, volumetric::VoxelSemanticDataProbabilistic>(
        voxel_size, block_size);
}

void refsem_destroy(void *h) { delete static_cast<ISem *>(h); }
void refsem_clear(void *h) { static_cast<ISem *>(h)->clear(); }

// class-static parameters (voxel_data_semantic.h:107-108, 251-254; set from Python,
// volumetric_integrator_voxel_semantic_grid.py:170-189)
void refsem_set_depth_threshold(int kind, float v) {
    if (kind == 0)
        volumetric::VoxelSemanticData::kDepthThreshold = v;
    else
        volumetric::VoxelSemanticDataProbabilistic::kDepthThreshold = v;
}
void refsem_set_depth_decay_rate(float v) { volumetric::VoxelSemanticDataProbabilistic::kDepthDecayRate = v; }
float refsem_get_depth_threshold(int kind) {
    return kind == 0 ? volumetric::VoxelSemanticData::kDepthThreshold
                     : volumetric::VoxelSemanticDataProbabilistic::kDepthThreshold;
}
float refsem_get_depth_decay_rate(void) { return volumetric::VoxelSemanticDataProbabilistic::kDepthDecayRate; }

void refsem_integrate(void *h, const double *pts, int64_t n, const float *cols, const int *cls, const int *inst,
                      const float *depths) {
    static_cast<ISem *>(h)->integrate(pts, n, cols, cls, inst, depths);
}

void refsem_integrate_f32(void *h, const float *pts, int64_t n, const float *cols, const int *cls, const int *inst,
                          const float *depths) {
    static_cast<ISem *>(h)->integrate_f32(pts, n, cols, cls, inst, depths);
}

int64_t refsem_num_blocks(void *h) { return static_cast<ISem *>(h)->num_blocks(); }

int64_t refsem_dump_blocks(void *h, int32_t *keys, uint64_t *hashes, int32_t *count, double *pos, float *col,
                           int32_t *obj, int32_t *cls, float *conf, int32_t *aux, int K, int32_t *lab_obj,
                           int32_t *lab_cls, float *lab_logp) {
    return static_cast<ISem *>(h)->dump(keys, hashes, count, pos, col, obj, cls, conf, aux, K, lab_obj, lab_cls,
                                        lab_logp);
}

int64_t refsem_get_voxels(void *h, int min_count, float min_conf, double *pts, float *cols, int32_t *cls,
                          int32_t *obj, float *conf) {
    return static_cast<ISem *>(h)->get_voxels(min_count, min_conf, pts, cols, cls, obj, conf);
}

static volumetric::CameraFrustrum sem_frustum(const float *K, int width, int height, const double *Tcw,
                                              float depth_max, float depth_min) {
    Eigen::Matrix4d T;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) T(r, c) = Tcw[4 * r + c];
    return volumetric::CameraFrustrum(K[0], K[1], K[2], K[3], width, height, T, depth_max, depth_min);
}

// VoxelBlockSemanticGridT::assign_object_ids_to_instance_ids (voxel_block_semantic_grid.hpp; the algorithm is
// voxel_semantic_data_association.h:69-373).  Images are row-major int32 / float32 HxW; depth may be NULL.
// Returns the number of (instance id -> object id) pairs; the first `cap` are written.
int64_t refsem_assign_object_ids(void *h, const float *K, int width, int height, const double *Tcw, float depth_max,
                                 float depth_min, const int32_t *class_image, const int32_t *instance_image,
                                 const float *depth_image, float depth_threshold, int do_carving,
                                 float min_vote_ratio, int min_votes, int32_t *ids, int32_t *objs, int64_t cap) {
    const auto fr = sem_frustum(K, width, height, Tcw, depth_max, depth_min);
    cv::Mat cls(height, width, CV_MAKETYPE(CV_32S, 1), const_cast<int32_t *>(class_image));
    cv::Mat inst(height, width, CV_MAKETYPE(CV_32S, 1), const_cast<int32_t *>(instance_image));
    cv::Mat depth;
    if (depth_image) depth = cv::Mat(height, width, CV_32FC1, const_cast<float *>(depth_image));
    return static_cast<ISem *>(h)->assign(fr, cls, inst, depth, depth_threshold, do_carving != 0, min_vote_ratio,
                                          min_votes, ids, objs, cap);
}

// VoxelBlockGridT::carve (voxel_block_grid.hpp:616-622; voxel_grid_carving.h:47-80) on a semantic grid
void refsem_carve(void *h, const float *K, int width, int height, const double *Tcw, float depth_max, float depth_min,
                  const float *depth_image, float depth_threshold) {
    const auto fr = sem_frustum(K, width, height, Tcw, depth_max, depth_min);
    cv::Mat depth(height, width, CV_32FC1, const_cast<float *>(depth_image));
    static_cast<ISem *>(h)->carve(fr, depth, depth_threshold);
}

// get_voxels_in_camera_frustrum / get_voxels_in_bb (voxel_block_grid.hpp:1019-1195, 822-1016); K == NULL selects the box
int64_t refsem_query(void *h, const float *K, int width, int height, const double *Tcw, float depth_max,
                     float depth_min, const double *bbox, int min_count, float min_conf, double *pts, float *cols,
                     int32_t *cls, int32_t *obj, float *conf) {
    if (K) {
        const auto fr = sem_frustum(K, width, height, Tcw, depth_max, depth_min);
        return static_cast<ISem *>(h)->query(&fr, nullptr, min_count, min_conf, pts, cols, cls, obj, conf);
    }
    return static_cast<ISem *>(h)->query(nullptr, bbox, min_count, min_conf, pts, cols, cls, obj, conf);
}

// process-wide object-id allocator (voxel_semantic_shared_data.h:27-33)
void refsem_set_next_object_id(int32_t v) { volumetric::VoxelSemanticSharedData::next_object_id.store(v); }
int32_t refsem_get_next_object_id(void) { return volumetric::VoxelSemanticSharedData::next_object_id.load(); }

void refsem_integrate_segment(void *h, const double *pts, int64_t n, const float *cols, int cls, int obj) {
    static_cast<ISem *>(h)->integrate_segment(pts, n, cols, cls, obj);
}
int64_t refsem_segments(void *h, int by_class, int min_count, float min_conf, int32_t *ids, int32_t *cls, int64_t *npts,
                        float *conf_min, float *conf_max, double *obb, double *pts, float *cols, int64_t *total_points) {
    return static_cast<ISem *>(h)->segments(by_class, min_count, min_conf, ids, cls, npts, conf_min, conf_max, obb, pts,
                                            cols, total_points);
}
void refsem_remove_low_count_voxels(void *h, int min_count) { static_cast<ISem *>(h)->remove_low_count(min_count); }
void refsem_remove_low_confidence_segments(void *h, int min_confidence) {
    static_cast<ISem *>(h)->remove_low_confidence(min_confidence);
}
void refsem_merge_segments(void *h, int a, int b) { static_cast<ISem *>(h)->merge_segments(a, b); }
void refsem_remove_segment(void *h, int id) { static_cast<ISem *>(h)->remove_segment(id); }

} // extern "C"
