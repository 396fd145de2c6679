// b2v_api.cu — the C ABI (include/b2v.h): volume / grid lifetime, frame staging and stream
// pipelining, parity hooks.  Host code only; kernels live in b2v_tsdf.cu, b2v_mesh.cu, b2v_grid.cu.
//
// Per frame (b2v_integrate):   copy stream:    H2D depth, colour  -> event ready[s]
//                              compute stream: wait ready[s]; allocate_kernel; integrate_kernel;
//                                              event free[s]
// with a ring of kStage device staging slots, so the upload of frame f+1 overlaps the kernels of
// frame f.  Nothing synchronises with the host until b2v_synchronize / an inspection call.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b2v.h"
#include "b2v_internal.h"

using namespace b2v;

namespace b2v {

// Host-side pose algebra, same operation order as the oracle (and -ffp-contract=off on the host
// compiler), so allocation keys agree bit for bit.
void fill_frame_params(FrameParams *p, const double K[4], const double Tcw[16], int H, int W,
                       const VolumeGeometry &g, uint32_t frame_id) {
    p->fx = K[0];
    p->fy = K[1];
    p->cx = K[2];
    p->cy = K[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) p->pose.Rwc[3 * i + j] = Tcw[4 * j + i];
    for (int i = 0; i < 3; ++i)
        p->pose.twc[i] = -((p->pose.Rwc[3 * i + 0] * Tcw[3] + p->pose.Rwc[3 * i + 1] * Tcw[7]) +
                           p->pose.Rwc[3 * i + 2] * Tcw[11]);
    p->tau_d = g.unit_shift > 0 ? g.tau_d : static_cast<double>(g.tau);
    p->unit_len = g.voxel_length * static_cast<double>(kB << g.unit_shift);
    IntFrame &I = p->I;
    for (int i = 0; i < 12; ++i) I.E[i] = static_cast<float>(Tcw[i]);
    for (int i = 0; i < 3; ++i) I.Es[i] = I.E[4 * i + 2] * g.vs;  // extrinsic_f * voxel_length_f, column 2
    I.fxf = static_cast<float>(K[0]);
    I.fyf = static_cast<float>(K[1]);
    I.cxf = static_cast<float>(K[2]);
    I.cyf = static_cast<float>(K[3]);
    I.safe_w = static_cast<float>(W) - 0.0001f;
    I.safe_h = static_cast<float>(H) - 0.0001f;
    I.tau = g.tau;
    I.inv_tau = 1.0f / g.tau;
    I.W = W;
    I.tex = nullptr;
    p->inv_fx = 1.0f / I.fxf;
    p->inv_fy = 1.0f / I.fyf;
    p->inv_vs = 1.0f / g.vs;  // voxel_block_grid.hpp:6
    p->depth_trunc = g.depth_trunc;
    p->unit_shift = g.unit_shift;
    p->H = H;
    p->W = W;
    p->stride = g.stride;
    p->frame_id = frame_id;
    p->shard_rank = g.shard_rank;
    p->shard_count = g.shard_count;
    p->group_bit = -1;
    p->group_buf = 0;
}

VolumeConsts volume_consts(const VolumeGeometry &g) {
    VolumeConsts c;
    c.unit_len = g.voxel_length * static_cast<double>(kB << g.unit_shift);
    c.vs = g.vs;
    c.half_vs = g.vs * 0.5f;
    c.unit_shift = g.unit_shift;
    return c;
}

}  // namespace b2v

namespace {

static_assert(kCtrNew0 == kCtrActive0 + kActiveRing, "the ring counters are cleared with one memset");
constexpr int kFrameStage = 4;                       // staging ring of the frame-by-frame path
constexpr int kGroupStage = kGroupBufs * kMaxGroup;  // group buffers x kMaxGroup frames
constexpr int kStage = kGroupStage + kFrameStage;    // raw-frame staging slots (device copies of host frames)

uint32_t next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return static_cast<uint32_t>(p);
}

bool is_device_pointer(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

struct b2v_volume {
    b2v_config cfg{};
    VolumeGeometry geo{};
    cudaStream_t compute = nullptr, copy = nullptr, alloc = nullptr;
    cudaStream_t last_stream = nullptr;  // caller stream of the most recent frame (synchronised on reads)
    bool overlap = true;                 // allocate(f+1) on its own stream, concurrent with integrate(f)
    bool use_tma = true;                 // stage image tiles with TMA when the layout allows it
    bool inputs_fenced = false;          // batch call: device inputs already ordered before the alloc stream
    bool fuse = true;                    // b2v_integrate_batch fuses groups of up to kMaxGroup frames
    int group_frames = 16;               // frames per fused group (1..kMaxGroup), b2v_set_group_size
    LambdaMap lam_map{};
    bool lam_map_ok = false;
    cudaEvent_t input_event = nullptr;   // b2v_set_input_event: readiness of the next batch's device inputs
    bool rings_stale = false;            // a fused batch advanced frame_id: the per-frame ring counters must be re-armed
    float4 *d_gtex[kGroupBufs * kMaxGroup] = {};  // texel images of the group buffers
    size_t gtex_pixels = 0;
    uint32_t group_id = 0;
    cudaEvent_t ev_galloc[kGroupBufs] = {}, ev_group_done[kGroupBufs] = {};
    int last_group_buf = -1, last_group_count = 0;  // most recent frame came from a fused group
    int64_t prof_frames = 0, prof_int_launches = 0;
    // optional rectification stage (b2v_set_rectification)
    float *d_mapx = nullptr, *d_mapy = nullptr;
    int rect_H = 0, rect_W = 0, rect_swap = 0;
    float *d_rdepth[kStage] = {};    // rectified frames (same slot layout as the raw staging)
    uint8_t *d_rcolor[kStage] = {};
    // TMA descriptors are cached per image address (encoding costs ~1 us of host time each)
    std::unordered_map<uintptr_t, FrameMaps> map_cache;
    int map_H = 0, map_W = 0;
    const float *map_lam = nullptr;
    cudaEvent_t ev_in = nullptr, ev_alloc_done[kActiveRing] = {}, ev_int_done[kActiveRing] = {};
    // raw 16-bit depth input (b2v_integrate_u16 / b2v_integrate_batch_u16): uploaded as is, widened on the device
    uint16_t *d_depth16[kStage] = {};   // same slot layout as d_depth; allocated on first use
    size_t stage16_pixels = 0;
    float in_u16_scale = 0.0f;          // > 0 while a *_u16 entry point runs: `depth` pointers are uint16_t
    float *d_depth[kStage] = {};
    uint8_t *d_color[kStage] = {};
    float4 *d_texel[kStage] = {};   // packed {depth, lambda, rgbx} frames read by integrate_kernel
    float *d_lambda = nullptr;      // lambda image of the cached intrinsics
    double lam_K[4] = {0, 0, 0, 0};
    int lam_H = 0, lam_W = 0;
    size_t stage_pixels = 0;
    cudaEvent_t ev_ready[kStage] = {}, ev_free[kStage] = {};
    HashTable table{};
    PoolMeta meta{};
    uint32_t frame_id = 0;  // frames integrated since reset (stamp = frame_id + 1)
    int grid_ctas = 0;
    int sm_count = 0;
    int64_t launches = 0;
    uint32_t *h_counters = nullptr;  // pinned mirror
    std::string err;
    // mesh / point-cloud extraction
    MeshBuffers mb{};
    uint32_t mesh_blocks_cap = 0;
    size_t mesh_v_cap = 0, mesh_t_cap = 0;
    int64_t last_nv = 0, last_nt = 0;
    uint32_t *h_totals = nullptr;
    // optional per-kernel timing (b2v_profile_*)
    bool prof_enabled = false;
    std::vector<cudaEvent_t> prof_events;  // quadruples: allocate begin/end, integrate begin/end
    size_t prof_used = 0;
};

#define B2V_CUDA(v, call)                                                                  \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            (v)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            return B2V_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

static int volume_clear_device(b2v_volume *v) {
    const size_t tcap = static_cast<size_t>(v->table.mask) + 1;
    B2V_CUDA(v, cudaMemsetAsync(v->table.entries, 0xFF, tcap * sizeof(uint4), v->compute));
    B2V_CUDA(v, cudaMemsetAsync(v->table.stamp, 0, tcap * sizeof(uint32_t), v->compute));
    B2V_CUDA(v, cudaMemsetAsync(v->meta.group_mask, 0, tcap * kGroupBufs * sizeof(uint32_t), v->compute));
    B2V_CUDA(v, cudaMemsetAsync(v->meta.counters, 0, kNumCounters * sizeof(uint32_t), v->compute));
    return B2V_OK;
}

extern "C" int b2v_version(void) { return 100; }

extern "C" int b2v_selftest_division(int32_t device, uint64_t pairs, uint64_t *bad_reciprocals, uint64_t *bad_quotients) {
    if (cudaSetDevice(device) != cudaSuccess) return B2V_ERR_CUDA;
    unsigned long long *d = nullptr, h[2] = {0, 0};
    if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) return B2V_ERR_CUDA;
    cudaError_t e = cudaMemset(d, 0, sizeof(h));
    if (e == cudaSuccess) e = launch_selftest_division(d, pairs, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return B2V_ERR_CUDA;
    if (bad_reciprocals) *bad_reciprocals = h[0];
    if (bad_quotients) *bad_quotients = h[1];
    return B2V_OK;
}

extern "C" int b2v_device_sm_count(int32_t device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    return n;
}

extern "C" const char *b2v_last_error(const b2v_volume *v) { return v ? v->err.c_str() : "null volume"; }

extern "C" int b2v_create(const b2v_config *cfg, b2v_volume **out) {
    if (!cfg || !out) return B2V_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (cfg->block_size != B2V_BLOCK_SIZE || !(cfg->voxel_size > 0.0f) || !(cfg->sdf_trunc > 0.0f) ||
        !(cfg->depth_trunc > 0.0f) || cfg->capacity_blocks == 0 || cfg->shard_count < 1 ||
        cfg->shard_rank < 0 || cfg->shard_rank >= cfg->shard_count ||
        (cfg->unit_resolution != 0 && cfg->unit_resolution != 8 && cfg->unit_resolution != 16) ||
        static_cast<float>(cfg->voxel_length > 0.0 ? cfg->voxel_length : cfg->voxel_size) != cfg->voxel_size ||
        static_cast<float>(cfg->sdf_trunc_d > 0.0 ? cfg->sdf_trunc_d : cfg->sdf_trunc) != cfg->sdf_trunc)
        return B2V_ERR_INVALID_ARGUMENT;
    // allocate_kernel packs 21 bits per axis inside one frustum
    if (static_cast<double>(cfg->depth_trunc) / (static_cast<double>(cfg->voxel_size) * kB) > 5.0e5)
        return B2V_ERR_UNSUPPORTED;
    b2v_volume *v = new (std::nothrow) b2v_volume();
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    v->cfg = *cfg;
    if (v->cfg.depth_stride < 1) v->cfg.depth_stride = 4;
    if (v->cfg.unit_resolution == 0) v->cfg.unit_resolution = 16;
    if (!(v->cfg.voxel_length > 0.0)) v->cfg.voxel_length = static_cast<double>(cfg->voxel_size);
    if (!(v->cfg.sdf_trunc_d > 0.0)) v->cfg.sdf_trunc_d = static_cast<double>(cfg->sdf_trunc);
    v->geo.vs = cfg->voxel_size;
    v->geo.tau = cfg->sdf_trunc;
    v->geo.depth_trunc = cfg->depth_trunc;
    v->geo.voxel_length = v->cfg.voxel_length;
    v->geo.tau_d = v->cfg.sdf_trunc_d;
    v->geo.unit_shift = v->cfg.unit_resolution == 16 ? 1 : 0;
    v->geo.stride = v->cfg.depth_stride;
    v->geo.shard_rank = cfg->shard_rank;
    v->geo.shard_count = cfg->shard_count;
    *out = v;  // returned even on CUDA failure so the caller can read b2v_last_error and destroy
    B2V_CUDA(v, cudaSetDevice(cfg->device));
    B2V_CUDA(v, cudaStreamCreateWithFlags(&v->compute, cudaStreamNonBlocking));
    B2V_CUDA(v, cudaStreamCreateWithFlags(&v->copy, cudaStreamNonBlocking));
    B2V_CUDA(v, cudaStreamCreateWithFlags(&v->alloc, cudaStreamNonBlocking));
    B2V_CUDA(v, cudaEventCreateWithFlags(&v->ev_in, cudaEventDisableTiming));
    for (int r = 0; r < kActiveRing; ++r) {
        B2V_CUDA(v, cudaEventCreateWithFlags(&v->ev_alloc_done[r], cudaEventDisableTiming));
        B2V_CUDA(v, cudaEventCreateWithFlags(&v->ev_int_done[r], cudaEventDisableTiming));
    }
    if (const char *e = std::getenv("B2V_OVERLAP")) v->overlap = std::atoi(e) != 0;
    if (const char *e = std::getenv("B2V_TMA")) v->use_tma = std::atoi(e) != 0;
    if (const char *e = std::getenv("B2V_FUSE")) v->fuse = std::atoi(e) != 0;
    if (const char *e = std::getenv("B2V_GROUP")) {
        const int n = std::atoi(e);
        if (n >= 1 && n <= kMaxGroup) v->group_frames = n;
    }
    for (int b = 0; b < kGroupBufs; ++b) {
        B2V_CUDA(v, cudaEventCreateWithFlags(&v->ev_galloc[b], cudaEventDisableTiming));
        B2V_CUDA(v, cudaEventCreateWithFlags(&v->ev_group_done[b], cudaEventDisableTiming));
    }
    for (int s = 0; s < kStage; ++s) {
        B2V_CUDA(v, cudaEventCreateWithFlags(&v->ev_ready[s], cudaEventDisableTiming));
        B2V_CUDA(v, cudaEventCreateWithFlags(&v->ev_free[s], cudaEventDisableTiming));
    }
    const uint32_t cap = cfg->capacity_blocks;
    const uint32_t tcap = next_pow2(static_cast<uint64_t>(cap) * 2);
    v->table.mask = tcap - 1;
    v->meta.capacity = cap;
    B2V_CUDA(v, cudaMalloc(&v->table.entries, static_cast<size_t>(tcap) * sizeof(uint4)));
    B2V_CUDA(v, cudaMalloc(&v->table.stamp, static_cast<size_t>(tcap) * sizeof(uint32_t)));
    B2V_CUDA(v, cudaMalloc(&v->meta.pool, static_cast<size_t>(cap) * kBlockFloats * sizeof(float)));
    B2V_CUDA(v, cudaMalloc(&v->meta.block_keys, static_cast<size_t>(cap) * sizeof(int4)));
    B2V_CUDA(v, cudaMalloc(&v->meta.counters, kNumCounters * sizeof(uint32_t)));
    B2V_CUDA(v, cudaMalloc(&v->meta.active_slots, static_cast<size_t>(cap) * kActiveRing * sizeof(uint32_t)));
    B2V_CUDA(v, cudaMalloc(&v->meta.group_mask, static_cast<size_t>(tcap) * kGroupBufs * sizeof(uint32_t)));
    B2V_CUDA(v, cudaMalloc(&v->meta.union_slots, static_cast<size_t>(cap) * kGroupBufs * sizeof(uint32_t)));
    B2V_CUDA(v, cudaMalloc(&v->meta.block_flags, static_cast<size_t>(cap) * sizeof(uint32_t)));
    B2V_CUDA(v, cudaMemsetAsync(v->meta.block_flags, 0, static_cast<size_t>(cap) * sizeof(uint32_t), v->compute));
    B2V_CUDA(v, cudaMallocHost(&v->h_counters, kNumCounters * sizeof(uint32_t)));
    B2V_CUDA(v, cudaMallocHost(&v->h_totals, kNumMeshTotals * sizeof(uint32_t)));
    std::memset(v->h_totals, 0, kNumMeshTotals * sizeof(uint32_t));
    B2V_CUDA(v, cudaMemsetAsync(v->meta.pool, 0, static_cast<size_t>(cap) * kBlockFloats * sizeof(float),
                                v->compute));
    int rc = volume_clear_device(v);
    if (rc != B2V_OK) return rc;
    const int sms = b2v_device_sm_count(cfg->device);
    // persistent grid: exactly one wave of resident CTAs (B2V_INT_CTAS_PER_SM overrides, for tuning)
    int per_sm = integrate_max_resident_ctas_per_sm();
    if (const char *e = std::getenv("B2V_INT_CTAS_PER_SM")) {
        const int n = std::atoi(e);
        if (n >= 1 && n <= 32) per_sm = n;
    }
    v->grid_ctas = (sms > 0 ? sms : 148) * per_sm;
    v->sm_count = sms;
    B2V_CUDA(v, cudaStreamSynchronize(v->compute));
    return B2V_OK;
}

extern "C" int b2v_destroy(b2v_volume *v) {
    if (!v) return B2V_OK;
    cudaSetDevice(v->cfg.device);
    cudaDeviceSynchronize();
    if (v->ev_in) cudaEventDestroy(v->ev_in);
    for (int r = 0; r < kActiveRing; ++r) {
        if (v->ev_alloc_done[r]) cudaEventDestroy(v->ev_alloc_done[r]);
        if (v->ev_int_done[r]) cudaEventDestroy(v->ev_int_done[r]);
    }
    cudaFree(v->d_depth[0]);  // slots 1.. point into the same two allocations
    cudaFree(v->d_color[0]);
    for (int s = 0; s < kStage; ++s) {
        cudaFree(v->d_texel[s]);
        if (v->ev_ready[s]) cudaEventDestroy(v->ev_ready[s]);
        if (v->ev_free[s]) cudaEventDestroy(v->ev_free[s]);
    }
    cudaFree(v->d_lambda);
    cudaFree(v->d_mapx);
    cudaFree(v->d_mapy);
    cudaFree(v->d_rdepth[0]);
    cudaFree(v->d_rcolor[0]);
    cudaFree(v->d_depth16[0]);
    for (float4 *t : v->d_gtex) cudaFree(t);
    for (int b = 0; b < kGroupBufs; ++b) {
        if (v->ev_galloc[b]) cudaEventDestroy(v->ev_galloc[b]);
        if (v->ev_group_done[b]) cudaEventDestroy(v->ev_group_done[b]);
    }
    cudaFree(v->meta.group_mask);
    cudaFree(v->meta.union_slots);
    cudaFree(v->meta.block_flags);
    cudaFree(v->table.entries);
    cudaFree(v->table.stamp);
    cudaFree(v->meta.pool);
    cudaFree(v->meta.block_keys);
    cudaFree(v->meta.counters);
    cudaFree(v->meta.active_slots);
    cudaFree(v->mb.nbr);
    cudaFree(v->mb.cube);
    cudaFree(v->mb.edge_mask);
    cudaFree(v->mb.local);
    cudaFree(v->mb.sums);
    cudaFree(v->mb.offs);
    cudaFree(v->mb.work);
    cudaFree(v->mb.totals);
    cudaFree(v->mb.partials);
    cudaFree(v->mb.vertices);
    cudaFree(v->mb.colors);
    cudaFree(v->mb.edge_ids);
    cudaFree(v->mb.triangles);
    cudaFreeHost(v->h_counters);
    cudaFreeHost(v->h_totals);
    for (cudaEvent_t e : v->prof_events)
        if (e) cudaEventDestroy(e);
    if (v->compute) cudaStreamDestroy(v->compute);
    if (v->copy) cudaStreamDestroy(v->copy);
    if (v->alloc) cudaStreamDestroy(v->alloc);
    delete v;
    return B2V_OK;
}

static int read_counters(b2v_volume *v) {
    B2V_CUDA(v, cudaSetDevice(v->cfg.device));
    B2V_CUDA(v, cudaStreamSynchronize(v->copy));
    B2V_CUDA(v, cudaStreamSynchronize(v->alloc));
    if (v->last_stream) B2V_CUDA(v, cudaStreamSynchronize(v->last_stream));
    B2V_CUDA(v, cudaMemcpyAsync(v->h_counters, v->meta.counters, kNumCounters * sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, v->compute));
    B2V_CUDA(v, cudaStreamSynchronize(v->compute));
    if (v->h_counters[kCtrError]) {
        v->err = (v->h_counters[kCtrError] & 2u) ? "hash table full: raise capacity_blocks"
                                                 : "block pool full: raise capacity_blocks";
        return B2V_ERR_CAPACITY;
    }
    return B2V_OK;
}

static uint32_t block_count(const b2v_volume *v) {
    const uint32_t n = v->h_counters[kCtrPool];
    return n < v->meta.capacity ? n : v->meta.capacity;
}

extern "C" int b2v_reset(b2v_volume *v) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    int rc = read_counters(v);
    if (rc == B2V_ERR_CUDA) return rc;
    const uint32_t nb = block_count(v);
    B2V_CUDA(v, cudaMemsetAsync(v->meta.pool, 0, static_cast<size_t>(nb) * kBlockFloats * sizeof(float),
                                v->compute));
    B2V_CUDA(v, cudaMemsetAsync(v->meta.block_flags, 0, static_cast<size_t>(nb) * sizeof(uint32_t), v->compute));
    rc = volume_clear_device(v);
    if (rc != B2V_OK) return rc;
    v->frame_id = 0;
    v->group_id = 0;
    v->last_group_buf = -1;
    v->rings_stale = false;
    v->err.clear();
    B2V_CUDA(v, cudaStreamSynchronize(v->compute));
    return B2V_OK;
}

static int ensure_staging(b2v_volume *v, size_t pixels) {
    if (pixels <= v->stage_pixels) return B2V_OK;
    B2V_CUDA(v, cudaStreamSynchronize(v->compute));
    B2V_CUDA(v, cudaStreamSynchronize(v->copy));
    B2V_CUDA(v, cudaStreamSynchronize(v->alloc));
    if (v->last_stream) B2V_CUDA(v, cudaStreamSynchronize(v->last_stream));
    // raw staging slots are carved out of two contiguous allocations, so the frames of a group (which
    // are contiguous in the caller's arrays) upload with ONE copy per image type
    cudaFree(v->d_depth[0]);
    cudaFree(v->d_color[0]);
    for (int s = 0; s < kStage; ++s) {
        cudaFree(v->d_texel[s]);
        v->d_depth[s] = nullptr;
        v->d_color[s] = nullptr;
        v->d_texel[s] = nullptr;
    }
    v->stage_pixels = 0;  // stays 0 if an allocation below fails
    float *dbase = nullptr;
    uint8_t *cbase = nullptr;
    B2V_CUDA(v, cudaMalloc(&dbase, pixels * sizeof(float) * kStage));
    v->d_depth[0] = dbase;
    B2V_CUDA(v, cudaMalloc(&cbase, pixels * 3 * kStage));
    v->d_color[0] = cbase;
    for (int s = 0; s < kStage; ++s) {
        v->d_depth[s] = dbase + pixels * s;
        v->d_color[s] = cbase + pixels * 3 * s;
        if (s >= kGroupStage)  // texel images of the per-frame path (the group buffers have their own)
            B2V_CUDA(v, cudaMalloc(&v->d_texel[s], pixels * sizeof(float4)));
    }
    cudaFree(v->d_lambda);
    v->d_lambda = nullptr;
    B2V_CUDA(v, cudaMalloc(&v->d_lambda, pixels * sizeof(float)));
    v->lam_H = v->lam_W = 0;
    v->stage_pixels = pixels;
    return B2V_OK;
}

static int grow_profile_events(b2v_volume *v, size_t need) {
    if (v->prof_used + need > v->prof_events.size()) {
        const size_t old = v->prof_events.size();
        v->prof_events.resize(old + 4 * 256, nullptr);
        for (size_t k = old; k < v->prof_events.size(); ++k) B2V_CUDA(v, cudaEventCreate(&v->prof_events[k]));
    }
    return B2V_OK;
}

// cached TMA descriptors of a frame (keyed by the depth image address; colour address is checked)
static const FrameMaps *frame_maps(b2v_volume *v, const float *d_depth, const uint8_t *d_color, int H, int W) {
    if (!v->use_tma || !tma_tiles_usable(W, v->cfg.depth_stride, d_depth, d_color, v->d_lambda)) return nullptr;
    if (v->map_H != H || v->map_W != W || v->map_lam != v->d_lambda || v->map_cache.size() > 4096) {
        v->map_cache.clear();
        v->map_H = H;
        v->map_W = W;
        v->map_lam = v->d_lambda;
        v->lam_map_ok = encode_lambda_map(&v->lam_map, v->d_lambda, H, W, 32);
    }
    if (!v->lam_map_ok) return nullptr;
    auto it = v->map_cache.find(reinterpret_cast<uintptr_t>(d_depth));
    if (it != v->map_cache.end() && it->second.color_ptr == d_color) return &it->second;
    FrameMaps m;
    if (!encode_frame_maps(&m, d_depth, d_color, H, W, 32)) return nullptr;
    m.color_ptr = d_color;
    auto res = v->map_cache.insert_or_assign(reinterpret_cast<uintptr_t>(d_depth), m);
    return &res.first->second;
}

static int rectify_frame(b2v_volume *v, const float **d_depth, const uint8_t **d_color, int H, int W, int slot,
                         cudaStream_t stream);

// raw uint16 staging, one contiguous allocation carved into the same slots as the float staging
static int ensure_staging16(b2v_volume *v, size_t pixels) {
    if (pixels <= v->stage16_pixels) return B2V_OK;
    B2V_CUDA(v, cudaStreamSynchronize(v->compute));
    B2V_CUDA(v, cudaStreamSynchronize(v->copy));
    B2V_CUDA(v, cudaStreamSynchronize(v->alloc));
    if (v->last_stream) B2V_CUDA(v, cudaStreamSynchronize(v->last_stream));
    cudaFree(v->d_depth16[0]);
    for (int s = 0; s < kStage; ++s) v->d_depth16[s] = nullptr;
    v->stage16_pixels = 0;
    uint16_t *base = nullptr;
    B2V_CUDA(v, cudaMalloc(&base, pixels * sizeof(uint16_t) * kStage));
    for (int s = 0; s < kStage; ++s) v->d_depth16[s] = base + pixels * s;
    v->stage16_pixels = pixels;
    return B2V_OK;
}

static int integrate_frame(b2v_volume *v, const float *depth, const uint8_t *color, int32_t height,
                           int32_t width, const double K[4], const double Tcw[16], void *stream,
                           int dev_hint = -1) {
    if (!depth || !color || !K || !Tcw || height <= 0 || width <= 0) {
        v->err = "b2v_integrate: null pointer or non-positive image size";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    if (!(K[0] > 0.0) || !(K[1] > 0.0)) {
        v->err = "b2v_integrate: focal lengths must be positive";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    const size_t pixels = static_cast<size_t>(height) * width;
    bool dev_depth, dev_color;
    if (dev_hint >= 0) {  // batch call: queried once for the whole batch
        dev_depth = dev_color = dev_hint != 0;
    } else {
        B2V_CUDA(v, cudaSetDevice(v->cfg.device));
        dev_depth = is_device_pointer(depth);
        dev_color = is_device_pointer(color);
    }
    if (stream != nullptr && !(dev_depth && dev_color)) {
        v->err = "b2v_integrate: a caller stream requires device image pointers";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t cs = stream ? static_cast<cudaStream_t>(stream) : v->compute;
    cudaStream_t as = v->overlap ? v->alloc : cs;  // stream of the allocate kernel
    v->last_stream = stream ? cs : nullptr;
    const int s = kGroupStage + static_cast<int>(v->frame_id % kFrameStage);
    const int ring = static_cast<int>(v->frame_id % kActiveRing);
    const float *d_depth = depth;
    const uint8_t *d_color = color;
    const bool staged = !(dev_depth && dev_color);
    const bool u16 = v->in_u16_scale > 0.0f;  // `depth` is really const uint16_t *
    {
        int rc = ensure_staging(v, pixels);
        if (rc == B2V_OK && u16) rc = ensure_staging16(v, pixels);
        if (rc != B2V_OK) return rc;
    }
    if (staged) {
        B2V_CUDA(v, cudaStreamWaitEvent(v->copy, v->ev_free[s], 0));
        if (!dev_depth) {
            if (u16)
                B2V_CUDA(v, cudaMemcpyAsync(v->d_depth16[s], depth, pixels * sizeof(uint16_t), cudaMemcpyHostToDevice,
                                            v->copy));
            else
                B2V_CUDA(v, cudaMemcpyAsync(v->d_depth[s], depth, pixels * sizeof(float),
                                            cudaMemcpyHostToDevice, v->copy));
            d_depth = v->d_depth[s];
        }
        if (!dev_color) {
            B2V_CUDA(v, cudaMemcpyAsync(v->d_color[s], color, pixels * 3, cudaMemcpyHostToDevice, v->copy));
            d_color = v->d_color[s];
        }
        B2V_CUDA(v, cudaEventRecord(v->ev_ready[s], v->copy));
        B2V_CUDA(v, cudaStreamWaitEvent(as, v->ev_ready[s], 0));
    } else if (v->overlap && !v->inputs_fenced) {
        // device inputs were produced by earlier work on the caller's stream (a batch call fences once:
        // an event recorded now would also wait for the previous frame's integrate kernel)
        B2V_CUDA(v, cudaEventRecord(v->ev_in, cs));
        B2V_CUDA(v, cudaStreamWaitEvent(as, v->ev_in, 0));
    }
    if (v->rings_stale) {
        // first single frame after a fused batch: the batch advanced frame_id without passing through the
        // per-frame rings, so the ring this frame counts into may still hold an old frame's counts (only the
        // previous per-frame allocate re-arms the next ring).  Rare transition: order it after everything on the
        // compute stream and clear all ring counters.
        B2V_CUDA(v, cudaEventRecord(v->ev_in, cs));
        B2V_CUDA(v, cudaStreamWaitEvent(as, v->ev_in, 0));
        B2V_CUDA(v, cudaMemsetAsync(v->meta.counters + kCtrActive0, 0, 2 * kActiveRing * sizeof(uint32_t), as));
        v->rings_stale = false;
    } else if (v->overlap && v->frame_id >= 3) {
        // allocate(f) recycles the ring slot / texel buffer last read by integrate(f - 3) .. (f - 4)
        B2V_CUDA(v, cudaStreamWaitEvent(as, v->ev_int_done[(v->frame_id - 3) % kActiveRing], 0));
    }
    if (u16) {  // widen the raw depth into the float staging slot: float(u16) * scale, one rounding
        const uint16_t *src = dev_depth ? reinterpret_cast<const uint16_t *>(depth) : v->d_depth16[s];
        B2V_CUDA(v, launch_depth_u16_to_f32(src, v->d_depth[s], pixels, v->in_u16_scale, as));
        d_depth = v->d_depth[s];
        v->launches += 1;
    }
    {
        const int rc = rectify_frame(v, &d_depth, &d_color, height, width, s, as);
        if (rc != B2V_OK) return rc;
    }
    FrameParams P;
    fill_frame_params(&P, K, Tcw, height, width, v->geo, v->frame_id + 1);
    if (v->lam_H != height || v->lam_W != width || std::memcmp(v->lam_K, K, sizeof(v->lam_K)) != 0) {
        if (v->overlap) {  // the lambda image is read by allocate kernels that may still be in flight
            B2V_CUDA(v, cudaStreamSynchronize(v->alloc));
        }
        B2V_CUDA(v, launch_lambda(P, v->d_lambda, as));
        std::memcpy(v->lam_K, K, sizeof(v->lam_K));
        v->lam_H = height;
        v->lam_W = width;
        v->launches += 1;
    }
    cudaEvent_t *pe = nullptr;
    if (v->prof_enabled) {
        const int rc = grow_profile_events(v, 4);
        if (rc != B2V_OK) return rc;
        pe = &v->prof_events[v->prof_used];
        v->prof_used += 4;
        B2V_CUDA(v, cudaEventRecord(pe[0], as));
    }
    float4 *tex = v->d_texel[s];
    P.I.tex = tex;
    B2V_CUDA(v, launch_allocate(P, d_depth, d_color, v->d_lambda, tex, v->table, v->meta, ring,
                                frame_maps(v, d_depth, d_color, height, width), &v->lam_map, as));
    if (staged || u16) B2V_CUDA(v, cudaEventRecord(v->ev_free[s], as));  // the raw frame is consumed by allocate only
    v->launches += 1;
    v->frame_id += 1;
    if (v->prof_enabled) v->prof_frames += 1;
    v->last_group_buf = -1;
    if (pe) B2V_CUDA(v, cudaEventRecord(pe[1], as));
    if (v->overlap) {
        B2V_CUDA(v, cudaEventRecord(v->ev_alloc_done[ring], as));
        B2V_CUDA(v, cudaStreamWaitEvent(cs, v->ev_alloc_done[ring], 0));
    }
    if (pe) B2V_CUDA(v, cudaEventRecord(pe[2], cs));
    B2V_CUDA(v, launch_integrate(P, volume_consts(v->geo), v->table, v->meta, ring, v->grid_ctas, cs));
    if (pe) B2V_CUDA(v, cudaEventRecord(pe[3], cs));
    if (v->overlap) B2V_CUDA(v, cudaEventRecord(v->ev_int_done[ring], cs));
    v->launches += 1;
    if (v->prof_enabled) v->prof_int_launches += 1;
    return B2V_OK;
}

extern "C" int b2v_integrate(b2v_volume *v, const float *depth, const uint8_t *color, int32_t height,
                             int32_t width, const double K[4], const double Tcw[16], void *stream) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    return integrate_frame(v, depth, color, height, width, K, Tcw, stream);
}

extern "C" int b2v_set_rectification(b2v_volume *v, const float *map_x, const float *map_y, int32_t height,
                                     int32_t width, int32_t swap_rb) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    const int rc = read_counters(v);  // drains every stream
    if (rc == B2V_ERR_CUDA) return rc;
    cudaFree(v->d_mapx);
    cudaFree(v->d_mapy);
    cudaFree(v->d_rdepth[0]);
    cudaFree(v->d_rcolor[0]);
    v->d_mapx = v->d_mapy = nullptr;
    for (int s = 0; s < kStage; ++s) {
        v->d_rdepth[s] = nullptr;
        v->d_rcolor[s] = nullptr;
    }
    v->rect_H = v->rect_W = 0;
    v->rect_swap = swap_rb;
    v->map_cache.clear();
    if (!map_x || !map_y) return B2V_OK;
    if (height <= 0 || width <= 0) {
        v->err = "b2v_set_rectification: bad image size";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    const size_t pixels = static_cast<size_t>(height) * width;
    B2V_CUDA(v, cudaMalloc(&v->d_mapx, pixels * sizeof(float)));
    B2V_CUDA(v, cudaMalloc(&v->d_mapy, pixels * sizeof(float)));
    B2V_CUDA(v, cudaMemcpy(v->d_mapx, map_x, pixels * sizeof(float), cudaMemcpyHostToDevice));
    B2V_CUDA(v, cudaMemcpy(v->d_mapy, map_y, pixels * sizeof(float), cudaMemcpyHostToDevice));
    float *dbase = nullptr;
    uint8_t *cbase = nullptr;
    B2V_CUDA(v, cudaMalloc(&dbase, pixels * sizeof(float) * kStage));
    B2V_CUDA(v, cudaMalloc(&cbase, pixels * 3 * kStage));
    for (int s = 0; s < kStage; ++s) {
        v->d_rdepth[s] = dbase + pixels * s;
        v->d_rcolor[s] = cbase + pixels * 3 * s;
    }
    v->rect_H = height;
    v->rect_W = width;
    return B2V_OK;
}

extern "C" int b2v_remap(const void *src, int32_t kind, int32_t height, int32_t width, const float *map_x,
                         const float *map_y, void *dst, int32_t swap_rb, int32_t device) {
    if (!src || !dst || !map_x || !map_y || height <= 0 || width <= 0 || (kind != 0 && kind != 1))
        return B2V_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return B2V_ERR_CUDA;
    const size_t pixels = static_cast<size_t>(height) * width;
    const size_t bytes = pixels * (kind == 0 ? 3 : 4);
    void *d_src = nullptr, *d_dst = nullptr;
    float *d_mx = nullptr, *d_my = nullptr;
    cudaError_t e = cudaMalloc(&d_src, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_dst, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d_mx, pixels * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_my, pixels * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_src, src, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_mx, map_x, pixels * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_my, map_y, pixels * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = kind == 0 ? launch_remap_u8c3_linear(static_cast<const uint8_t *>(d_src), height, width, d_mx, d_my,
                                                 static_cast<uint8_t *>(d_dst), swap_rb, nullptr)
                      : launch_remap_b32_nearest(d_src, height, width, d_mx, d_my, d_dst, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(dst, d_dst, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_src);
    cudaFree(d_dst);
    cudaFree(d_mx);
    cudaFree(d_my);
    return e == cudaSuccess ? B2V_OK : B2V_ERR_CUDA;
}

// rectify one frame (raw staging or caller device buffers -> rectified slot), on the allocate stream
static int rectify_frame(b2v_volume *v, const float **d_depth, const uint8_t **d_color, int H, int W, int slot,
                         cudaStream_t as) {
    if (!v->d_mapx) return B2V_OK;
    if (H != v->rect_H || W != v->rect_W) {
        v->err = "rectification maps were installed for a different image size";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    B2V_CUDA(v, launch_remap_b32_nearest(*d_depth, H, W, v->d_mapx, v->d_mapy, v->d_rdepth[slot], as));
    B2V_CUDA(v, launch_remap_u8c3_linear(*d_color, H, W, v->d_mapx, v->d_mapy, v->d_rcolor[slot], v->rect_swap, as));
    *d_depth = v->d_rdepth[slot];
    *d_color = v->d_rcolor[slot];
    v->launches += 2;
    return B2V_OK;
}

static int ensure_group_buffers(b2v_volume *v, size_t pixels) {
    if (pixels <= v->gtex_pixels) return B2V_OK;
    B2V_CUDA(v, cudaDeviceSynchronize());
    v->gtex_pixels = 0;  // stays 0 if an allocation below fails
    for (float4 *&t : v->d_gtex) {
        cudaFree(t);
        t = nullptr;
        B2V_CUDA(v, cudaMalloc(&t, pixels * sizeof(float4)));
    }
    v->gtex_pixels = pixels;
    return B2V_OK;
}

extern "C" int b2v_integrate_batch(b2v_volume *v, int32_t n_frames, const float *depth,
                                   const uint8_t *color, int32_t height, int32_t width,
                                   const double K[4], const double *Tcw, void *stream) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    if (n_frames < 0 || (n_frames > 0 && (!depth || !color || !Tcw || !K)) || height <= 0 || width <= 0) {
        v->err = "b2v_integrate_batch: bad arguments";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    if (n_frames == 0) return B2V_OK;
    if (!(K[0] > 0.0) || !(K[1] > 0.0)) {
        v->err = "b2v_integrate_batch: focal lengths must be positive";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    B2V_CUDA(v, cudaSetDevice(v->cfg.device));
    const size_t pixels = static_cast<size_t>(height) * width;
    const bool dd = is_device_pointer(depth), dc = is_device_pointer(color);
    const int dev_hint = dd == dc ? (dd ? 1 : 0) : -1;
    if (stream != nullptr && dev_hint != 1) {
        v->err = "b2v_integrate_batch: a caller stream requires device image pointers";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t cs = stream ? static_cast<cudaStream_t>(stream) : v->compute;
    cudaStream_t as = v->overlap ? v->alloc : cs;
    struct FenceGuard {  // every exit path (errors included) drops the batch-wide input fence
        b2v_volume *v;
        ~FenceGuard() { v->inputs_fenced = false; }
    } fence_guard{v};
    if (v->overlap) {
        // one fence for the whole batch: by default everything enqueued on the caller's stream so far (the inputs'
        // producers, earlier per-frame work) happens before the batch's allocate kernels.  A caller that knows better
        // (b2v_set_input_event: "the inputs are ready when this event fires") keeps the allocate kernels of this batch
        // from also waiting for the update kernels of the previous one.
        if (v->input_event && dev_hint == 1) {
            B2V_CUDA(v, cudaStreamWaitEvent(v->alloc, v->input_event, 0));
        } else {
            B2V_CUDA(v, cudaEventRecord(v->ev_in, cs));
            B2V_CUDA(v, cudaStreamWaitEvent(v->alloc, v->ev_in, 0));
        }
        v->inputs_fenced = true;
    } else if (v->input_event && dev_hint == 1) {
        B2V_CUDA(v, cudaStreamWaitEvent(cs, v->input_event, 0));
    }
    v->input_event = nullptr;
    int rc = B2V_OK;
    const bool u16 = v->in_u16_scale > 0.0f;  // `depth` is really const uint16_t *
    const uint16_t *depth16 = reinterpret_cast<const uint16_t *>(depth);
    if (!v->fuse || n_frames < 2) {
        for (int32_t f = 0; f < n_frames && rc == B2V_OK; ++f)
            rc = integrate_frame(v, u16 ? reinterpret_cast<const float *>(depth16 + pixels * f) : depth + pixels * f,
                                 color + pixels * 3 * f, height, width, K, Tcw + 16 * static_cast<size_t>(f), stream,
                                 dev_hint);
        return rc;
    }
    rc = ensure_group_buffers(v, pixels);
    if (rc == B2V_OK) rc = ensure_staging(v, pixels);
    if (rc == B2V_OK && u16) rc = ensure_staging16(v, pixels);
    if (rc != B2V_OK) return rc;
    v->rings_stale = true;
    v->last_stream = stream ? cs : nullptr;
    const bool staged = dev_hint != 1;
    const int gsz = std::max(1, std::min(v->group_frames, kMaxGroup));
    for (int32_t g0 = 0; g0 < n_frames; g0 += gsz) {
        const int count = std::min<int32_t>(gsz, n_frames - g0);
        const int buf = static_cast<int>(v->group_id % kGroupBufs);
        // the group buffer (masks, union list, texel images) was last used by group id - kGroupBufs
        B2V_CUDA(v, cudaStreamWaitEvent(as, v->ev_group_done[buf], 0));
        B2V_CUDA(v, cudaMemsetAsync(v->meta.counters + group_ctr(buf, 0), 0, kGroupCtrStride * sizeof(uint32_t), as));
        static thread_local GroupAllocArgs aargs;  // ~15 KB: keep it off the stack
        static thread_local GroupArgs args;
        std::memset(&args, 0, sizeof(args));
        args.V = volume_consts(v->geo);
        args.count = count;
        aargs.count = count;
        aargs.use_tma = 1;
        if (staged) {
            // the raw staging slots of this buffer were consumed by the allocate launch of group id - kGroupBufs;
            // the group's frames are contiguous on both sides: one H2D copy per image type
            B2V_CUDA(v, cudaStreamWaitEvent(v->copy, v->ev_galloc[buf], 0));
            const int s0 = buf * kMaxGroup;
            if (u16)
                B2V_CUDA(v, cudaMemcpyAsync(v->d_depth16[s0], depth16 + pixels * g0, pixels * sizeof(uint16_t) * count,
                                            cudaMemcpyHostToDevice, v->copy));
            else
                B2V_CUDA(v, cudaMemcpyAsync(v->d_depth[s0], depth + pixels * g0, pixels * sizeof(float) * count,
                                            cudaMemcpyHostToDevice, v->copy));
            B2V_CUDA(v, cudaMemcpyAsync(v->d_color[s0], color + pixels * 3 * g0, pixels * 3 * count,
                                        cudaMemcpyHostToDevice, v->copy));
        }
        for (int k = 0; k < count; ++k) {
            const size_t f = static_cast<size_t>(g0 + k);
            const float *d_depth = u16 ? nullptr : depth + pixels * f;
            const uint8_t *d_color = color + pixels * 3 * f;
            if (staged || u16) d_depth = v->d_depth[buf * kMaxGroup + k];  // (widened) float staging slot
            if (staged) d_color = v->d_color[buf * kMaxGroup + k];
            FrameParams P;
            fill_frame_params(&P, K, Tcw + 16 * f, height, width, v->geo, v->frame_id + 1);
            P.group_bit = k;
            P.group_buf = buf;
            aargs.pose[k] = P.pose;
            if (k == 0) {
                aargs.P = P;
                aargs.frame_id0 = v->frame_id + 1;
            }
            if (k == 0 && (v->lam_H != height || v->lam_W != width || std::memcmp(v->lam_K, K, sizeof(v->lam_K)) != 0)) {
                if (v->overlap) B2V_CUDA(v, cudaStreamSynchronize(v->alloc));
                B2V_CUDA(v, launch_lambda(P, v->d_lambda, as));
                std::memcpy(v->lam_K, K, sizeof(v->lam_K));
                v->lam_H = height;
                v->lam_W = width;
                v->launches += 1;
            }
            float4 *tex = v->d_gtex[buf * kMaxGroup + k];
            aargs.depth[k] = d_depth;
            aargs.color[k] = d_color;
            aargs.tex[k] = tex;
            P.I.tex = tex;
            args.f[k] = P.I;
            v->frame_id += 1;
        }
        if (staged) {
            B2V_CUDA(v, cudaEventRecord(v->ev_ready[buf], v->copy));  // all frames of the group uploaded
            B2V_CUDA(v, cudaStreamWaitEvent(as, v->ev_ready[buf], 0));
        }
        if (u16) {  // widen the group's raw depth into its (contiguous) float staging slots in one launch
            const uint16_t *src = staged ? v->d_depth16[buf * kMaxGroup] : depth16 + pixels * g0;
            B2V_CUDA(v, launch_depth_u16_to_f32(src, v->d_depth[buf * kMaxGroup], pixels * count, v->in_u16_scale, as));
            v->launches += 1;
        }
        for (int k = 0; k < count; ++k) {  // optional rectification, then the TMA descriptors of the final images
            const int rrc = rectify_frame(v, &aargs.depth[k], &aargs.color[k], height, width, buf * kMaxGroup + k, as);
            if (rrc != B2V_OK) return rrc;
            const FrameMaps *fm = frame_maps(v, aargs.depth[k], aargs.color[k], height, width);
            if (fm) aargs.maps[k] = *fm; else aargs.use_tma = 0;
        }
        cudaEvent_t *pe = nullptr;
        if (v->prof_enabled) {
            const int prc = grow_profile_events(v, 4);
            if (prc != B2V_OK) return prc;
            pe = &v->prof_events[v->prof_used];
            v->prof_used += 4;
            v->prof_frames += count;
            v->prof_int_launches += 1;
            B2V_CUDA(v, cudaEventRecord(pe[0], as));
        }
        aargs.lmap = v->lam_map;
        B2V_CUDA(v, launch_allocate_group(aargs, v->d_lambda, v->table, v->meta, as));
        if (pe) B2V_CUDA(v, cudaEventRecord(pe[1], as));
        B2V_CUDA(v, cudaEventRecord(v->ev_galloc[buf], as));
        if (v->overlap) B2V_CUDA(v, cudaStreamWaitEvent(cs, v->ev_galloc[buf], 0));
        if (pe) B2V_CUDA(v, cudaEventRecord(pe[2], cs));
        B2V_CUDA(v, launch_integrate_group(args, v->table, v->meta, buf, v->grid_ctas, cs));
        if (pe) B2V_CUDA(v, cudaEventRecord(pe[3], cs));
        B2V_CUDA(v, cudaEventRecord(v->ev_group_done[buf], cs));
        v->launches += 3;
        v->last_group_buf = buf;
        v->last_group_count = count;
        v->group_id += 1;
    }
    return B2V_OK;
}

// Raw 16-bit depth (e.g. TUM / ScanNet PNGs): uploaded as uint16 (2 instead of 4 bytes per pixel over PCIe) and
// widened on the device to float(u16) * depth_scale in float32 - the value numpy's
// `depth.astype(np.float32) * depth_factor` produces (volumetric_integrator_base.py:1008-1015).
extern "C" int b2v_integrate_batch_u16(b2v_volume *v, int32_t n_frames, const uint16_t *depth, float depth_scale,
                                       const uint8_t *color, int32_t height, int32_t width, const double K[4],
                                       const double *Tcw, void *stream) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    if (!(depth_scale > 0.0f)) {
        v->err = "b2v_integrate_batch_u16: depth_scale must be positive";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    v->in_u16_scale = depth_scale;
    const int rc = b2v_integrate_batch(v, n_frames, reinterpret_cast<const float *>(depth), color, height, width, K,
                                       Tcw, stream);
    v->in_u16_scale = 0.0f;
    return rc;
}

extern "C" int b2v_integrate_u16(b2v_volume *v, const uint16_t *depth, float depth_scale, const uint8_t *color,
                                 int32_t height, int32_t width, const double K[4], const double Tcw[16],
                                 void *stream) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    if (!(depth_scale > 0.0f)) {
        v->err = "b2v_integrate_u16: depth_scale must be positive";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    v->in_u16_scale = depth_scale;
    const int rc = integrate_frame(v, reinterpret_cast<const float *>(depth), color, height, width, K, Tcw, stream);
    v->in_u16_scale = 0.0f;
    return rc;
}

extern "C" int b2v_set_input_event(b2v_volume *v, void *event) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    v->input_event = static_cast<cudaEvent_t>(event);
    return B2V_OK;
}

extern "C" int b2v_set_group_size(b2v_volume *v, int32_t frames) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    if (frames < 1 || frames > kMaxGroup) {
        v->err = "b2v_set_group_size: 1..32 frames";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    const int rc = read_counters(v);
    if (rc == B2V_ERR_CUDA) return rc;
    v->group_frames = frames;
    return B2V_OK;
}

extern "C" int b2v_set_fusion(b2v_volume *v, int32_t enable) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    const int rc = read_counters(v);
    if (rc == B2V_ERR_CUDA) return rc;
    v->fuse = enable != 0;
    return B2V_OK;
}

extern "C" int b2v_synchronize(b2v_volume *v) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    return read_counters(v);
}

extern "C" int64_t b2v_num_blocks(b2v_volume *v) {
    if (!v) return -1;
    const int rc = read_counters(v);
    if (rc == B2V_ERR_CUDA) return -1;
    return block_count(v);
}

extern "C" int b2v_last_frame_stats(b2v_volume *v, int64_t *touched_blocks, int64_t *new_blocks) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    const int rc = read_counters(v);
    if (rc == B2V_ERR_CUDA) return rc;
    const int ring = v->frame_id ? static_cast<int>((v->frame_id - 1) % kActiveRing) : 0;
    if (touched_blocks) {
        if (v->frame_id == 0)
            *touched_blocks = 0;
        else if (v->last_group_buf >= 0)
            *touched_blocks = v->h_counters[group_ctr(v->last_group_buf, kGcTouched0) + v->last_group_count - 1];
        else
            *touched_blocks = v->h_counters[kCtrActive0 + ring];
    }
    if (new_blocks) {
        if (v->frame_id == 0)
            *new_blocks = 0;
        else if (v->last_group_buf >= 0)  // after a fused batch: blocks allocated by the last group
            *new_blocks = v->h_counters[group_ctr(v->last_group_buf, kGcNew)];
        else
            *new_blocks = v->h_counters[kCtrNew0 + ring];
    }
    return rc;
}

extern "C" int b2v_set_overlap(b2v_volume *v, int32_t enable) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    const int rc = read_counters(v);  // drains every stream first
    if (rc == B2V_ERR_CUDA) return rc;
    v->overlap = enable != 0;
    return B2V_OK;
}

extern "C" int b2v_profile_enable(b2v_volume *v, int32_t enable) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    v->prof_enabled = enable != 0;
    v->prof_used = 0;
    v->prof_frames = 0;
    v->prof_int_launches = 0;
    return B2V_OK;
}

extern "C" int b2v_profile_read(b2v_volume *v, double *allocate_ms, double *integrate_ms, int64_t *frames,
                                int64_t *integrate_launches) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    B2V_CUDA(v, cudaSetDevice(v->cfg.device));
    double a = 0.0, b = 0.0;
    if (std::getenv("B2V_DEBUG_TIMELINE") && v->prof_used >= 4) {  // debug: event times relative to the first
        for (size_t k = 0; k + 3 < v->prof_used && k < 4 * 12; k += 4) {
            float t[4];
            for (int j = 0; j < 4; ++j) {
                cudaEventSynchronize(v->prof_events[k + j]);
                cudaEventElapsedTime(&t[j], v->prof_events[0], v->prof_events[k + j]);
            }
            std::fprintf(stderr, "[b2v timeline] launch %zu: alloc %.1f..%.1f us  integrate %.1f..%.1f us\n", k / 4,
                         1e3 * t[0], 1e3 * t[1], 1e3 * t[2], 1e3 * t[3]);
        }
    }
    for (size_t k = 0; k + 3 < v->prof_used; k += 4) {
        B2V_CUDA(v, cudaEventSynchronize(v->prof_events[k + 1]));
        B2V_CUDA(v, cudaEventSynchronize(v->prof_events[k + 3]));
        float ms = 0.0f;
        B2V_CUDA(v, cudaEventElapsedTime(&ms, v->prof_events[k], v->prof_events[k + 1]));
        a += ms;
        B2V_CUDA(v, cudaEventElapsedTime(&ms, v->prof_events[k + 2], v->prof_events[k + 3]));
        b += ms;
    }
    if (allocate_ms) *allocate_ms = a;
    if (integrate_ms) *integrate_ms = b;
    if (frames) *frames = v->prof_frames;
    if (integrate_launches) *integrate_launches = v->prof_int_launches;
    v->prof_used = 0;
    v->prof_frames = 0;
    v->prof_int_launches = 0;
    return B2V_OK;
}

extern "C" int b2v_counters(b2v_volume *v, int64_t *block_updates, int64_t *kernel_launches,
                            int64_t *block_visits) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    const int rc = read_counters(v);
    if (rc == B2V_ERR_CUDA) return rc;
    if (block_updates)
        *block_updates = static_cast<int64_t>(static_cast<uint64_t>(v->h_counters[kCtrUpdatesLo]) |
                                              (static_cast<uint64_t>(v->h_counters[kCtrUpdatesHi]) << 32));
    if (kernel_launches) *kernel_launches = v->launches;
    if (block_visits)
        *block_visits = static_cast<int64_t>(static_cast<uint64_t>(v->h_counters[kCtrVisitsLo]) |
                                             (static_cast<uint64_t>(v->h_counters[kCtrVisitsHi]) << 32));
    return rc;
}

extern "C" int64_t b2v_dump_blocks(b2v_volume *v, int32_t *keys, uint64_t *hashes, float *voxels) {
    if (!v) return -1;
    if (read_counters(v) == B2V_ERR_CUDA) return -1;
    const uint32_t nb = block_count(v);
    if (nb == 0) return 0;
    if (keys) {
        std::vector<int4> tmp(nb);
        if (cudaMemcpy(tmp.data(), v->meta.block_keys, nb * sizeof(int4), cudaMemcpyDeviceToHost) != cudaSuccess)
            return -1;
        for (uint32_t i = 0; i < nb; ++i) {
            keys[3 * i + 0] = tmp[i].x;
            keys[3 * i + 1] = tmp[i].y;
            keys[3 * i + 2] = tmp[i].z;
        }
    }
    if (hashes) {
        uint64_t *d_h = nullptr;
        if (cudaMalloc(&d_h, nb * sizeof(uint64_t)) != cudaSuccess) return -1;
        cudaError_t e = launch_block_hashes(v->meta.block_keys, d_h, nb, v->compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(v->compute);
        if (e == cudaSuccess) e = cudaMemcpy(hashes, d_h, nb * sizeof(uint64_t), cudaMemcpyDeviceToHost);
        cudaFree(d_h);
        v->launches += 1;
        if (e != cudaSuccess) return -1;
    }
    if (voxels) {
        if (cudaMemcpy(voxels, v->meta.pool, static_cast<size_t>(nb) * kBlockFloats * sizeof(float),
                       cudaMemcpyDeviceToHost) != cudaSuccess)
            return -1;
    }
    return nb;
}

extern "C" int b2v_upload_blocks(b2v_volume *v, int64_t n_blocks, const int32_t *keys, const float *voxels) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    if (n_blocks < 0 || (n_blocks > 0 && (!keys || !voxels))) {
        v->err = "b2v_upload_blocks: bad arguments";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    if (n_blocks == 0) return B2V_OK;
    B2V_CUDA(v, cudaSetDevice(v->cfg.device));
    const size_t n = static_cast<size_t>(n_blocks);
    std::vector<int4> k4(n);
    for (size_t i = 0; i < n; ++i) k4[i] = make_int4(keys[3 * i], keys[3 * i + 1], keys[3 * i + 2], 0);
    int4 *d_k = nullptr;
    float *d_v = nullptr;
    uint32_t *d_i = nullptr;
    cudaError_t e = cudaMalloc(&d_k, n * sizeof(int4));
    if (e == cudaSuccess) e = cudaMalloc(&d_v, n * kBlockFloats * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_i, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_k, k4.data(), n * sizeof(int4), cudaMemcpyHostToDevice, v->compute);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_v, voxels, n * kBlockFloats * sizeof(float), cudaMemcpyHostToDevice, v->compute);
    if (e == cudaSuccess)
        e = launch_upload_blocks(d_k, d_v, static_cast<uint32_t>(n), d_i, v->table, v->meta, v->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->compute);
    cudaFree(d_k);
    cudaFree(d_v);
    cudaFree(d_i);
    v->launches += 2;
    if (e != cudaSuccess) {
        v->err = std::string("b2v_upload_blocks: ") + cudaGetErrorString(e);
        return B2V_ERR_CUDA;
    }
    return read_counters(v);
}

// Device-to-device block exchange (multi-GPU mesh gather): keys as int32 x 4 {x, y, z, 0}
extern "C" int64_t b2v_export_blocks_device(b2v_volume *v, int32_t *d_keys4, float *d_voxels, int64_t max_blocks) {
    if (!v) return -1;
    if (read_counters(v) == B2V_ERR_CUDA) return -1;
    const uint32_t nb = block_count(v);
    if (!d_keys4 && !d_voxels) return nb;
    if (static_cast<int64_t>(nb) > max_blocks) {
        v->err = "b2v_export_blocks_device: destination too small";
        return -1;
    }
    if (nb == 0) return 0;
    cudaError_t e = cudaSuccess;
    if (d_keys4) e = cudaMemcpyAsync(d_keys4, v->meta.block_keys, nb * sizeof(int4), cudaMemcpyDeviceToDevice, v->compute);
    if (e == cudaSuccess && d_voxels)
        e = cudaMemcpyAsync(d_voxels, v->meta.pool, static_cast<size_t>(nb) * kBlockFloats * sizeof(float),
                            cudaMemcpyDeviceToDevice, v->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->compute);
    if (e != cudaSuccess) {
        v->err = std::string("b2v_export_blocks_device: ") + cudaGetErrorString(e);
        return -1;
    }
    return nb;
}

extern "C" int b2v_import_blocks_device(b2v_volume *v, int64_t n_blocks, const int32_t *d_keys4, const float *d_voxels) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    if (n_blocks < 0 || (n_blocks > 0 && (!d_keys4 || !d_voxels))) {
        v->err = "b2v_import_blocks_device: bad arguments";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    if (n_blocks == 0) return B2V_OK;
    B2V_CUDA(v, cudaSetDevice(v->cfg.device));
    uint32_t *d_i = nullptr;
    cudaError_t e = cudaMalloc(&d_i, static_cast<size_t>(n_blocks) * sizeof(uint32_t));
    if (e == cudaSuccess)
        e = launch_upload_blocks(reinterpret_cast<const int4 *>(d_keys4), d_voxels, static_cast<uint32_t>(n_blocks), d_i,
                                 v->table, v->meta, v->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->compute);
    cudaFree(d_i);
    v->launches += 2;
    if (e != cudaSuccess) {
        v->err = std::string("b2v_import_blocks_device: ") + cudaGetErrorString(e);
        return B2V_ERR_CUDA;
    }
    return read_counters(v);
}

extern "C" int64_t b2v_last_touched_keys(b2v_volume *v, int32_t *keys, int64_t max_keys) {
    if (!v) return -1;
    if (read_counters(v) == B2V_ERR_CUDA) return -1;
    if (v->frame_id == 0) return 0;
    const int ring = static_cast<int>((v->frame_id - 1) % kActiveRing);
    const bool grp = v->last_group_buf >= 0;  // after a fused batch: the last group's union of touched blocks
    uint32_t n = grp ? v->h_counters[group_ctr(v->last_group_buf, kGcUnion)] : v->h_counters[kCtrActive0 + ring];
    if (n > v->meta.capacity) n = v->meta.capacity;
    if (!keys) return n;
    if (static_cast<int64_t>(n) > max_keys) n = static_cast<uint32_t>(max_keys);
    if (n == 0) return 0;
    int4 *d_k = nullptr;
    if (cudaMalloc(&d_k, n * sizeof(int4)) != cudaSuccess) return -1;
    std::vector<int4> tmp(n);
    const uint32_t *list = grp ? v->meta.union_slots + static_cast<size_t>(v->last_group_buf) * v->meta.capacity
                               : v->meta.active_slots + static_cast<size_t>(ring) * v->meta.capacity;
    cudaError_t e = launch_gather_active_keys(v->table, list, n, d_k, v->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->compute);
    if (e == cudaSuccess) e = cudaMemcpy(tmp.data(), d_k, n * sizeof(int4), cudaMemcpyDeviceToHost);
    cudaFree(d_k);
    v->launches += 1;
    if (e != cudaSuccess) return -1;
    for (uint32_t i = 0; i < n; ++i) {
        keys[3 * i + 0] = tmp[i].x;
        keys[3 * i + 1] = tmp[i].y;
        keys[3 * i + 2] = tmp[i].z;
    }
    return n;
}

// ---- mesh / point cloud ---------------------------------------------------------------------

template <typename T> static cudaError_t regrow(T **p, size_t n) {
    cudaFree(*p);
    *p = nullptr;
    return cudaMalloc(p, (n ? n : 1) * sizeof(T));
}

static int ensure_mesh_scratch(b2v_volume *v, uint32_t nb) {
    if (!v->mb.totals) B2V_CUDA(v, cudaMalloc(&v->mb.totals, kNumMeshTotals * sizeof(uint32_t)));
    if (nb <= v->mesh_blocks_cap) return B2V_OK;
    const size_t n = nb;
    B2V_CUDA(v, regrow(&v->mb.nbr, n * 8));
    B2V_CUDA(v, regrow(&v->mb.cube, n * kVox));
    B2V_CUDA(v, regrow(&v->mb.edge_mask, n * (kVox / 4)));
    B2V_CUDA(v, regrow(&v->mb.local, n * kVox));
    B2V_CUDA(v, regrow(&v->mb.sums, n * 2));
    B2V_CUDA(v, regrow(&v->mb.offs, n * 2));
    B2V_CUDA(v, regrow(&v->mb.partials, 2 * ((n + 1023) / 1024)));
    B2V_CUDA(v, regrow(&v->mb.work, n * 4));
    v->mesh_blocks_cap = nb;
    return B2V_OK;
}

static int extract_common(b2v_volume *v, bool mesh, int64_t *n_vertices, int64_t *n_triangles) {
    int rc = read_counters(v);
    if (rc == B2V_ERR_CUDA) return rc;
    const uint32_t nb = block_count(v);
    rc = ensure_mesh_scratch(v, nb);
    if (rc != B2V_OK) return rc;
    v->mb.n_blocks = nb;
    cudaStream_t cs = v->compute;
    const int sms = v->sm_count > 0 ? v->sm_count : 148;
    if (mesh) {
        B2V_CUDA(v, launch_mesh_classify(v->table, v->meta, v->mb, sms, cs));
    } else {
        B2V_CUDA(v, launch_point_masks(v->table, v->meta, v->mb, sms, cs));
    }
    B2V_CUDA(v, launch_mesh_scan(v->mb, sms, cs));
    B2V_CUDA(v, cudaMemcpyAsync(v->h_totals, v->mb.totals, kNumMeshTotals * sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
    B2V_CUDA(v, cudaStreamSynchronize(cs));
    const size_t nv = v->h_totals[kMtVertices], nt = v->h_totals[kMtTriangles];
    if (nv > v->mesh_v_cap) {
        B2V_CUDA(v, regrow(&v->mb.vertices, nv * 3));
        B2V_CUDA(v, regrow(&v->mb.colors, nv * 3));
        B2V_CUDA(v, regrow(&v->mb.edge_ids, nv * 4));
        v->mesh_v_cap = nv;
    }
    if (nt > v->mesh_t_cap) {
        B2V_CUDA(v, regrow(&v->mb.triangles, nt * 3));
        v->mesh_t_cap = nt;
    }
    B2V_CUDA(v, launch_mesh_vertices(v->meta, v->mb, v->geo.voxel_length, v->geo.unit_shift, !mesh, v->h_totals[kMtVertexBlocks], cs));
    if (mesh) B2V_CUDA(v, launch_mesh_triangles(v->mb, v->h_totals[kMtTriangleBlocks], cs));
    B2V_CUDA(v, cudaStreamSynchronize(cs));
    v->launches += mesh ? 6 : 5;
    v->last_nv = static_cast<int64_t>(nv);
    v->last_nt = mesh ? static_cast<int64_t>(nt) : 0;
    if (n_vertices) *n_vertices = v->last_nv;
    if (n_triangles) *n_triangles = v->last_nt;
    return rc;
}

extern "C" int b2v_last_mesh_stats(b2v_volume *v, int64_t stats[5]) {
    if (!v || !stats) return B2V_ERR_INVALID_ARGUMENT;
    stats[0] = v->mb.n_blocks;
    stats[1] = v->h_totals[kMtCandidates];
    stats[2] = v->h_totals[kMtTiles];
    stats[3] = v->h_totals[kMtVertexBlocks];
    stats[4] = v->h_totals[kMtTriangleBlocks];
    return B2V_OK;
}

extern "C" int b2v_extract_mesh(b2v_volume *v, int64_t *n_vertices, int64_t *n_triangles) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    return extract_common(v, true, n_vertices, n_triangles);
}

extern "C" int b2v_copy_mesh(b2v_volume *v, double *vertices, double *colors, int32_t *edge_ids,
                             int32_t *triangles) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    const size_t nv = static_cast<size_t>(v->last_nv), nt = static_cast<size_t>(v->last_nt);
    if (vertices && nv) B2V_CUDA(v, cudaMemcpy(vertices, v->mb.vertices, nv * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (colors && nv) B2V_CUDA(v, cudaMemcpy(colors, v->mb.colors, nv * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    if (edge_ids && nv) B2V_CUDA(v, cudaMemcpy(edge_ids, v->mb.edge_ids, nv * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (triangles && nt) B2V_CUDA(v, cudaMemcpy(triangles, v->mb.triangles, nt * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return B2V_OK;
}

extern "C" int b2v_extract_points(b2v_volume *v, int64_t *n_points) {
    if (!v) return B2V_ERR_INVALID_ARGUMENT;
    return extract_common(v, false, n_points, nullptr);
}

extern "C" int b2v_copy_points(b2v_volume *v, double *points, double *colors) {
    return b2v_copy_mesh(v, points, colors, nullptr, nullptr);
}

// ---- point-average grid (duck type B) ---------------------------------------------------------

struct b2v_grid {
    float voxel_size = 0.0f, inv_voxel_size = 0.0f;
    int device = 0;
    cudaStream_t stream = nullptr;
    HashTable table{};
    GridMeta meta{};
    uint32_t *h_counters = nullptr;
    float *d_pts = nullptr, *d_cols = nullptr;
    size_t stage_points = 0;
    uint32_t *d_sums = nullptr, *d_offs = nullptr, *d_total = nullptr;
    uint32_t scan_cap = 0;
    float *d_out_pts = nullptr, *d_out_cols = nullptr;
    size_t out_cap = 0;
    int64_t last_n = 0;
    std::string err;
};

constexpr int kGridBlockWordsHost = 7 * kVox;

extern "C" const char *b2v_grid_last_error(const b2v_grid *g) { return g ? g->err.c_str() : "null grid"; }

static int grid_clear_device(b2v_grid *g, uint32_t used_blocks) {
    const size_t tcap = static_cast<size_t>(g->table.mask) + 1;
    B2V_CUDA(g, cudaMemsetAsync(g->table.entries, 0xFF, tcap * sizeof(uint4), g->stream));
    B2V_CUDA(g, cudaMemsetAsync(g->meta.counters, 0, kNumCounters * sizeof(uint32_t), g->stream));
    B2V_CUDA(g, cudaMemsetAsync(g->meta.pool, 0,
                                static_cast<size_t>(used_blocks) * kGridBlockWordsHost * sizeof(uint32_t),
                                g->stream));
    return B2V_OK;
}

extern "C" int b2v_grid_create(float voxel_size, int32_t block_size, uint32_t capacity_blocks,
                               int32_t device, b2v_grid **out) {
    if (!out) return B2V_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (block_size != B2V_BLOCK_SIZE || !(voxel_size > 0.0f) || capacity_blocks == 0)
        return B2V_ERR_INVALID_ARGUMENT;
    b2v_grid *g = new (std::nothrow) b2v_grid();
    if (!g) return B2V_ERR_INVALID_ARGUMENT;
    g->voxel_size = voxel_size;
    g->inv_voxel_size = 1.0f / voxel_size;  // voxel_block_grid.hpp:6
    g->device = device;
    *out = g;
    B2V_CUDA(g, cudaSetDevice(device));
    B2V_CUDA(g, cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    const uint32_t tcap = next_pow2(static_cast<uint64_t>(capacity_blocks) * 2);
    g->table.mask = tcap - 1;
    g->table.stamp = nullptr;
    g->meta.capacity = capacity_blocks;
    B2V_CUDA(g, cudaMalloc(&g->table.entries, static_cast<size_t>(tcap) * sizeof(uint4)));
    B2V_CUDA(g, cudaMalloc(&g->meta.pool, static_cast<size_t>(capacity_blocks) * kGridBlockWordsHost * sizeof(uint32_t)));
    B2V_CUDA(g, cudaMalloc(&g->meta.block_keys, static_cast<size_t>(capacity_blocks) * sizeof(int4)));
    B2V_CUDA(g, cudaMalloc(&g->meta.counters, kNumCounters * sizeof(uint32_t)));
    B2V_CUDA(g, cudaMalloc(&g->d_total, sizeof(uint32_t)));
    B2V_CUDA(g, cudaMallocHost(&g->h_counters, kNumCounters * sizeof(uint32_t)));
    int rc = grid_clear_device(g, capacity_blocks);
    if (rc != B2V_OK) return rc;
    B2V_CUDA(g, cudaStreamSynchronize(g->stream));
    return B2V_OK;
}

extern "C" int b2v_grid_destroy(b2v_grid *g) {
    if (!g) return B2V_OK;
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    cudaFree(g->table.entries);
    cudaFree(g->meta.pool);
    cudaFree(g->meta.block_keys);
    cudaFree(g->meta.counters);
    cudaFree(g->d_pts);
    cudaFree(g->d_cols);
    cudaFree(g->d_sums);
    cudaFree(g->d_offs);
    cudaFree(g->d_total);
    cudaFree(g->d_out_pts);
    cudaFree(g->d_out_cols);
    cudaFreeHost(g->h_counters);
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
    return B2V_OK;
}

static int grid_read_counters(b2v_grid *g) {
    B2V_CUDA(g, cudaSetDevice(g->device));
    B2V_CUDA(g, cudaMemcpyAsync(g->h_counters, g->meta.counters, kNumCounters * sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, g->stream));
    B2V_CUDA(g, cudaStreamSynchronize(g->stream));
    if (g->h_counters[kCtrError]) {
        g->err = (g->h_counters[kCtrError] & 2u) ? "hash table full: raise capacity_blocks"
                                                 : "block pool full: raise capacity_blocks";
        return B2V_ERR_CAPACITY;
    }
    return B2V_OK;
}

static uint32_t grid_block_count(const b2v_grid *g) {
    const uint32_t n = g->h_counters[kCtrPool];
    return n < g->meta.capacity ? n : g->meta.capacity;
}

extern "C" int b2v_grid_clear(b2v_grid *g) {
    if (!g) return B2V_ERR_INVALID_ARGUMENT;
    int rc = grid_read_counters(g);
    if (rc == B2V_ERR_CUDA) return rc;
    rc = grid_clear_device(g, grid_block_count(g));
    if (rc != B2V_OK) return rc;
    g->err.clear();
    B2V_CUDA(g, cudaStreamSynchronize(g->stream));
    return B2V_OK;
}

static int grid_integrate_any(b2v_grid *g, const void *points, bool f64, const void *colors, bool u8, int64_t n_points) {
    if (!g) return B2V_ERR_INVALID_ARGUMENT;
    if (n_points < 0 || (n_points > 0 && !points)) {
        g->err = "b2v_grid_integrate: bad arguments";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    if (n_points == 0) return B2V_OK;  // voxel_block_grid.hpp:22-24,121-123
    B2V_CUDA(g, cudaSetDevice(g->device));
    const void *d_p = points;
    const void *d_c = colors;
    const bool dev_p = is_device_pointer(points);
    const bool dev_c = colors ? is_device_pointer(colors) : true;
    if (!dev_p || !dev_c) {
        if (static_cast<size_t>(n_points) > g->stage_points) {
            B2V_CUDA(g, cudaStreamSynchronize(g->stream));
            B2V_CUDA(g, regrow(&g->d_pts, static_cast<size_t>(n_points) * 3 * 2));  // room for float64 points
            B2V_CUDA(g, regrow(&g->d_cols, static_cast<size_t>(n_points) * 3));
            g->stage_points = static_cast<size_t>(n_points);
        }
        if (!dev_p) {
            B2V_CUDA(g, cudaMemcpyAsync(g->d_pts, points, static_cast<size_t>(n_points) * 3 * (f64 ? sizeof(double) : sizeof(float)),
                                        cudaMemcpyHostToDevice, g->stream));
            d_p = g->d_pts;
        }
        if (colors && !dev_c) {
            B2V_CUDA(g, cudaMemcpyAsync(g->d_cols, colors, static_cast<size_t>(n_points) * 3 * (u8 ? 1 : sizeof(float)),
                                        cudaMemcpyHostToDevice, g->stream));
            d_c = g->d_cols;
        }
    }
    B2V_CUDA(g, launch_grid_integrate(d_p, f64, d_c, u8, n_points, g->inv_voxel_size, g->table, g->meta, g->stream));
    return B2V_OK;
}

extern "C" int b2v_grid_integrate(b2v_grid *g, const float *points, const float *colors, int64_t n_points) {
    return grid_integrate_any(g, points, false, colors, false, n_points);
}

extern "C" int b2v_grid_integrate_f64(b2v_grid *g, const double *points, const float *colors, int64_t n_points) {
    return grid_integrate_any(g, points, true, colors, false, n_points);
}

extern "C" int b2v_grid_integrate_ex(b2v_grid *g, const void *points, int32_t points_f64, const void *colors,
                                     int32_t colors_u8, int64_t n_points) {
    return grid_integrate_any(g, points, points_f64 != 0, colors, colors_u8 != 0, n_points);
}

extern "C" int b2v_filter_shadow_points(const float *depth, int32_t height, int32_t width, int32_t delta_x,
                                        int32_t delta_y, float fill_value, float *out, int32_t device) {
    if (!depth || !out || height <= 0 || width <= 0 || delta_x < 0 || delta_y < 0 || delta_x >= width ||
        delta_y >= height)
        return B2V_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return B2V_ERR_CUDA;
    const size_t pixels = static_cast<size_t>(height) * width;
    const bool din = is_device_pointer(depth), dout = is_device_pointer(out);
    float *d_in = nullptr, *d_out = nullptr;
    void *scratch = nullptr;
    cudaError_t e = cudaMalloc(&scratch, kShadowScratchBytes);
    if (e == cudaSuccess && !din) {
        e = cudaMalloc(&d_in, pixels * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(d_in, depth, pixels * sizeof(float), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && !dout) e = cudaMalloc(&d_out, pixels * sizeof(float));
    if (e == cudaSuccess)
        e = launch_filter_shadow_points(din ? depth : d_in, height, width, delta_x, delta_y, fill_value,
                                        dout ? out : d_out, scratch, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && !dout) e = cudaMemcpy(out, d_out, pixels * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(scratch);
    cudaFree(d_in);
    cudaFree(d_out);
    return e == cudaSuccess ? B2V_OK : B2V_ERR_CUDA;
}

extern "C" int b2v_grid_integrate_rgbd(b2v_grid *g, const float *depth, const uint8_t *color, int32_t height,
                                       int32_t width, const double K[4], const double Twc[16], float max_depth,
                                       float min_depth, int32_t filter_shadow_points) {
    if (!g) return B2V_ERR_INVALID_ARGUMENT;
    if (!depth || !color || !K || !Twc || height <= 0 || width <= 0) {
        g->err = "b2v_grid_integrate_rgbd: bad arguments";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    B2V_CUDA(g, cudaSetDevice(g->device));
    const size_t pixels = static_cast<size_t>(height) * width;
    const float *d_depth = depth;
    const uint8_t *d_color = color;
    float *tmp_d = nullptr;
    uint8_t *tmp_c = nullptr;
    cudaError_t e = cudaSuccess;
    if (!is_device_pointer(depth)) {
        e = cudaMalloc(&tmp_d, pixels * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpyAsync(tmp_d, depth, pixels * sizeof(float), cudaMemcpyHostToDevice, g->stream);
        d_depth = tmp_d;
    }
    if (e == cudaSuccess && !is_device_pointer(color)) {
        e = cudaMalloc(&tmp_c, pixels * 3);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tmp_c, color, pixels * 3, cudaMemcpyHostToDevice, g->stream);
        d_color = tmp_c;
    }
    float *filtered = nullptr;
    void *scratch = nullptr;
    if (e == cudaSuccess && filter_shadow_points) {  // voxel_grid.py:238-245: depth2pointcloud sees the filtered depth
        e = cudaMalloc(&filtered, pixels * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&scratch, kShadowScratchBytes);
        if (e == cudaSuccess && (height <= 2 || width <= 2)) e = cudaErrorInvalidValue;
        if (e == cudaSuccess) e = launch_filter_shadow_points(d_depth, height, width, 2, 2, -1.0f, filtered, scratch, g->stream);
        d_depth = filtered;
    }
    if (e == cudaSuccess) {
        RgbdParams P;
        P.fx_inv = 1.0 / K[0];
        P.fy_inv = 1.0 / K[1];
        P.cx = K[2];
        P.cy = K[3];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) P.R[3 * i + j] = Twc[4 * i + j];
            P.t[i] = Twc[4 * i + 3];
        }
        P.min_depth = min_depth;
        P.max_depth = max_depth;
        P.H = height;
        P.W = width;
        e = launch_grid_integrate_rgbd(P, d_depth, d_color, g->inv_voxel_size, g->table, g->meta, g->stream);
    }
    if (e == cudaSuccess && (tmp_d || tmp_c || filtered)) e = cudaStreamSynchronize(g->stream);
    cudaFree(tmp_d);
    cudaFree(tmp_c);
    cudaFree(filtered);
    cudaFree(scratch);
    if (e != cudaSuccess) {
        g->err = std::string("b2v_grid_integrate_rgbd: ") + cudaGetErrorString(e);
        return B2V_ERR_CUDA;
    }
    return B2V_OK;
}

extern "C" int b2v_grid_synchronize(b2v_grid *g) {
    if (!g) return B2V_ERR_INVALID_ARGUMENT;
    return grid_read_counters(g);
}

extern "C" int64_t b2v_grid_num_blocks(b2v_grid *g) {
    if (!g) return -1;
    if (grid_read_counters(g) == B2V_ERR_CUDA) return -1;
    return grid_block_count(g);
}

static int grid_ensure_scan(b2v_grid *g, uint32_t nb) {
    if (nb <= g->scan_cap) return B2V_OK;
    B2V_CUDA(g, regrow(&g->d_sums, static_cast<size_t>(nb)));
    B2V_CUDA(g, regrow(&g->d_offs, static_cast<size_t>(nb)));
    g->scan_cap = nb;
    return B2V_OK;
}

static int64_t grid_count(b2v_grid *g, int32_t min_count, uint32_t *nb_out) {
    if (grid_read_counters(g) == B2V_ERR_CUDA) return -1;
    const uint32_t nb = grid_block_count(g);
    if (nb_out) *nb_out = nb;
    if (grid_ensure_scan(g, nb) != B2V_OK) return -1;
    if (launch_grid_count(g->meta, nb, min_count, g->d_sums, g->d_offs, g->d_total, g->stream) != cudaSuccess)
        return -1;
    uint32_t total = 0;
    if (cudaMemcpyAsync(&total, g->d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream) != cudaSuccess)
        return -1;
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) return -1;
    return total;
}

extern "C" int64_t b2v_grid_size(b2v_grid *g) {
    if (!g) return -1;
    return grid_count(g, 1, nullptr);
}

extern "C" int64_t b2v_grid_get_voxels(b2v_grid *g, int32_t min_count) {
    if (!g) return -1;
    uint32_t nb = 0;
    const int64_t n = grid_count(g, min_count, &nb);
    if (n < 0) return -1;
    if (static_cast<size_t>(n) > g->out_cap) {
        if (regrow(&g->d_out_pts, static_cast<size_t>(n) * 3) != cudaSuccess) return -1;
        if (regrow(&g->d_out_cols, static_cast<size_t>(n) * 3) != cudaSuccess) return -1;
        g->out_cap = static_cast<size_t>(n);
    }
    if (launch_grid_emit(g->meta, nb, min_count, g->d_offs, g->d_out_pts, g->d_out_cols, g->stream) != cudaSuccess)
        return -1;
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) return -1;
    g->last_n = n;
    return n;
}

extern "C" int b2v_grid_copy_voxels(b2v_grid *g, float *points, float *colors) {
    if (!g) return B2V_ERR_INVALID_ARGUMENT;
    const size_t n = static_cast<size_t>(g->last_n);
    if (points && n) B2V_CUDA(g, cudaMemcpy(points, g->d_out_pts, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    if (colors && n) B2V_CUDA(g, cudaMemcpy(colors, g->d_out_cols, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return B2V_OK;
}

extern "C" int b2v_grid_remove_low_count_voxels(b2v_grid *g, int32_t min_count) {
    if (!g) return B2V_ERR_INVALID_ARGUMENT;
    const int rc = grid_read_counters(g);
    if (rc == B2V_ERR_CUDA) return rc;
    B2V_CUDA(g, launch_grid_remove_low_count(g->meta, grid_block_count(g), min_count, g->stream));
    return B2V_OK;
}

// frustum -> GridQuery: AABB of the 8 frustum corners (camera_frustrum.cpp:209-260) -> voxel key bounds in
// double (voxel_block_grid.hpp:1340-1345: get_voxel_key_inv<double,double> with the float inv_voxel_size)
static void fill_key_bounds(GridQuery *q, float inv_vs) {
    for (int a = 0; a < 3; ++a) {
        q->min_key[a] = static_cast<int32_t>(std::floor(q->bb[a] * static_cast<double>(inv_vs)));
        q->max_key[a] = static_cast<int32_t>(std::floor(q->bb[3 + a] * static_cast<double>(inv_vs)));
    }
}

void b2v::fill_frustum_query(GridQuery *q, const float K[4], int W, int H, const double Tcw[16],
                             float depth_max, float depth_min, int min_count, float inv_vs) {
    std::memset(q, 0, sizeof(*q));
    q->mode = 1;
    q->min_count = min_count;
    q->fx = K[0];
    q->fy = K[1];
    q->cx = K[2];
    q->cy = K[3];
    q->depth_min = depth_min;
    q->depth_max = depth_max;
    q->W = W;
    q->H = H;
    double Rwc[9], twc[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            q->R[3 * i + j] = Tcw[4 * i + j];
            Rwc[3 * i + j] = Tcw[4 * j + i];
        }
        q->t[i] = Tcw[4 * i + 3];
    }
    for (int i = 0; i < 3; ++i) twc[i] = -(Rwc[3 * i] * q->t[0] + Rwc[3 * i + 1] * q->t[1] + Rwc[3 * i + 2] * q->t[2]);
    for (int a = 0; a < 3; ++a) {
        q->bb[a] = 1e300;
        q->bb[3 + a] = -1e300;
    }
    const double us[4] = {0.0, static_cast<double>(W), static_cast<double>(W), 0.0};
    const double vs[4] = {0.0, 0.0, static_cast<double>(H), static_cast<double>(H)};
    for (int c = 0; c < 4; ++c) {
        const double xn = (us[c] - static_cast<double>(q->cx)) / static_cast<double>(q->fx);
        const double yn = (vs[c] - static_cast<double>(q->cy)) / static_cast<double>(q->fy);
        for (int far = 0; far < 2; ++far) {
            const double d = far ? static_cast<double>(depth_max) : static_cast<double>(depth_min);
            const double pc[3] = {xn * d, yn * d, d};
            for (int a = 0; a < 3; ++a) {
                const double w = Rwc[3 * a] * pc[0] + Rwc[3 * a + 1] * pc[1] + Rwc[3 * a + 2] * pc[2] + twc[a];
                q->bb[a] = std::min(q->bb[a], w);
                q->bb[3 + a] = std::max(q->bb[3 + a], w);
            }
        }
    }
    fill_key_bounds(q, inv_vs);
}

static int64_t grid_run_query(b2v_grid *g, const GridQuery &q) {
    if (grid_read_counters(g) == B2V_ERR_CUDA) return -1;
    const uint32_t nb = grid_block_count(g);
    if (grid_ensure_scan(g, nb) != B2V_OK) return -1;
    if (launch_grid_query_count(g->meta, nb, q, g->d_sums, g->d_offs, g->d_total, g->stream) != cudaSuccess) return -1;
    uint32_t total = 0;
    if (cudaMemcpyAsync(&total, g->d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) return -1;
    if (static_cast<size_t>(total) > g->out_cap) {
        if (regrow(&g->d_out_pts, static_cast<size_t>(total) * 3) != cudaSuccess) return -1;
        if (regrow(&g->d_out_cols, static_cast<size_t>(total) * 3) != cudaSuccess) return -1;
        g->out_cap = total;
    }
    if (launch_grid_query_emit(g->meta, nb, q, g->d_offs, g->d_out_pts, g->d_out_cols, g->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) return -1;
    g->last_n = total;
    return total;
}

extern "C" int64_t b2v_grid_get_voxels_in_frustum(b2v_grid *g, const float K[4], int32_t width, int32_t height,
                                                  const double Tcw[16], float depth_max, float depth_min,
                                                  int32_t min_count) {
    if (!g || !K || !Tcw || width <= 0 || height <= 0) return -1;
    GridQuery q;
    fill_frustum_query(&q, K, width, height, Tcw, depth_max, depth_min, min_count, g->inv_voxel_size);
    return grid_run_query(g, q);
}

extern "C" int64_t b2v_grid_get_voxels_in_bb(b2v_grid *g, const double bbox[6], int32_t min_count) {
    if (!g || !bbox) return -1;
    GridQuery q;
    std::memset(&q, 0, sizeof(q));
    q.mode = 0;
    q.min_count = min_count;
    for (int a = 0; a < 6; ++a) q.bb[a] = bbox[a];
    fill_key_bounds(&q, g->inv_voxel_size);
    return grid_run_query(g, q);
}

extern "C" int b2v_grid_carve(b2v_grid *g, const float K[4], int32_t width, int32_t height, const double Tcw[16],
                              float depth_max, float depth_min, const float *depth, float depth_threshold) {
    if (!g) return B2V_ERR_INVALID_ARGUMENT;
    if (!K || !Tcw || !depth || width <= 0 || height <= 0) {
        g->err = "b2v_grid_carve: bad arguments";
        return B2V_ERR_INVALID_ARGUMENT;
    }
    const int rc = grid_read_counters(g);
    if (rc == B2V_ERR_CUDA) return rc;
    const size_t pixels = static_cast<size_t>(width) * height;
    const float *d_depth = depth;
    float *tmp = nullptr;
    if (!is_device_pointer(depth)) {
        B2V_CUDA(g, cudaMalloc(&tmp, pixels * sizeof(float)));
        cudaError_t e = cudaMemcpyAsync(tmp, depth, pixels * sizeof(float), cudaMemcpyHostToDevice, g->stream);
        if (e != cudaSuccess) {
            cudaFree(tmp);
            g->err = cudaGetErrorString(e);
            return B2V_ERR_CUDA;
        }
        d_depth = tmp;
    }
    GridQuery q;
    fill_frustum_query(&q, K, width, height, Tcw, depth_max, depth_min, 1, g->inv_voxel_size);
    cudaError_t e = launch_grid_carve(g->meta, grid_block_count(g), q, d_depth, depth_threshold, g->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) {
        g->err = cudaGetErrorString(e);
        return B2V_ERR_CUDA;
    }
    return B2V_OK;
}

extern "C" int64_t b2v_grid_dump_blocks(b2v_grid *g, int32_t *keys, uint64_t *hashes, int32_t *count,
                                        float *pos_sum, float *col_sum) {
    if (!g) return -1;
    if (grid_read_counters(g) == B2V_ERR_CUDA) return -1;
    const uint32_t nb = grid_block_count(g);
    if (nb == 0) return 0;
    std::vector<int4> k(nb);
    if (cudaMemcpy(k.data(), g->meta.block_keys, nb * sizeof(int4), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    for (uint32_t i = 0; i < nb; ++i) {
        if (keys) {
            keys[3 * i + 0] = k[i].x;
            keys[3 * i + 1] = k[i].y;
            keys[3 * i + 2] = k[i].z;
        }
        if (hashes) hashes[i] = block_key_hash(k[i].x, k[i].y, k[i].z);
    }
    if (count || pos_sum || col_sum) {
        std::vector<uint32_t> raw(static_cast<size_t>(nb) * kGridBlockWordsHost);
        if (cudaMemcpy(raw.data(), g->meta.pool, raw.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
            return -1;
        for (size_t b = 0; b < nb; ++b) {
            const uint32_t *blk = raw.data() + b * kGridBlockWordsHost;
            const float *fb = reinterpret_cast<const float *>(blk);
            for (int l = 0; l < kVox; ++l) {
                if (count) count[b * kVox + l] = static_cast<int32_t>(blk[l]);
                for (int c = 0; c < 3; ++c) {
                    if (pos_sum) pos_sum[(b * kVox + l) * 3 + c] = fb[(1 + c) * kVox + l];
                    if (col_sum) col_sum[(b * kVox + l) * 3 + c] = fb[(4 + c) * kVox + l];
                }
            }
        }
    }
    return nb;
}
