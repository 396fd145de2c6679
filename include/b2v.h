/* b2v.h — C ABI of the B200-native volumetric integrator (libb2v.so).
 *
 * Drop-in boundary for pySLAM's dense-mapping plugin path.  Plain pointers and sizes only; no
 * torch / pybind types; status codes instead of exceptions.  Every entry point names the
 * reference interface it replaces (paths relative to the pySLAM tree).
 *
 * Conventions (same as the reference front-end):
 *   - depth  : float32 [H*W], metres, row-major           (pyslam/dense/volumetric_integrator_base.py:713)
 *   - color  : uint8   [H*W*3], RGB interleaved, row-major (base.py:1054)
 *   - K      : float64 [4] = {fx, fy, cx, cy}              (volumetric_integrator_tsdf.py:110-119)
 *   - Tcw    : float64 [16] row-major world->camera pose   (base.py:116; tsdf.py:223 `pose`)
 *   - image / point pointers may be HOST or DEVICE memory; the library detects which.
 *     Pinned host memory makes the host->device copies asynchronous.
 *   - all calls on one volume must come from one thread at a time (the reference's integrator
 *     process is single-consumer, base.py:789-967).
 */
#ifndef B2V_H
#define B2V_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2V_OK 0
#define B2V_ERR_INVALID_ARGUMENT 1
#define B2V_ERR_CUDA 2
#define B2V_ERR_CAPACITY 3 /* block pool or hash table full: raise capacity_blocks */
#define B2V_ERR_UNSUPPORTED 4

#define B2V_BLOCK_SIZE 8                 /* voxels per block side (config_parameters.py:313) */
#define B2V_BLOCK_VOXELS 512
#define B2V_VOXEL_PLANES 5               /* tsdf, weight, r, g, b : float32 planes per block */

typedef struct b2v_volume b2v_volume;    /* TSDF volume (Open3D-like duck type A)           */
typedef struct b2v_grid b2v_grid;        /* point-average voxel block grid (duck type B)    */

typedef struct b2v_config {
    float voxel_size;        /* kVolumetricIntegrationVoxelLength   (config_parameters.py:311) */
    int32_t block_size;      /* must be 8                            (config_parameters.py:313) */
    float sdf_trunc;         /* kVolumetricIntegrationTSdfTrunc      (config_parameters.py:349) */
    float depth_trunc;       /* ...TsdfDepthTruncIndoor/Outdoor      (config_parameters.py:350-351) */
    int32_t depth_stride;    /* Open3D depth_sampling_stride, default 4 (tsdf.py:104-108)      */
    uint32_t capacity_blocks;/* block-pool capacity (10 KiB per block)                          */
    int32_t device;          /* CUDA device ordinal                                             */
    int32_t shard_rank;      /* this GPU's shard; a block is owned iff                          */
    int32_t shard_count;     /*   BlockKeyHash(key) % shard_count == shard_rank (1 = own all)   */
    int32_t unit_resolution; /* Open3D volume_unit_resolution: 16 (0 = default; the reference's value,
                              * tsdf.py:104-108): allocation by ScalableTSDFVolume::LocateVolumeUnit, every 8^3
                              * block of a touched 16^3 unit; 8: SURVEY decision D1, allocation by the float32
                              * pyslam key range (voxel_hashing.h:69-75) of the +-sdf_trunc box               */
    double voxel_length;     /* the float64 voxel length / truncation Open3D holds (Python floats); 0 = widen  */
    double sdf_trunc_d;      /*   the float32 fields.  (float)voxel_length must equal voxel_size, same for tau */
} b2v_config;

/* ---- lifetime: replaces o3d.pipelines.integration.ScalableTSDFVolume(...) (tsdf.py:104-108) ---- */
int b2v_create(const b2v_config *cfg, b2v_volume **out);
int b2v_destroy(b2v_volume *v);
/* replaces self.volume.reset() (tsdf.py:156; base.py:642) */
int b2v_reset(b2v_volume *v);
const char *b2v_last_error(const b2v_volume *v);

/* ---- integrate: replaces self.volume.integrate(rgbd, intrinsic, pose) (tsdf.py:215-223) ----
 * Asynchronous: returns once the work is enqueued.  `stream` (a cudaStream_t) may be non-NULL only
 * with DEVICE image pointers; NULL uses the library's own streams. */
int b2v_integrate(b2v_volume *v, const float *depth, const uint8_t *color, int32_t height,
                  int32_t width, const double K[4], const double Tcw[16], void *stream);
/* Rectification on the GPU (volumetric_integrator_base.py:1017-1054): with maps installed, the frames given
 * to b2v_integrate / b2v_integrate_batch are the RAW (distorted) images; colour is remapped like
 * cv2.remap(..., INTER_LINEAR), depth like cv2.remap(..., INTER_NEAREST) (bit-exact with OpenCV's fixed-point
 * arithmetic, constant zero border) before allocation.  map_x / map_y: float32 [height*width] as produced by
 * cv2.initUndistortRectifyMap(..., CV_32FC1) (host); NULL maps remove the stage.  swap_rb != 0 also converts
 * BGR input to RGB (cv2.cvtColor(COLOR_BGR2RGB), base.py:1054).  Synchronises. */
int b2v_set_rectification(b2v_volume *v, const float *map_x, const float *map_y, int32_t height,
                          int32_t width, int32_t swap_rb);
/* stand-alone cv2.remap equivalents on host arrays (tests, other callers): kind 0 = uint8 x3 bilinear,
 * kind 1 = 32-bit pixels (float32 depth / int32 labels) nearest */
int b2v_remap(const void *src, int32_t kind, int32_t height, int32_t width, const float *map_x,
              const float *map_y, void *dst, int32_t swap_rb, int32_t device);
/* n frames back to back (the rebuild(map) bulk path, base.py:1242-1318): depth [n*H*W],
 * color [n*H*W*3], Tcw [n*16]; same K for all.  By default groups of up to 16 frames are FUSED: a block
 * is read once, updated by the frames of the group in frame order, and written once - bit-identical
 * to frame-by-frame integration (b2v_set_fusion(v, 0) forces frame-by-frame). */
int b2v_integrate_batch(b2v_volume *v, int32_t n_frames, const float *depth, const uint8_t *color,
                        int32_t height, int32_t width, const double K[4], const double *Tcw,
                        void *stream);
/* The same two calls for RAW 16-bit depth (TUM / ScanNet style PNG payloads): `depth` is uint16 [height][width]
 * (per frame), uploaded as is - 2 instead of 4 bytes per pixel over PCIe - and widened on the device to
 * float32(depth) * depth_scale in float32 arithmetic, the value numpy's `depth.astype(np.float32) * depth_factor`
 * produces in the reference (volumetric_integrator_base.py:1008-1015).  Everything else as above. */
int b2v_integrate_u16(b2v_volume *v, const uint16_t *depth, float depth_scale, const uint8_t *color, int32_t height,
                      int32_t width, const double K[4], const double Tcw[16], void *stream);
int b2v_integrate_batch_u16(b2v_volume *v, int32_t n_frames, const uint16_t *depth, float depth_scale,
                            const uint8_t *color, int32_t height, int32_t width, const double K[4], const double *Tcw,
                            void *stream);
/* Pipelined callers with DEVICE frames (pyslam_b200.sharding.FrameIngest): the next b2v_integrate_batch* call's inputs
 * are ready when `event` (a cudaEvent_t recorded by the producer of the frames) fires.  Without it the batch waits
 * for everything enqueued so far on the caller's stream - including the update kernels of the previous batch, which
 * its allocate kernels could overlap.  One-shot: consumed by the next batch call. */
int b2v_set_input_event(b2v_volume *v, void *event);
/* wait for all enqueued work; returns B2V_ERR_CAPACITY if a frame overflowed the pool */
int b2v_synchronize(b2v_volume *v);

/* ---- inspection / parity hooks ---- */
int64_t b2v_num_blocks(b2v_volume *v);                 /* synchronises */
/* blocks touched / newly allocated by the most recent frame (synchronises) */
int b2v_last_frame_stats(b2v_volume *v, int64_t *touched_blocks, int64_t *new_blocks);
/* bench accounting since create/reset: (block, frame) updates applied; kernel launches; block visits
 * (a visit = one block read + written; equals the updates frame by frame, fewer in fused batches) */
int b2v_counters(b2v_volume *v, int64_t *block_updates, int64_t *kernel_launches, int64_t *block_visits);
/* how the most recent b2v_extract_mesh / b2v_extract_points narrowed its work: stats[0] = blocks of the map, [1] = tiles
 * (a block + its +1 halo) whose blocks' sign summaries admit a surface crossing, [2] = tiles that hold both signs,
 * [3] = blocks with vertices, [4] = blocks with triangles */
int b2v_last_mesh_stats(b2v_volume *v, int64_t stats[5]);
/* Scheduling option: 1 (default) runs allocate(f+1) on its own stream concurrently with integrate(f)
 * (it has no data dependency on it); 0 serialises both kernels on one stream (clean per-kernel timing).
 * Results are bit-identical either way.  Synchronises. */
int b2v_set_overlap(b2v_volume *v, int32_t enable);
int b2v_set_fusion(b2v_volume *v, int32_t enable);
/* frames per fused group of b2v_integrate_batch: 1..32, default 16.  Larger groups amortise launch and latency costs
 * (hash-sharded ranks with few blocks each); results do not depend on it.  Synchronises. */
int b2v_set_group_size(b2v_volume *v, int32_t frames);
/* Per-kernel device timing (CUDA events on the launching stream around each launch), for the
 * roofline figure: enable, run frames, then read the summed durations (synchronises, resets). */
int b2v_profile_enable(b2v_volume *v, int32_t enable);
int b2v_profile_read(b2v_volume *v, double *allocate_ms, double *integrate_ms, int64_t *frames,
                     int64_t *integrate_launches);
/* keys int32[nb*3], hashes uint64[nb] (= reference BlockKeyHash, cpp/volumetric/voxel_hashing.h:106-113),
 * voxels float32[nb*5*512] (planes tsdf, weight, r, g, b; voxel index lx + 8*ly + 64*lz,
 * cpp/volumetric/voxel_block.h:67-70).  HOST outputs, any may be NULL; returns nb or <0. */
int64_t b2v_dump_blocks(b2v_volume *v, int32_t *keys, uint64_t *hashes, float *voxels);
/* Restore / seed blocks from HOST arrays keys int32[n*3] (unique), voxels float32[n*5*512]; existing
 * blocks are overwritten.  The reference's load() is a stub (base.py:595-604); this is the restore
 * half of b2v_dump_blocks, also used to gather shards onto one GPU and by the tests. */
int b2v_upload_blocks(b2v_volume *v, int64_t n_blocks, const int32_t *keys, const float *voxels);
/* the same exchange with DEVICE buffers (multi-GPU mesh gather): d_keys4 int32 [n][4] = {x, y, z, 0},
 * d_voxels float32 [n][5][512].  export returns the block count (with both pointers NULL: just the count). */
int64_t b2v_export_blocks_device(b2v_volume *v, int32_t *d_keys4, float *d_voxels, int64_t max_blocks);
int b2v_import_blocks_device(b2v_volume *v, int64_t n_blocks, const int32_t *d_keys4, const float *d_voxels);
/* keys int32[n*3] of the blocks touched by the most recent b2v_integrate frame - or, after b2v_integrate_batch, by
 * the frames of the batch's last fused group (their union; b2v_last_frame_stats still counts the last frame alone);
 * returns n or <0 */
int64_t b2v_last_touched_keys(b2v_volume *v, int32_t *keys, int64_t max_keys);

/* ---- mesh: replaces self.volume.extract_triangle_mesh() (tsdf.py:239,260) ----
 * Two-call pattern: b2v_extract_mesh runs the kernels and returns the sizes; b2v_copy_mesh copies
 * the result of the last extraction into HOST arrays vertices f64[nv*3], colors f64[nv*3] in [0,1] (float64 like
 * Open3D's TriangleMesh, computed with Open3D's float64 formulas), edge_ids int32[nv*4] (canonical weld key:
 * voxel x,y,z + axis), triangles int32[nt*3]. */
int b2v_extract_mesh(b2v_volume *v, int64_t *n_vertices, int64_t *n_triangles);
int b2v_copy_mesh(b2v_volume *v, double *vertices, double *colors, int32_t *edge_ids,
                  int32_t *triangles);
/* replaces self.volume.extract_point_cloud() (tsdf.py:246,267): zero crossings along +x,+y,+z */
int b2v_extract_points(b2v_volume *v, int64_t *n_points);
int b2v_copy_points(b2v_volume *v, double *points, double *colors);   /* float64 [n*3], Open3D's formulas */

/* ---- duck type B: pySLAM's own volumetric.VoxelBlockGrid (point-average grid) ----
 * replaces VoxelBlockGridT<VoxelData> (cpp/volumetric/voxel_block_grid.h:61-233) behind the pybind
 * class registered at cpp/volumetric/volumetric_grid_module.h:732-935. */
int b2v_grid_create(float voxel_size, int32_t block_size, uint32_t capacity_blocks, int32_t device,
                    b2v_grid **out);
int b2v_grid_destroy(b2v_grid *g);
int b2v_grid_clear(b2v_grid *g);                       /* clear()/reset() */
const char *b2v_grid_last_error(const b2v_grid *g);
/* integrate(points f32[n*3], colors f32[n*3] | NULL)  (volumetric_grid_module.h:131-467 ->
 * voxel_block_grid.hpp:115-136) */
int b2v_grid_integrate(b2v_grid *g, const float *points, const float *colors, int64_t n_points);
/* the float64-points overload (volumetric_grid_module.h:737-749): voxel keys from the float64 coordinates
 * (floor(x * (double)inv_voxel_size), voxel_hashing.h:69-75), sums accumulate static_cast<float>(x) */
int b2v_grid_integrate_f64(b2v_grid *g, const double *points, const float *colors, int64_t n_points);
/* every dtype combination of the pybind overloads (volumetric_grid_module.h:737-802): points float32 | float64,
 * colours float32 | uint8 (scaled on the device by the float32 constant 1/255, voxel_data.h:79-97) | NULL */
int b2v_grid_integrate_ex(b2v_grid *g, const void *points, int32_t points_f64, const void *colors, int32_t colors_u8,
                          int64_t n_points);
/* Fused front-end of VolumetricIntegratorVoxelGrid: depth2pointcloud (pyslam/utilities/depth.py:45-85) +
 * world transform + integrate (pyslam/dense/volumetric_integrator_voxel_grid.py:247-300) in one call, no
 * point cloud materialised.  depth float32 [H*W], color uint8 RGB [H*W*3] (host or device), K = {fx,fy,cx,cy}
 * float64, Twc float64[16] row-major camera->world (the reference's inv_T(pose)), valid pixels are
 * min_depth < d < max_depth.  Same keys / counts as integrating the front-end's float32 points. */
int b2v_grid_integrate_rgbd(b2v_grid *g, const float *depth, const uint8_t *color, int32_t height,
                            int32_t width, const double K[4], const double Twc[16], float max_depth,
                            float min_depth, int32_t filter_shadow_points);
/* filter_shadow_points(depth, delta_depth=None, delta_x, delta_y, fill_value) (pyslam/utilities/depth.py:103-146)
 * on the GPU: exact global median (radix select) of the positive depth differences, threshold
 * 3 * 1.4826 * median, pixels on either side of a larger jump are set to fill_value.
 * depth / out: float32 [height*width], host or device (out may alias depth only on the host). */
int b2v_filter_shadow_points(const float *depth, int32_t height, int32_t width, int32_t delta_x,
                             int32_t delta_y, float fill_value, float *out, int32_t device);
int b2v_grid_synchronize(b2v_grid *g);
int64_t b2v_grid_num_blocks(b2v_grid *g);              /* num_blocks() */
int64_t b2v_grid_size(b2v_grid *g);                    /* size(): voxels with count > 0 */
/* get_voxels(min_count) (voxel_block_grid.hpp:717-819): returns n; then copy */
int64_t b2v_grid_get_voxels(b2v_grid *g, int32_t min_count);
int b2v_grid_copy_voxels(b2v_grid *g, float *points, float *colors);
/* remove_low_count_voxels(min_count) (voxel_block_grid.hpp:625-647) */
int b2v_grid_remove_low_count_voxels(b2v_grid *g, int32_t min_count);
/* carve(camera_frustrum, depth_image, depth_threshold) (voxel_block_grid.hpp:616-622;
 * voxel_grid_carving.h:47-80; CameraFrustrum: camera_frustrum.h:36-48): K = {fx,fy,cx,cy} float32,
 * Tcw float64[16] row-major, depth float32 [height*width] (host or device). */
int b2v_grid_carve(b2v_grid *g, const float K[4], int32_t width, int32_t height, const double Tcw[16],
                   float depth_max, float depth_min, const float *depth, float depth_threshold);
/* get_voxels_in_camera_frustrum(frustum, min_count) (voxel_block_grid.hpp:1019-1195) and
 * get_voxels_in_bb(bbox, min_count) (voxel_block_grid.hpp:822-1016), bbox = {min xyz, max xyz} float64.
 * Return n; fetch with b2v_grid_copy_voxels. */
int64_t b2v_grid_get_voxels_in_frustum(b2v_grid *g, const float K[4], int32_t width, int32_t height,
                                       const double Tcw[16], float depth_max, float depth_min,
                                       int32_t min_count);
int64_t b2v_grid_get_voxels_in_bb(b2v_grid *g, const double bbox[6], int32_t min_count);
/* parity hook: keys int32[nb*3], hashes u64[nb], count int32[nb*512], pos_sum f32[nb*512*3],
 * col_sum f32[nb*512*3]  (same layout as the reference's VoxelData, voxel_data.h:118-133) */
int64_t b2v_grid_dump_blocks(b2v_grid *g, int32_t *keys, uint64_t *hashes, int32_t *count,
                             float *pos_sum, float *col_sum);

/* ---- semantic voxel-block grids -------------------------------------------
 * Drop-in for volumetric.VoxelBlockSemanticGrid (voting) and volumetric.VoxelBlockSemanticProbabilisticGrid
 * (cpp/volumetric/voxel_block_semantic_grid.h:59-121; pybind: volumetric_grid_module.h), for
 *   integrate(points, colors, class_ids, instance_ids, depths)   voxel_block_grid.hpp:12-112
 *   get_voxels(min_count, min_confidence)                        voxel_block_grid.hpp:717-819
 *   set_depth_threshold / set_depth_decay_rate                   voxel_block_semantic_grid.hpp:22-36
 *   remove_low_count_voxels, remove_low_confidence_segments, merge_segments, remove_segment
 *                                                                voxel_block_grid.hpp:625-647, semantic_grid.hpp:101-183
 * Observations reach a voxel in input order, like the reference's sequential build: counts, float64 position
 * sums, float32 colour sums, labels and log-evidence are bit-identical.  The depth threshold / decay rate are
 * per grid here (class-static, i.e. process-wide, in the reference: voxel_data_semantic.h:107-108, 251-254). */
typedef struct b2v_sgrid b2v_sgrid;
#define B2V_SEM_VOTING 0         /* VoxelSemanticData: (object, class, counter), voxel_data_semantic.h:106-199 */
#define B2V_SEM_PROBABILISTIC 1  /* VoxelSemanticDataProbabilistic: joint log-evidence per pair, :249-672 */
#define B2V_SEM_MAX_LABELS 8     /* label pairs kept per Bayesian voxel (the reference's map is unbounded) */
int b2v_sgrid_create(double voxel_size, int32_t block_size, uint32_t capacity_blocks, int32_t kind,
                     int32_t device, b2v_sgrid **out);
int b2v_sgrid_destroy(b2v_sgrid *g);
const char *b2v_sgrid_last_error(const b2v_sgrid *g);
int b2v_sgrid_clear(b2v_sgrid *g);
int b2v_sgrid_set_depth_threshold(b2v_sgrid *g, float depth_threshold);
int b2v_sgrid_set_depth_decay_rate(b2v_sgrid *g, float depth_decay_rate);
/* points: float32 or float64 [n][3] (points_f64); colors: NULL, float32 [n][3] in [0,1] or uint8 [n][3]
 * (colors_u8); class_ids / instance_ids / depths: NULL or [n].  Without colours only positions are integrated
 * and without instance ids the object id is 0, as in the reference (voxel_block_grid.hpp:228-231, 259-286).
 * Host or device pointers; synchronous (the inputs are free when the call returns). */
int b2v_sgrid_integrate(b2v_sgrid *g, int64_t n, const void *points, int32_t points_f64, const void *colors,
                        int32_t colors_u8, const int32_t *class_ids, const int32_t *instance_ids,
                        const float *depths);
/* Fused front-end of the semantic integrator (volumetric_integrator_voxel_semantic_grid.py:332-461): optional
 * shadow-point filter, depth2pointcloud (depth.py:45-85) with class / object-id images, camera->world transform
 * (Twc, float64 then float32) and integrate, without materialising the point cloud.  Points reach a voxel in
 * row-major pixel order, the order of the reference's point arrays.  class_image / object_image: NULL or int32
 * [H][W]; use_depths: weight the evidence by camera depth (kVolumetricSemanticProbabilisticIntegrationUseDepth). */
int b2v_sgrid_integrate_rgbd(b2v_sgrid *g, const float *depth, const uint8_t *color, const int32_t *class_image,
                             const int32_t *object_image, int32_t height, int32_t width, const double K[4],
                             const double Twc[16], float max_depth, float min_depth, int32_t use_depths,
                             int32_t filter_shadow_points);
int64_t b2v_sgrid_num_blocks(b2v_sgrid *g);
/* two-step read-out: get_voxels returns the count (or -1), copy_voxels fills caller arrays (any may be NULL):
 * points f64 [n][3], colors f32 [n][3], class_ids / object_ids i32 [n], confidences f32 [n] */
int64_t b2v_sgrid_get_voxels(b2v_sgrid *g, int32_t min_count, float min_confidence);
int b2v_sgrid_copy_voxels(b2v_sgrid *g, double *points, float *colors, int32_t *class_ids, int32_t *object_ids,
                          float *confidences);
/* spatial read-outs, same two-step pattern (voxel_block_grid.hpp:822-1016 get_voxels_in_bb, bbox = min xyz, max xyz;
 * :1019-1195 get_voxels_in_camera_frustrum) */
int64_t b2v_sgrid_get_voxels_in_bb(b2v_sgrid *g, const double bbox[6], int32_t min_count, float min_confidence);
int64_t b2v_sgrid_get_voxels_in_frustum(b2v_sgrid *g, const float K[4], int32_t width, int32_t height,
                                        const double Tcw[16], float depth_max, float depth_min, int32_t min_count,
                                        float min_confidence);
int b2v_sgrid_remove_low_count_voxels(b2v_sgrid *g, int32_t min_count);
int b2v_sgrid_remove_low_confidence_segments(b2v_sgrid *g, int32_t min_confidence);
int b2v_sgrid_merge_segments(b2v_sgrid *g, int32_t object_id1, int32_t object_id2);
int b2v_sgrid_remove_segment(b2v_sgrid *g, int32_t object_id);
/* carve (voxel_block_grid.hpp:616-622; voxel_grid_carving.h:47-80) on a semantic grid: K = {fx, fy, cx, cy} float,
 * Tcw row-major 4x4, depth float32 [height][width] (host or device) */
int b2v_sgrid_carve(b2v_sgrid *g, const float K[4], int32_t width, int32_t height, const double Tcw[16],
                    float depth_max, float depth_min, const float *depth, float depth_threshold);
/* assign_object_ids_to_instance_ids (voxel_block_semantic_grid.h:67-71; voxel_semantic_data_association.h:69-373):
 * voxels in the frustum whose class equals the pixel's class and that lie on the observed surface vote
 * "2-D instance id -> 3-D object id"; returns the number of (instance, object) pairs of the resulting map, or -1;
 * b2v_sgrid_copy_instance_map copies them out (ascending instance id; object id -1 = no confident match).
 * class_image / instance_image: int32 [height][width]; depth_image: float32 or NULL.  New object ids come from a
 * per-grid counter (process-wide in the reference, voxel_semantic_shared_data.h:27-33) handed out in ascending
 * instance-id order (block-iteration order in the reference): maps agree up to that renumbering. */
int64_t b2v_sgrid_assign_object_ids_to_instance_ids(b2v_sgrid *g, const float K[4], int32_t width, int32_t height,
                                                    const double Tcw[16], float depth_max, float depth_min,
                                                    const int32_t *class_image, const int32_t *instance_image,
                                                    const float *depth_image, float depth_threshold,
                                                    int32_t do_carving, float min_vote_ratio, int32_t min_votes);
int b2v_sgrid_copy_instance_map(b2v_sgrid *g, int32_t *instance_ids, int32_t *object_ids);
int b2v_sgrid_set_next_object_id(b2v_sgrid *g, int32_t next_object_id);
int32_t b2v_sgrid_get_next_object_id(const b2v_sgrid *g);
/* number of label pairs dropped because a Bayesian voxel saw more than B2V_SEM_MAX_LABELS distinct pairs */
int b2v_sgrid_label_overflows(b2v_sgrid *g, uint64_t *out);
/* parity hook: arrays [nb][512]...; aux = voting counter / number of label pairs; lab_* [nb][512][K] in
 * ascending (object, class) order padded with (-1, -1, -inf); any output may be NULL */
int64_t b2v_sgrid_dump_blocks(b2v_sgrid *g, int32_t *keys, uint64_t *hashes, int32_t *count, double *pos_sum,
                              float *col_sum, int32_t *object_id, int32_t *class_id, float *confidence,
                              int32_t *aux, int32_t K, int32_t *lab_obj, int32_t *lab_cls, float *lab_logp);

/* Self-test of the update kernels' IEEE division fast path (shared correctly rounded reciprocal + two residual
 * corrections instead of the compiler's div.rn expansion): counts inputs whose result differs from __frcp_rn over all
 * 2^23 significands (x3 exponents) and from __fdiv_rn over `pairs` pseudo-random operand pairs.  Both must be 0. */
int b2v_selftest_division(int32_t device, uint64_t pairs, uint64_t *bad_reciprocals, uint64_t *bad_quotients);

/* library / device info */
int b2v_version(void);
int b2v_device_sm_count(int32_t device);

#ifdef __cplusplus
}
#endif
#endif /* B2V_H */
