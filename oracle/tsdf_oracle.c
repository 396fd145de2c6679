/* TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product; never linked, imported or
 * executed by pyslam_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file.
 *
 * PARITY vs a RUNNING Open3D is unpinned (pinned against oracle/open3d_order.c, the literal restatement of the
 * upstream source, instead): the reference's TSDF engine is Open3D's legacy
 * ScalableTSDFVolume (called from /root/reference/pyslam/dense/volumetric_integrator_tsdf.py:104-108,
 * 215-223,239,260), an un-vendored dependency (pin 02674268f706be4b004bbbf3d39b95fa9de35f74,
 * /root/reference/scripts/install_open3d_python.sh:114-118; conda open3d-0.19.0) that is absent
 * from /root/reference and not installed here, and the reference holds no golden vector for it
 * (SURVEY.md §8c).  This file restates Open3D's published algorithm
 * under decision D1: pyslam key arithmetic and 8^3 blocks.
 * The key / block / hash half IS pinned: tests check it against the compiled, unmodified
 * reference (oracle/_ref/libref_grid.so) and the in-header floor_div table
 * (/root/reference/cpp/volumetric/voxel_hashing.h:129-142).
 *
 * Arithmetic contract (shared with the CUDA kernels so values can be compared bit-exactly;
 * compile with -ffp-contract=off, every fused multiply-add is an explicit fmaf()):
 *
 *   keys      v = (int32)floorf(x_f32 * inv_vs_f32)            voxel_hashing.h:69-75
 *             b = floor_div(v, B), l = v - b*B                  voxel_hashing.h:139-161
 *             voxel flat index lx + ly*B + lz*B^2               voxel_block.h:67-70
 *             hash = (u64)(i64)x ^ ((u64)(i64)y << 1) ^ ((u64)(i64)z << 2)   voxel_hashing.h:106-113
 *   allocate  (A.2) every `stride`-th pixel with 0 < d < depth_trunc: back-project in f64,
 *             p_w = Twc * p_c (rigid inverse of Tcw), touch every block in the key range of
 *             [p_w - tau, p_w + tau]
 *   update    (A.3) OPEN3D'S OPERATION ORDER (contract v3): a block is a sub-block of its R^3 volume unit; voxel
 *             centre h = (float)((double)(vl/2 + vl*x) + unit*L) (z: the unit's first voxel), p = ((E0*h0 + E1*h1)
 *             + E2*h2) + E3 without FMA, then p += vl*E[:,2] once per z step from the unit's z = 0;
 *             u_f = p.x*fx/p.z + cx + 0.5f with true divisions; Open3D's 0.0001 image margin;
 *             sdf = (d - p.z) * lambda(u,v);  if sdf > -tau:  t = min(1, sdf*(1/tau)),
 *             tsdf = (tsdf*w + t)/(w+1) (mul, add, div);  w += 1.  In unit-16 mode tsdf and weight are BIT-IDENTICAL
 *             to oracle/open3d_order.c (tests/test_oracle_open3d.py).  Colour is a float32 running mean
 *             rgb_k = fmaf(rgb_k,w,RGB_k)*(1/(w+1)) where Open3D keeps float64 (compared with a tolerance).
 *   mesh      (A.4) classic tables; vertex on edge (voxel e, axis a) in float64 like Open3D:
 *             pt = vl/2 + vl*e;  pt[a] += |f0|*vl/(|f0|+|f1|);  colour (|f1|*c0/255 + |f0|*c1/255)/(|f0|+|f1|)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "mc_tables.h"

typedef struct {
    int32_t x, y, z;
} key3;

typedef struct tsdf_oracle {
    float vs, inv_vs, tau, inv_tau, depth_trunc;
    int B, nvox, stride;
    /* block pool: vox[b][5][nvox] planes = tsdf, weight, r, g, b (colour range 0..255) */
    int64_t nb, cap;
    key3 *keys;
    float *vox;
    int64_t *stamp; /* frame id of last touch */
    /* key -> block index map (open addressing, linear probing) */
    int64_t tcap;
    int64_t *table;
    /* blocks touched by the last integrate call, in first-touch order */
    int64_t ntouched, touched_cap;
    int64_t *touched;
    int64_t frame;
    /* Open3D volume units of R = 8*unit_blocks voxels.  R = 16 (the reference's setting): allocation by
     * LocateVolumeUnit in float64, every block of a touched unit.  R = 8 (decision D1): allocation by the float32
     * pyslam key range of the +-tau box; the update arithmetic is Open3D's with 8^3 units. */
    int unit_blocks;
    double unit_len, tau_d;
} tsdf_oracle;

static inline uint64_t mix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}

static inline uint64_t slot_hash(key3 k) {
    uint64_t h = (uint64_t)(uint32_t)k.x * 0x9E3779B97F4A7C15ULL;
    h ^= mix64((uint64_t)(uint32_t)k.y + 0x632BE59BD9B4E019ULL);
    h = mix64(h ^ ((uint64_t)(uint32_t)k.z << 21));
    return h;
}

/* voxel_hashing.h:106-113 with std::hash<int32_t> = identity (sign-extending) */
uint64_t tsdf_oracle_block_key_hash(int32_t x, int32_t y, int32_t z) {
    const uint64_t h1 = (uint64_t)(int64_t)x;
    const uint64_t h2 = (uint64_t)(int64_t)y;
    const uint64_t h3 = (uint64_t)(int64_t)z;
    return h1 ^ (h2 << 1) ^ (h3 << 2);
}

/* voxel_hashing.h:139-142 */
int64_t tsdf_oracle_floor_div(int64_t a, int64_t b) { return (a >= 0) ? (a / b) : ((a - b + 1) / b); }

/* voxel_hashing.h:69-75 (Tp = Tv = float) */
int32_t tsdf_oracle_voxel_coord(float x, float inv_vs) { return (int32_t)floorf(x * inv_vs); }

static void table_rebuild(tsdf_oracle *o, int64_t tcap) {
    free(o->table);
    o->tcap = tcap;
    o->table = (int64_t *)malloc(sizeof(int64_t) * (size_t)tcap);
    for (int64_t i = 0; i < tcap; ++i) o->table[i] = -1;
    for (int64_t b = 0; b < o->nb; ++b) {
        uint64_t s = slot_hash(o->keys[b]) & (uint64_t)(tcap - 1);
        while (o->table[s] >= 0) s = (s + 1) & (uint64_t)(tcap - 1);
        o->table[s] = b;
    }
}

static int64_t block_find(const tsdf_oracle *o, key3 k) {
    uint64_t s = slot_hash(k) & (uint64_t)(o->tcap - 1);
    for (;;) {
        const int64_t b = o->table[s];
        if (b < 0) return -1;
        if (o->keys[b].x == k.x && o->keys[b].y == k.y && o->keys[b].z == k.z) return b;
        s = (s + 1) & (uint64_t)(o->tcap - 1);
    }
}

static int64_t block_find_or_create(tsdf_oracle *o, key3 k) {
    int64_t b = block_find(o, k);
    if (b >= 0) return b;
    if (o->nb == o->cap) {
        o->cap = o->cap ? o->cap * 2 : 1024;
        o->keys = (key3 *)realloc(o->keys, sizeof(key3) * (size_t)o->cap);
        o->vox = (float *)realloc(o->vox, sizeof(float) * 5 * (size_t)o->nvox * (size_t)o->cap);
        o->stamp = (int64_t *)realloc(o->stamp, sizeof(int64_t) * (size_t)o->cap);
    }
    b = o->nb++;
    o->keys[b] = k;
    o->stamp[b] = -1;
    memset(o->vox + (size_t)b * 5 * o->nvox, 0, sizeof(float) * 5 * (size_t)o->nvox);
    if (o->nb * 2 > o->tcap) {
        table_rebuild(o, o->tcap * 2);
    } else {
        uint64_t s = slot_hash(k) & (uint64_t)(o->tcap - 1);
        while (o->table[s] >= 0) s = (s + 1) & (uint64_t)(o->tcap - 1);
        o->table[s] = b;
    }
    return b;
}

tsdf_oracle *tsdf_oracle_create(float voxel_size, int block_size, float sdf_trunc, float depth_trunc,
                                int stride) {
    tsdf_oracle *o = (tsdf_oracle *)calloc(1, sizeof(tsdf_oracle));
    o->vs = voxel_size;
    o->inv_vs = 1.0f / voxel_size; /* voxel_block_grid.hpp:6 */
    o->tau = sdf_trunc;
    o->inv_tau = 1.0f / sdf_trunc;
    o->depth_trunc = depth_trunc;
    o->B = block_size;
    o->nvox = block_size * block_size * block_size;
    o->stride = stride < 1 ? 1 : stride;
    o->unit_blocks = 1;
    o->unit_len = (double)voxel_size * (double)block_size;
    o->tau_d = (double)sdf_trunc;
    table_rebuild(o, 1 << 12);
    return o;
}

/* Open3D allocation granularity: unit_resolution voxels per volume-unit side (a multiple of the block side; 16 in
 * the reference, volumetric_integrator_tsdf.py:104-108; 8 = decision D1), voxel_length / sdf_trunc as the float64
 * values Open3D holds. */
void tsdf_oracle_set_units(tsdf_oracle *o, int unit_resolution, double voxel_length, double sdf_trunc) {
    o->unit_blocks = unit_resolution > o->B ? unit_resolution / o->B : 1;
    o->unit_len = voxel_length * (double)(o->unit_blocks * o->B);
    o->tau_d = sdf_trunc;
}

void tsdf_oracle_destroy(tsdf_oracle *o) {
    if (!o) return;
    free(o->keys);
    free(o->vox);
    free(o->stamp);
    free(o->table);
    free(o->touched);
    free(o);
}

void tsdf_oracle_reset(tsdf_oracle *o) {
    o->nb = 0;
    o->ntouched = 0;
    o->frame = 0;
    table_rebuild(o, 1 << 12);
}

int64_t tsdf_oracle_num_blocks(const tsdf_oracle *o) { return o->nb; }
int64_t tsdf_oracle_num_touched(const tsdf_oracle *o) { return o->ntouched; }

static void touch(tsdf_oracle *o, int64_t b) {
    if (o->stamp[b] == o->frame) return;
    o->stamp[b] = o->frame;
    if (o->ntouched == o->touched_cap) {
        o->touched_cap = o->touched_cap ? o->touched_cap * 2 : 4096;
        o->touched = (int64_t *)realloc(o->touched, sizeof(int64_t) * (size_t)o->touched_cap);
    }
    o->touched[o->ntouched++] = b;
}

static inline int depth_valid(float d, float depth_trunc) { return d > 0.0f && d < depth_trunc; }

/* A.2: allocation / touched set of one frame. */
static void allocate_frame(tsdf_oracle *o, const float *depth, int H, int W, const double K[4],
                           const double Tcw[16]) {
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    /* rigid inverse: Rwc = Rcw^T, twc = -(Rwc * tcw) */
    double R[3][3], t[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = Tcw[4 * j + i];
    for (int i = 0; i < 3; ++i)
        t[i] = -((R[i][0] * Tcw[3] + R[i][1] * Tcw[7]) + R[i][2] * Tcw[11]);
    const double tau = o->unit_blocks > 1 ? o->tau_d : (double)o->tau;
    const int B = o->B;
    for (int i = 0; i < H; i += o->stride) {
        for (int j = 0; j < W; j += o->stride) {
            const float d = depth[(size_t)i * W + j];
            if (!depth_valid(d, o->depth_trunc)) continue;
            const double z = (double)d;
            const double x = ((double)j - cx) * z / fx;
            const double y = ((double)i - cy) * z / fy;
            int32_t lo[3], hi[3];
            for (int a = 0; a < 3; ++a) {
                const double pw = ((R[a][0] * x + R[a][1] * y) + R[a][2] * z) + t[a];
                if (o->unit_blocks > 1) { /* ScalableTSDFVolume::LocateVolumeUnit: floor(p / unit length), float64 */
                    const int32_t ulo = (int32_t)floor((pw - tau) / o->unit_len);
                    const int32_t uhi = (int32_t)floor((pw + tau) / o->unit_len);
                    lo[a] = ulo * o->unit_blocks;
                    hi[a] = uhi * o->unit_blocks + o->unit_blocks - 1;
                    continue;
                }
                const int32_t vlo = tsdf_oracle_voxel_coord((float)(pw - tau), o->inv_vs);
                const int32_t vhi = tsdf_oracle_voxel_coord((float)(pw + tau), o->inv_vs);
                lo[a] = (int32_t)tsdf_oracle_floor_div(vlo, B);
                hi[a] = (int32_t)tsdf_oracle_floor_div(vhi, B);
            }
            for (int32_t bx = lo[0]; bx <= hi[0]; ++bx)
                for (int32_t by = lo[1]; by <= hi[1]; ++by)
                    for (int32_t bz = lo[2]; bz <= hi[2]; ++bz) {
                        const key3 k = {bx, by, bz};
                        touch(o, block_find_or_create(o, k));
                    }
        }
    }
}

/* A.3: projective update of one block, in Open3D's operation order
 * (UniformTSDFVolume::IntegrateWithDepthToCameraDistanceMultiplier) on the voxels of the block:
 * the block is sub-block (s = key - unit*S) of the volume unit `unit = floor_div(key, S)`; voxel (x,y,z) of the
 * unit has x = 8*s.x + lx etc.  Returns the number of voxels updated. */
static int64_t integrate_block(float *vox, key3 key, int B, int S, double unit_len, float vs, float tau,
                               float depth_trunc, const float *depth, const uint8_t *rgb, int H, int W, float fx,
                               float fy, float cxf, float cyf, const float E[12]) {
    const int nvox = B * B * B;
    float *p_tsdf = vox, *p_w = vox + nvox, *p_r = vox + 2 * nvox, *p_g = vox + 3 * nvox, *p_b = vox + 4 * nvox;
    const float inv_fx = 1.0f / fx, inv_fy = 1.0f / fy;
    const float half = vs * 0.5f, inv_tau = 1.0f / tau;
    const float safe_w = (float)W - 0.0001f, safe_h = (float)H - 0.0001f;
    const float Es[3] = {E[2] * vs, E[6] * vs, E[10] * vs}; /* extrinsic_scaled_f(:, 2) */
    int32_t u[3], sb[3];
    const int32_t k3[3] = {key.x, key.y, key.z};
    for (int a = 0; a < 3; ++a) {
        u[a] = (int32_t)tsdf_oracle_floor_div(k3[a], S);
        sb[a] = k3[a] - u[a] * S;
    }
    const double origin[3] = {(double)u[0] * unit_len, (double)u[1] * unit_len, (double)u[2] * unit_len};
    int64_t updated = 0;
    for (int lx = 0; lx < B; ++lx) {
        const float h0 = (float)((double)(half + vs * (float)(sb[0] * B + lx)) + origin[0]);
        for (int ly = 0; ly < B; ++ly) {
            const float h1 = (float)((double)(half + vs * (float)(sb[1] * B + ly)) + origin[1]);
            const float h2 = (float)((double)half + origin[2]);
            float pc[3];
            for (int r = 0; r < 3; ++r)
                pc[r] = ((E[4 * r + 0] * h0 + E[4 * r + 1] * h1) + E[4 * r + 2] * h2) + E[4 * r + 3];
            /* the unit's z loop reaches this block after sb.z * B increments */
            for (int z = 0; z < sb[2] * B; ++z) {
                pc[0] += Es[0];
                pc[1] += Es[1];
                pc[2] += Es[2];
            }
            for (int lz = 0; lz < B; ++lz, pc[0] += Es[0], pc[1] += Es[1], pc[2] += Es[2]) {
                if (pc[2] <= 0) continue;
                const float u_f = pc[0] * fx / pc[2] + cxf + 0.5f;
                const float v_f = pc[1] * fy / pc[2] + cyf + 0.5f;
                if (!(u_f >= 0.0001f && u_f < safe_w && v_f >= 0.0001f && v_f < safe_h)) continue;
                const int uu = (int)u_f, vv = (int)v_f;
                const float d = depth[(size_t)vv * W + uu];
                if (!depth_valid(d, depth_trunc)) continue;
                const float xx = ((float)uu - cxf) * inv_fx;
                const float yy = ((float)vv - cyf) * inv_fy;
                const float lam = sqrtf(xx * xx + yy * yy + 1.0f);
                const float sdf = (d - pc[2]) * lam;
                if (sdf > -tau) {
                    const int idx = lx + ly * B + lz * B * B;
                    const float tval = fminf(1.0f, sdf * inv_tau);
                    const float w = p_w[idx];
                    const float wn = w + 1.0f;
                    const float r = 1.0f / wn;
                    const uint8_t *c = rgb + ((size_t)vv * W + uu) * 3;
                    p_tsdf[idx] = (p_tsdf[idx] * w + tval) / wn;
                    /* colour: float32 running mean (Open3D keeps float64; compared with a tolerance) */
                    p_r[idx] = fmaf(p_r[idx], w, (float)c[0]) * r;
                    p_g[idx] = fmaf(p_g[idx], w, (float)c[1]) * r;
                    p_b[idx] = fmaf(p_b[idx], w, (float)c[2]) * r;
                    p_w[idx] = wn;
                    ++updated;
                }
            }
        }
    }
    return updated;
}

/* integrate(depth f32[H*W] metres, rgb u8[H*W*3], K = {fx,fy,cx,cy} f64, Tcw f64[16] row-major).
 * nthreads <= 1: scalar; otherwise OpenMP over the touched blocks (Open3D itself is OpenMP).
 * Returns the number of blocks touched by this frame. */
int64_t tsdf_oracle_integrate(tsdf_oracle *o, const float *depth, const uint8_t *rgb, int H, int W,
                              const double K[4], const double Tcw[16], int nthreads) {
    o->ntouched = 0;
    o->frame += 1;
    allocate_frame(o, depth, H, W, K, Tcw);
    float E[12];
    for (int i = 0; i < 12; ++i) E[i] = (float)Tcw[i];
    const float fx = (float)K[0], fy = (float)K[1], cx = (float)K[2], cy = (float)K[3];
    const int64_t n = o->ntouched;
    const size_t bstride = (size_t)5 * o->nvox;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 32) num_threads(nthreads > 1 ? nthreads : 1)
#endif
    for (int64_t i = 0; i < n; ++i) {
        const int64_t b = o->touched[i];
        integrate_block(o->vox + bstride * (size_t)b, o->keys[b], o->B, o->unit_blocks, o->unit_len, o->vs, o->tau,
                        o->depth_trunc, depth, rgb, H, W, fx, fy, cx, cy, E);
    }
    return n;
}

/* test hook: overwrite (creating it if needed) the planes of one block with vox[5*nvox] */
void tsdf_oracle_set_block(tsdf_oracle *o, int32_t bx, int32_t by, int32_t bz, const float *vox) {
    const key3 k = {bx, by, bz};
    const int64_t b = block_find_or_create(o, k);
    memcpy(o->vox + (size_t)b * 5 * o->nvox, vox, sizeof(float) * 5 * (size_t)o->nvox);
}

/* keys int32[n,3] of the blocks touched by the last frame */
int64_t tsdf_oracle_last_touched(const tsdf_oracle *o, int32_t *keys) {
    for (int64_t i = 0; i < o->ntouched; ++i) {
        const key3 k = o->keys[o->touched[i]];
        keys[3 * i + 0] = k.x;
        keys[3 * i + 1] = k.y;
        keys[3 * i + 2] = k.z;
    }
    return o->ntouched;
}

/* keys int32[nb,3], hashes u64[nb], vox f32[nb,5,nvox] (tsdf, weight, r, g, b). Any may be NULL. */
int64_t tsdf_oracle_dump(const tsdf_oracle *o, int32_t *keys, uint64_t *hashes, float *vox) {
    for (int64_t b = 0; b < o->nb; ++b) {
        if (keys) {
            keys[3 * b + 0] = o->keys[b].x;
            keys[3 * b + 1] = o->keys[b].y;
            keys[3 * b + 2] = o->keys[b].z;
        }
        if (hashes) hashes[b] = tsdf_oracle_block_key_hash(o->keys[b].x, o->keys[b].y, o->keys[b].z);
    }
    if (vox) memcpy(vox, o->vox, sizeof(float) * 5 * (size_t)o->nvox * (size_t)o->nb);
    return o->nb;
}

/* ------------------------------------------------------------------------------------------ */
/* A.4 marching cubes with vertex welding by canonical edge id (voxel + axis).                */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int32_t x, y, z, a;
} edge4;

typedef struct {
    int64_t cap, n;
    edge4 *key;
    int32_t *val;
} edge_map;

static uint64_t edge_hash(edge4 e) {
    key3 k = {e.x, e.y, e.z};
    return mix64(slot_hash(k) + (uint64_t)e.a);
}

static void edge_map_init(edge_map *m, int64_t cap) {
    m->cap = cap;
    m->n = 0;
    m->key = (edge4 *)malloc(sizeof(edge4) * (size_t)cap);
    m->val = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    for (int64_t i = 0; i < cap; ++i) m->val[i] = -1;
}

static void edge_map_grow(edge_map *m) {
    edge_map n;
    edge_map_init(&n, m->cap * 2);
    for (int64_t i = 0; i < m->cap; ++i) {
        if (m->val[i] < 0) continue;
        uint64_t s = edge_hash(m->key[i]) & (uint64_t)(n.cap - 1);
        while (n.val[s] >= 0) s = (s + 1) & (uint64_t)(n.cap - 1);
        n.key[s] = m->key[i];
        n.val[s] = m->val[i];
    }
    n.n = m->n;
    free(m->key);
    free(m->val);
    *m = n;
}

/* returns existing id or -(slot+1) for the empty slot where it would go */
static int64_t edge_map_find(const edge_map *m, edge4 e) {
    uint64_t s = edge_hash(e) & (uint64_t)(m->cap - 1);
    for (;;) {
        if (m->val[s] < 0) return -((int64_t)s + 1);
        const edge4 k = m->key[s];
        if (k.x == e.x && k.y == e.y && k.z == e.z && k.a == e.a) return m->val[s];
        s = (s + 1) & (uint64_t)(m->cap - 1);
    }
}

/* fetch (tsdf, w, r, g, b) of global voxel g; returns 0 if the owning block is missing */
static int voxel_fetch(const tsdf_oracle *o, int32_t gx, int32_t gy, int32_t gz, float out[5]) {
    const int B = o->B;
    key3 k;
    k.x = (int32_t)tsdf_oracle_floor_div(gx, B);
    k.y = (int32_t)tsdf_oracle_floor_div(gy, B);
    k.z = (int32_t)tsdf_oracle_floor_div(gz, B);
    const int64_t b = block_find(o, k);
    if (b < 0) return 0;
    const int lx = gx - k.x * B, ly = gy - k.y * B, lz = gz - k.z * B;
    const int idx = lx + ly * B + lz * B * B;
    const float *vox = o->vox + (size_t)b * 5 * o->nvox;
    for (int c = 0; c < 5; ++c) out[c] = vox[(size_t)c * o->nvox + idx];
    return 1;
}

typedef struct {
    int64_t nv, nt, vcap, tcap;
    double *vert64;  /* [nv,3] Open3D's float64 formula */
    double *color;   /* [nv,3] in [0,1], Open3D's float64 blend of the (float32) voxel colours */
    int32_t *edge;   /* [nv,4] canonical edge id (gx,gy,gz,axis) */
    int32_t *tri;    /* [nt,3] */
} mesh_out;

static void mesh_push_vertex(mesh_out *m, const double p64[3], const double c[3], edge4 e) {
    if (m->nv == m->vcap) {
        m->vcap = m->vcap ? m->vcap * 2 : 4096;
        m->vert64 = (double *)realloc(m->vert64, sizeof(double) * 3 * (size_t)m->vcap);
        m->color = (double *)realloc(m->color, sizeof(double) * 3 * (size_t)m->vcap);
        m->edge = (int32_t *)realloc(m->edge, sizeof(int32_t) * 4 * (size_t)m->vcap);
    }
    for (int k = 0; k < 3; ++k) {
        m->vert64[3 * m->nv + k] = p64[k];
        m->color[3 * m->nv + k] = c[k];
    }
    m->edge[4 * m->nv + 0] = e.x;
    m->edge[4 * m->nv + 1] = e.y;
    m->edge[4 * m->nv + 2] = e.z;
    m->edge[4 * m->nv + 3] = e.a;
    m->nv++;
}

static mesh_out g_mesh; /* result of the last extract, copied out by tsdf_oracle_mesh_copy */

/* Runs A.4 over the whole volume; returns counts via nv/nt. */
void tsdf_oracle_extract_mesh(const tsdf_oracle *o, int64_t *nv, int64_t *nt) {
    free(g_mesh.vert64);
    free(g_mesh.color);
    free(g_mesh.edge);
    free(g_mesh.tri);
    memset(&g_mesh, 0, sizeof(g_mesh));
    edge_map em;
    edge_map_init(&em, 1 << 16);
    const int B = o->B;
    const double vs64 = o->unit_len / (double)(o->unit_blocks * o->B), h64 = vs64 * 0.5; /* voxel_length_ */
    for (int64_t b = 0; b < o->nb; ++b) {
        const key3 bk = o->keys[b];
        for (int lz = 0; lz < B; ++lz)
            for (int ly = 0; ly < B; ++ly)
                for (int lx = 0; lx < B; ++lx) {
                    const int32_t gx = bk.x * B + lx, gy = bk.y * B + ly, gz = bk.z * B + lz;
                    float f[8], col[8][3];
                    int cube = 0, ok = 1;
                    for (int i = 0; i < 8 && ok; ++i) {
                        float v5[5];
                        if (!voxel_fetch(o, gx + MC_SHIFT[i][0], gy + MC_SHIFT[i][1],
                                         gz + MC_SHIFT[i][2], v5) ||
                            v5[1] == 0.0f) {
                            ok = 0;
                            break;
                        }
                        f[i] = v5[0];
                        col[i][0] = v5[2];
                        col[i][1] = v5[3];
                        col[i][2] = v5[4];
                        if (f[i] < 0.0f) cube |= (1 << i);
                    }
                    if (!ok || cube == 0 || cube == 255) continue;
                    int32_t vid[12];
                    for (int e = 0; e < 12; ++e) {
                        vid[e] = -1;
                        if (!(MC_EDGE_TABLE[cube] & (1 << e))) continue;
                        edge4 ek = {gx + MC_EDGE_SHIFT[e][0], gy + MC_EDGE_SHIFT[e][1],
                                    gz + MC_EDGE_SHIFT[e][2], MC_EDGE_SHIFT[e][3]};
                        int64_t r = edge_map_find(&em, ek);
                        if (r >= 0) {
                            vid[e] = (int32_t)r;
                            continue;
                        }
                        const int c0 = MC_EDGE_TO_VERT[e][0], c1 = MC_EDGE_TO_VERT[e][1];
                        const double f0 = fabs((double)f[c0]), f1 = fabs((double)f[c1]);
                        double p64[3] = {h64 + vs64 * ek.x, h64 + vs64 * ek.y, h64 + vs64 * ek.z};
                        p64[ek.a] += f0 * vs64 / (f0 + f1);
                        double cc[3];
                        for (int k = 0; k < 3; ++k)
                            cc[k] = (f1 * ((double)col[c0][k] / 255.0) + f0 * ((double)col[c1][k] / 255.0)) / (f0 + f1);
                        const int64_t s = -r - 1;
                        em.key[s] = ek;
                        em.val[s] = (int32_t)g_mesh.nv;
                        em.n++;
                        vid[e] = (int32_t)g_mesh.nv;
                        mesh_push_vertex(&g_mesh, p64, cc, ek);
                        if (em.n * 2 > em.cap) edge_map_grow(&em);
                    }
                    for (int t = 0; MC_TRI_TABLE[cube][t] != -1 && t < 15; t += 3) {
                        if (g_mesh.nt == g_mesh.tcap) {
                            g_mesh.tcap = g_mesh.tcap ? g_mesh.tcap * 2 : 4096;
                            g_mesh.tri = (int32_t *)realloc(g_mesh.tri,
                                                            sizeof(int32_t) * 3 * (size_t)g_mesh.tcap);
                        }
                        /* winding (i, i+2, i+1) as in A.4 */
                        g_mesh.tri[3 * g_mesh.nt + 0] = vid[(int)MC_TRI_TABLE[cube][t]];
                        g_mesh.tri[3 * g_mesh.nt + 1] = vid[(int)MC_TRI_TABLE[cube][t + 2]];
                        g_mesh.tri[3 * g_mesh.nt + 2] = vid[(int)MC_TRI_TABLE[cube][t + 1]];
                        g_mesh.nt++;
                    }
                }
    }
    free(em.key);
    free(em.val);
    *nv = g_mesh.nv;
    *nt = g_mesh.nt;
}

void tsdf_oracle_mesh_copy(double *vert64, double *color, int32_t *edge, int32_t *tri) {
    if (vert64) memcpy(vert64, g_mesh.vert64, sizeof(double) * 3 * (size_t)g_mesh.nv);
    if (color) memcpy(color, g_mesh.color, sizeof(double) * 3 * (size_t)g_mesh.nv);
    if (edge) memcpy(edge, g_mesh.edge, sizeof(int32_t) * 4 * (size_t)g_mesh.nv);
    if (tri) memcpy(tri, g_mesh.tri, sizeof(int32_t) * 3 * (size_t)g_mesh.nt);
}

int tsdf_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
