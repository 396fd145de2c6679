// b2v_mesh.cu — per-block marching-cubes triangle emission (sm_100a).
//
// Replaces Open3D ScalableTSDFVolume::ExtractTriangleMesh / ExtractPointCloud, called from
// pyslam/dense/volumetric_integrator_tsdf.py:239,246,260,267.  Semantics:
//   - a cube is rooted at every voxel; its 8 corners may live in up to 7 neighbouring blocks;
//     a cube with any zero-weight (or missing) corner is skipped; cases 0 / 255 emit nothing
//   - a vertex lives on an edge identified by (lower-corner voxel, axis) and is shared by every
//     cube around that edge (welding): position = voxel centre + |f0| vs / (|f0| + |f1|) along axis
//   - triangles come from the classic 256-case table with winding (i, i+2, i+1)
// Deterministic three-pass structure: classify -> exclusive scan -> emit.  Output order is
// (pool block, voxel index, axis) for vertices and (pool block, voxel index, case order) for
// triangles.
#include "b2v_internal.h"
#include "b2v_scan.cuh"
#include "mc_tables.h"

namespace b2v {

__constant__ unsigned short c_edge_table[256];
__constant__ signed char c_tri_table[256][16];
__constant__ unsigned char c_num_tris[256];
__constant__ signed char c_edge_shift[12][4];
__constant__ unsigned short c_halo[217];   // the 9^3 - 8^3 tile cells outside the own block: x | y << 4 | z << 8

static cudaError_t upload_tables_once() {
    static bool done = false;
    static int done_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (done && done_device == dev) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_edge_table, MC_EDGE_TABLE, sizeof(MC_EDGE_TABLE))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_tri_table, MC_TRI_TABLE, sizeof(MC_TRI_TABLE))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_num_tris, MC_NUM_TRIS, sizeof(MC_NUM_TRIS))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_edge_shift, MC_EDGE_SHIFT, sizeof(MC_EDGE_SHIFT))) != cudaSuccess) return e;
    unsigned short halo[217];
    int nh = 0;
    for (int z = 0; z < 9; ++z)
        for (int y = 0; y < 9; ++y)
            for (int x = 0; x < 9; ++x)
                if (x == 8 || y == 8 || z == 8) halo[nh++] = static_cast<unsigned short>(x | (y << 4) | (z << 8));
    if ((e = cudaMemcpyToSymbol(c_halo, halo, sizeof(halo))) != cudaSuccess) return e;
    done = true;
    done_device = dev;
    return cudaSuccess;
}

// ---- pass 0: neighbour block indices --------------------------------------------------------

// pool index of the block at +(o&1, o>>1&1, o>>2&1) of block b, -1 if it does not exist
__device__ __forceinline__ int32_t neighbor_block(const HashTable &T, const PoolMeta &M, uint32_t b, uint32_t o) {
    if (o == 0) return static_cast<int32_t>(b);
    const int4 k = M.block_keys[b];
    const uint32_t s = table_find(T, k.x + (o & 1), k.y + ((o >> 1) & 1), k.z + ((o >> 2) & 1));
    if (s == kEmpty) return -1;
    const uint32_t w = T.entries[s].w;
    return w < M.capacity ? static_cast<int32_t>(w) : -1;
}

// pass 0: one thread per (block, neighbour) - a separate, massively parallel launch so that the two dependent memory
// hops of a table probe are not on the critical path of every classify CTA
__global__ void mesh_neighbors_kernel(const HashTable T, const PoolMeta M, const MeshBuffers mb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < mb.n_blocks * 8u) mb.nbr[i] = neighbor_block(T, M, i >> 3, i & 7u);
}

// owner voxel of cube edge e rooted at local (lx,ly,lz): returns flat index into per-voxel arrays
__device__ __forceinline__ bool edge_owner(const int *s_nbr, int lx, int ly, int lz, int e,
                                           size_t *flat, int *axis) {
    const int ox = lx + c_edge_shift[e][0], oy = ly + c_edge_shift[e][1], oz = lz + c_edge_shift[e][2];
    *axis = c_edge_shift[e][3];
    const int nb = (ox >> 3) | ((oy >> 3) << 1) | ((oz >> 3) << 2);
    const int ob = s_nbr[nb];
    if (ob < 0) return false;
    *flat = static_cast<size_t>(ob) * kVox + ((ox & 7) + ((oy & 7) << 3) + ((oz & 7) << 6));
    return true;
}

// ---- pass 1a: marching-cubes case + vertex ownership (mesh) ---------------------------------

constexpr int kClsThreads = 128;   // 4 voxels per thread: 16 resident CTAs per SM keep more tile loads in flight

__global__ void __launch_bounds__(kClsThreads)
mesh_classify_kernel(const PoolMeta M, const MeshBuffers mb) {
    __shared__ float s_f[729];
    __shared__ float s_w[729];
    __shared__ int s_nbr[8];
    const uint32_t b = blockIdx.x;
    const int t = threadIdx.x;
    if (t < 8) s_nbr[t] = mb.nbr[b * 8 + t];
    // own voxels first (coalesced, independent of the neighbours): voxel t + 128 k
    const float *own = M.pool + static_cast<size_t>(b) * kBlockFloats;
    float f0[4], w0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        f0[k] = own[t + kClsThreads * k];
        w0[k] = own[kVox + t + kClsThreads * k];
    }
    const int lx = t & 7, ly = (t >> 3) & 7, lz0 = t >> 6;   // voxel k: lz = lz0 + 2 k
    bool neg = false, pos = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = lx + ly * 9 + (lz0 + 2 * k) * 81;
        s_f[c] = f0[k];
        s_w[c] = w0[k];
        neg |= w0[k] != 0.0f && f0[k] < 0.0f;
        pos |= w0[k] != 0.0f && !(f0[k] < 0.0f);
    }
    __syncthreads();
    // the 217 halo cells of the 9^3 tile
    for (int i = t; i < 217; i += kClsThreads) {
        const int h = c_halo[i];
        const int x = h & 15, y = (h >> 4) & 15, z = h >> 8;
        const int pb = s_nbr[(x >> 3) | ((y >> 3) << 1) | ((z >> 3) << 2)];
        float f = 0.0f, w = 0.0f;
        if (pb >= 0) {
            const float *blk = M.pool + static_cast<size_t>(pb) * kBlockFloats;
            const int v = (x & 7) + ((y & 7) << 3) + ((z & 7) << 6);
            f = blk[v];
            w = blk[kVox + v];
        }
        s_f[x + y * 9 + z * 81] = f;
        s_w[x + y * 9 + z * 81] = w;
        neg |= w != 0.0f && f < 0.0f;
        pos |= w != 0.0f && !(f < 0.0f);
    }
    // a tile whose observed voxels all lie on one side of the surface has no cube to emit (most blocks: free space or
    // behind the surface); mb.cube / mb.edge_mask were cleared by the launcher
    const int has_neg = __syncthreads_or(neg);
    const int has_pos = __syncthreads_or(pos);
    if (!(has_neg && has_pos)) return;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int lz = lz0 + 2 * k;
        int cube = 0;
        bool ok = true;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            // corner offsets {000,100,110,010,001,101,111,011} (SURVEY.md A.4 `shift`)
            const int sx = ((c + 1) >> 1) & 1, sy = (c >> 1) & 1, sz = c >> 2;
            const int i = (lx + sx) + (ly + sy) * 9 + (lz + sz) * 81;
            ok = ok && (s_w[i] != 0.0f);
            if (s_f[i] < 0.0f) cube |= (1 << c);
        }
        if (!ok || cube == 255) cube = 0;
        if (cube) {
            mb.cube[static_cast<size_t>(b) * kVox + t + kClsThreads * k] = static_cast<uint8_t>(cube);
            const unsigned em = c_edge_table[cube];
            for (int e = 0; e < 12; ++e) {
                if (!((em >> e) & 1u)) continue;
                size_t flat;
                int axis;
                if (edge_owner(s_nbr, lx, ly, lz, e, &flat, &axis))
                    atomicOr(mb.edge_mask + (flat >> 2), (1u << axis) << ((flat & 3) * 8));
            }
        }
    }
}

// ---- pass 1b: zero-crossing masks (point cloud) ---------------------------------------------

__global__ void __launch_bounds__(kVox)
point_masks_kernel(const PoolMeta M, const MeshBuffers mb) {
    __shared__ int s_nbr[8];
    __shared__ uint8_t s_m[kVox];
    const uint32_t b = blockIdx.x;
    const int t = threadIdx.x;
    if (t < 8) s_nbr[t] = mb.nbr[b * 8 + t];
    __syncthreads();
    const float *blk = M.pool + static_cast<size_t>(b) * kBlockFloats;
    const float f0 = blk[t], w0 = blk[kVox + t];
    const int l[3] = {t & 7, (t >> 3) & 7, t >> 6};
    unsigned m = 0;
    if (w0 != 0.0f && f0 < 0.98f && f0 >= -0.98f) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int q[3] = {l[0], l[1], l[2]};
            q[a] += 1;
            const int pb = s_nbr[(q[0] >> 3) | ((q[1] >> 3) << 1) | ((q[2] >> 3) << 2)];
            if (pb < 0) continue;
            const float *nb = M.pool + static_cast<size_t>(pb) * kBlockFloats;
            const int v = (q[0] & 7) + ((q[1] & 7) << 3) + ((q[2] & 7) << 6);
            const float f1 = nb[v], w1 = nb[kVox + v];
            if (w1 != 0.0f && f1 < 0.98f && f1 >= -0.98f && f0 * f1 < 0.0f) m |= 1u << a;
        }
    }
    s_m[t] = static_cast<uint8_t>(m);
    mb.cube[static_cast<size_t>(b) * kVox + t] = 0;
    __syncthreads();
    if (t < kVox / 4)
        mb.edge_mask[static_cast<size_t>(b) * (kVox / 4) + t] = reinterpret_cast<const uint32_t *>(s_m)[t];
}

// ---- pass 2: per-block sums and their exclusive scans ---------------------------------------

__global__ void __launch_bounds__(128)
mesh_block_sums_kernel(const MeshBuffers mb) {
    __shared__ uint32_t s_v[4], s_t[4];
    const uint32_t b = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t m4 = mb.edge_mask[static_cast<size_t>(b) * 128 + t];
    const uint32_t c4 = reinterpret_cast<const uint32_t *>(mb.cube)[static_cast<size_t>(b) * 128 + t];
    uint32_t pv[4], pt[4];   // vertices / triangles of the thread's four voxels
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        pv[k] = __popc((m4 >> (8 * k)) & 7u);
        pt[k] = c_num_tris[(c4 >> (8 * k)) & 0xFFu];
    }
    const uint32_t nv = pv[0] + pv[1] + pv[2] + pv[3], nt = pt[0] + pt[1] + pt[2] + pt[3];
    // exclusive scan over the 128 threads (vertices in the low half, triangles in the high half: <= 1536 / 2560)
    uint32_t x = nv | (nt << 16);
    const uint32_t mine = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) {
        s_v[wid] = x & 0xFFFFu;
        s_t[wid] = x >> 16;
    }
    __syncthreads();
    uint32_t ov = 0, ot = 0;
    for (int w = 0; w < wid; ++w) {
        ov += s_v[w];
        ot += s_t[w];
    }
    const uint32_t sv = s_v[0] + s_v[1] + s_v[2] + s_v[3], st = s_t[0] + s_t[1] + s_t[2] + s_t[3];
    if (sv | st) {  // per-voxel position inside the block: the emit kernels need no scan of their own
        uint32_t bv = ov + ((x - mine) & 0xFFFFu), bt = ot + ((x - mine) >> 16);
        uint4 out;
        uint32_t *o = &out.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[k] = bv | (bt << 16);
            bv += pv[k];
            bt += pt[k];
        }
        reinterpret_cast<uint4 *>(mb.local)[static_cast<size_t>(b) * 128 + t] = out;
    }
    if (t == 0) {
        mb.sums[b] = sv;
        mb.sums[mb.n_blocks + b] = st;
        // the emit kernels run over the blocks that have output only (list order does not matter: positions come
        // from the scan of the sums)
        if (sv) mb.work[atomicAdd(mb.totals + 2, 1u)] = b;
        if (st) mb.work[mb.n_blocks + atomicAdd(mb.totals + 3, 1u)] = b;
    }
}

// ---- pass 3a: vertices -----------------------------------------------------------------------

// points = false: ScalableTSDFVolume::ExtractTriangleMesh vertex / colour formulas (float64):
//     pt = vl/2 + vl * e;  pt[a] += |f0| vl / (|f0| + |f1|);  colour (|f1| c0/255 + |f0| c1/255) / (|f0| + |f1|)
// points = true: ExtractPointCloud: p0 = (vl/2 + vl * x_in_unit) + unit * L, p1 = p0 + vl on the axis,
//     p = (p0 r1 + p1 r0) / (r0 + r1) with float32 r0 = |f0|, r1 = |f1| (their sum in float32), colour
//     ((c0 r1 + c1 r0) / (r0 + r1)) / 255 in float32, widened
constexpr int kEmitThreads = 128;   // four CTAs per block, no barrier: threads without output leave at once

template <bool kPoints>
__global__ void __launch_bounds__(kEmitThreads)
mesh_vertices_kernel(const PoolMeta M, const MeshBuffers mb, const double vl, const int unit_shift) {
    const uint32_t b = mb.work[blockIdx.x >> 2];  // blocks with at least one vertex
    const int t = (blockIdx.x & 3) * kEmitThreads + threadIdx.x;
    const size_t flat = static_cast<size_t>(b) * kVox + t;
    const unsigned m = (mb.edge_mask[flat >> 2] >> ((flat & 3) * 8)) & 7u;
    if (m == 0) return;
    const int *s_nbr = mb.nbr + static_cast<size_t>(b) * 8;
    const float *blk = M.pool + static_cast<size_t>(b) * kBlockFloats;
    const float r0 = fabsf(blk[t]);
    const float c0[3] = {blk[2 * kVox + t], blk[3 * kVox + t], blk[4 * kVox + t]};
    const int4 key = M.block_keys[b];
    const uint32_t base = mb.offs[b] + (mb.local[flat] & 0xFFFFu);
    const int l[3] = {t & 7, (t >> 3) & 7, t >> 6};
    const int g[3] = {key.x * kB + l[0], key.y * kB + l[1], key.z * kB + l[2]};
    const double half = __dmul_rn(vl, 0.5);
    double ctr[3];
    if (kPoints) {
        const int kk[3] = {key.x, key.y, key.z};
        const double L = __dmul_rn(vl, static_cast<double>(kB << unit_shift));  // volume_unit_length_
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int u = kk[a] >> unit_shift;
            const int x = (kk[a] - (u << unit_shift)) * kB + l[a];
            ctr[a] = __dadd_rn(__dadd_rn(half, __dmul_rn(vl, static_cast<double>(x))), __dmul_rn(static_cast<double>(u), L));
        }
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) ctr[a] = __dadd_rn(half, __dmul_rn(vl, static_cast<double>(g[a])));
    }
    uint32_t vid = base;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (!((m >> a) & 1u)) continue;
        int q[3] = {l[0], l[1], l[2]};
        q[a] += 1;
        const int pb = s_nbr[(q[0] >> 3) | ((q[1] >> 3) << 1) | ((q[2] >> 3) << 2)];
        const float *nb = M.pool + static_cast<size_t>(pb) * kBlockFloats;
        const int v = (q[0] & 7) + ((q[1] & 7) << 3) + ((q[2] & 7) << 6);
        const float r1 = fabsf(nb[v]);
        double p[3] = {ctr[0], ctr[1], ctr[2]};
        double col[3];
        if (kPoints) {
            const float rs = __fadd_rn(r0, r1);
            const double p1 = __dadd_rn(ctr[a], vl);
            p[a] = __ddiv_rn(__dadd_rn(__dmul_rn(ctr[a], static_cast<double>(r1)), __dmul_rn(p1, static_cast<double>(r0))),
                             static_cast<double>(rs));
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float c1 = nb[(2 + k) * kVox + v];
                const float num = __fadd_rn(__fmul_rn(c0[k], r1), __fmul_rn(c1, r0));
                col[k] = static_cast<double>(__fdiv_rn(__fdiv_rn(num, rs), 255.0f));
            }
        } else {
            const double f0 = static_cast<double>(r0), f1 = static_cast<double>(r1);
            const double fs = __dadd_rn(f0, f1);
            p[a] = __dadd_rn(p[a], __ddiv_rn(__dmul_rn(f0, vl), fs));
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double d0 = __ddiv_rn(static_cast<double>(c0[k]), 255.0);
                const double d1 = __ddiv_rn(static_cast<double>(nb[(2 + k) * kVox + v]), 255.0);
                col[k] = __ddiv_rn(__dadd_rn(__dmul_rn(f1, d0), __dmul_rn(f0, d1)), fs);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mb.vertices[3 * static_cast<size_t>(vid) + k] = p[k];
            mb.colors[3 * static_cast<size_t>(vid) + k] = col[k];
        }
        reinterpret_cast<int4 *>(mb.edge_ids)[vid] = make_int4(g[0], g[1], g[2], a);
        ++vid;
    }
}

// ---- pass 3b: triangles ----------------------------------------------------------------------

__global__ void __launch_bounds__(kEmitThreads)
mesh_triangles_kernel(const MeshBuffers mb) {
    const uint32_t b = mb.work[mb.n_blocks + (blockIdx.x >> 2)];  // blocks with at least one triangle
    const int t = (blockIdx.x & 3) * kEmitThreads + threadIdx.x;
    const size_t own = static_cast<size_t>(b) * kVox + t;
    const int cube = mb.cube[own];
    const uint32_t nt = c_num_tris[cube];
    if (nt == 0) return;
    const uint32_t tbase = mb.offs[mb.n_blocks + b] + (mb.local[own] >> 16);
    const int *s_nbr = mb.nbr + static_cast<size_t>(b) * 8;
    const int lx = t & 7, ly = (t >> 3) & 7, lz = t >> 6;
    int vid[12];
    const unsigned em = c_edge_table[cube];
    for (int e = 0; e < 12; ++e) {
        vid[e] = -1;
        if (!((em >> e) & 1u)) continue;
        size_t flat;
        int axis;
        if (!edge_owner(s_nbr, lx, ly, lz, e, &flat, &axis)) continue;
        const unsigned m = (mb.edge_mask[flat >> 2] >> ((flat & 3) * 8)) & 7u;
        vid[e] = static_cast<int>(mb.offs[flat >> 9] + (mb.local[flat] & 0xFFFFu) + __popc(m & ((1u << axis) - 1u)));
    }
    for (uint32_t k = 0; k < nt; ++k) {
        int32_t *tri = mb.triangles + 3 * static_cast<size_t>(tbase + k);
        tri[0] = vid[c_tri_table[cube][3 * k + 0]];
        tri[1] = vid[c_tri_table[cube][3 * k + 2]];  // winding (i, i+2, i+1)
        tri[2] = vid[c_tri_table[cube][3 * k + 1]];
    }
}

// ---- launchers -------------------------------------------------------------------------------

cudaError_t launch_mesh_classify(const HashTable &table, const PoolMeta &meta, const MeshBuffers &mb,
                                 cudaStream_t stream) {
    cudaError_t e = upload_tables_once();
    if (e != cudaSuccess || mb.n_blocks == 0) return e;
    e = cudaMemsetAsync(mb.edge_mask, 0, static_cast<size_t>(mb.n_blocks) * kVox, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(mb.cube, 0, static_cast<size_t>(mb.n_blocks) * kVox, stream);
    if (e != cudaSuccess) return e;
    mesh_neighbors_kernel<<<(mb.n_blocks * 8u + 255u) / 256u, 256, 0, stream>>>(table, meta, mb);
    mesh_classify_kernel<<<mb.n_blocks, kClsThreads, 0, stream>>>(meta, mb);
    return cudaGetLastError();
}

cudaError_t launch_point_masks(const HashTable &table, const PoolMeta &meta, const MeshBuffers &mb,
                               cudaStream_t stream) {
    cudaError_t e = upload_tables_once();
    if (e != cudaSuccess || mb.n_blocks == 0) return e;
    mesh_neighbors_kernel<<<(mb.n_blocks * 8u + 255u) / 256u, 256, 0, stream>>>(table, meta, mb);
    point_masks_kernel<<<mb.n_blocks, kVox, 0, stream>>>(meta, mb);
    return cudaGetLastError();
}

cudaError_t launch_mesh_scan(const MeshBuffers &mb, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(mb.totals, 0, 4 * sizeof(uint32_t), stream);
    if (e != cudaSuccess || mb.n_blocks == 0) return e;
    mesh_block_sums_kernel<<<mb.n_blocks, 128, 0, stream>>>(mb);
    exclusive_scan_kernel<<<2, 1024, 0, stream>>>(mb.sums, mb.offs, mb.totals, mb.n_blocks);
    return cudaGetLastError();
}

cudaError_t launch_mesh_vertices(const PoolMeta &meta, const MeshBuffers &mb, double voxel_length, int unit_shift,
                                 bool points, uint32_t work_blocks, cudaStream_t stream) {
    if (work_blocks == 0) return cudaSuccess;
    if (points)
        mesh_vertices_kernel<true><<<work_blocks * 4u, kEmitThreads, 0, stream>>>(meta, mb, voxel_length, unit_shift);
    else
        mesh_vertices_kernel<false><<<work_blocks * 4u, kEmitThreads, 0, stream>>>(meta, mb, voxel_length, unit_shift);
    return cudaGetLastError();
}

cudaError_t launch_mesh_triangles(const MeshBuffers &mb, uint32_t work_blocks, cudaStream_t stream) {
    if (work_blocks == 0) return cudaSuccess;
    mesh_triangles_kernel<<<work_blocks * 4u, kEmitThreads, 0, stream>>>(mb);
    return cudaGetLastError();
}

}  // namespace b2v
