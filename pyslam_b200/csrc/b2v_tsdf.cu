// b2v_tsdf.cu — the two per-frame kernels of the TSDF path (sm_100a).
//
//   allocate_kernel   voxel-block hash allocation along each sampled depth ray
//                     (replaces Open3D ScalableTSDFVolume::Integrate's touched-unit loop, called
//                      from pyslam/dense/volumetric_integrator_tsdf.py:223; keys/hash follow
//                      cpp/volumetric/voxel_hashing.h:69-161)
//   integrate_kernel  per-voxel projective TSDF + colour weighted update of every touched block
//                     (replaces Open3D UniformTSDFVolume::IntegrateWithDepthToCameraDistanceMultiplier;
//                      block layout follows cpp/volumetric/voxel_block.h:45-70)
//
// Arithmetic contract: DESIGN.md §"Arithmetic contract".  Every floating-point operation that
// decides a key, a pixel or a stored value is written with an explicit-rounding intrinsic so the
// compiler can neither contract nor reorder it; the CPU oracle performs the same IEEE operations.
#include "b2v_internal.h"

namespace b2v {

// ------------------------------------------------------------------------------------------------
// allocation
// ------------------------------------------------------------------------------------------------

constexpr int kAllocTile = 8;       // 8 x 8 depth samples per CTA ...
constexpr int kAllocSub = 4;        // ... x 4 lanes per sample: a sample's (typically 27) candidate blocks
                                    // are split over 4 lanes, so the serial chain per warp is 7 long
constexpr int kAllocThreads = kAllocTile * kAllocTile * kAllocSub;
constexpr int kSetSize = 512;       // CTA-local de-duplication set (power of two)
constexpr int kListCap = 512;       // CTA-local lists of fresh / first-touched slots

// 21 low bits per axis: injective inside one frame's frustum (checked on the host at create)
__device__ __forceinline__ unsigned long long pack_key(int x, int y, int z) {
    return static_cast<unsigned long long>(x & 0x1FFFFF) |
           (static_cast<unsigned long long>(y & 0x1FFFFF) << 21) |
           (static_cast<unsigned long long>(z & 0x1FFFFF) << 42);
}

__device__ __forceinline__ void assign_block(const HashTable &T, const PoolMeta &M, uint32_t slot,
                                             uint32_t idx) {
    uint32_t *w = reinterpret_cast<uint32_t *>(T.entries + slot) + 3;
    if (idx < M.capacity) {
        const uint4 e = ld_entry(T.entries + slot);
        M.block_keys[idx] = make_int4(static_cast<int>(e.x), static_cast<int>(e.y),
                                      static_cast<int>(e.z), 0);
        *w = idx;
    } else {
        *w = kNoBlock;
        atomicOr(M.counters + kCtrError, 1u);
    }
}

// recover a coordinate from its 21-bit packed field, given any reference within 2^20 of it
__device__ __forceinline__ int unpack_axis(unsigned long long pk, int shift, int ref) {
    const int d = (static_cast<int>((pk >> shift) & 0x1FFFFF) - ref) & 0x1FFFFF;
    return ref + ((d ^ 0x100000) - 0x100000);  // sign-extend 21 bits
}

// Global find-or-insert of one block key + first-touch detection for this frame.  New slots and
// first-touched slots are queued in shared-memory lists (flushed with one atomic per CTA).
__device__ __forceinline__ void touch_key(const FrameParams &P, const HashTable &T, const PoolMeta &M,
                                          int ring, int kx, int ky, int kz, uint32_t *s_new,
                                          uint32_t *s_n_new, uint32_t *s_act, uint32_t *s_n_act) {
    if (P.shard_count > 1 &&
        static_cast<int>(block_key_hash(kx, ky, kz) % static_cast<uint64_t>(P.shard_count)) != P.shard_rank)
        return;
    bool is_new;
    const uint32_t slot = table_insert(T, kx, ky, kz, &is_new);
    if (slot == kEmpty) {
        atomicOr(M.counters + kCtrError, 2u);
        return;
    }
    if (is_new) {
        const uint32_t pos = atomicAdd(s_n_new, 1u);
        if (pos < kListCap) {
            s_new[pos] = slot;
        } else {  // list overflow: assign directly
            assign_block(T, M, slot, atomicAdd(M.counters + kCtrPool, 1u));
            atomicAdd(M.counters + kCtrNew0 + ring, 1u);
        }
    }
    if (atomicExch(T.stamp + slot, P.frame_id) != P.frame_id) {
        const uint32_t pos = atomicAdd(s_n_act, 1u);
        if (pos < kListCap) {
            s_act[pos] = slot;
        } else {
            const uint32_t g = atomicAdd(M.counters + kCtrActive0 + ring, 1u);
            if (g < M.capacity) M.active_slots[static_cast<size_t>(ring) * M.capacity + g] = slot;
        }
    }
}

// Phase A  every sampled pixel enumerates the blocks of its [p - tau, p + tau] box; duplicates die
//          first inside the warp (__match_any_sync ballot), then inside the CTA (shared-memory set).
// Phase B  the CTA's unique keys probe / insert into the global table one key per thread, all in
//          flight at once (the probes are independent L2 round trips), and exchange the frame stamp.
// Phase C  one atomic per CTA hands out contiguous pool indices and active-list positions.
__global__ void __launch_bounds__(kAllocThreads)
allocate_kernel(const FrameParams P, const float *__restrict__ depth, const uint8_t *__restrict__ rgb,
                const float *__restrict__ lam, float4 *__restrict__ tex, const HashTable T,
                const PoolMeta M, const int ring) {
    __shared__ unsigned long long s_set[kSetSize];
    __shared__ uint32_t s_new[kListCap];
    __shared__ uint32_t s_act[kListCap];
    __shared__ uint32_t s_n_new, s_n_act, s_base_new, s_base_act;
    __shared__ int s_ref[4];  // reference key for unpacking; s_ref[3]: 0 = unset, 1 = set

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    for (int i = tid; i < kSetSize; i += kAllocThreads) s_set[i] = ~0ull;
    if (tid == 0) {
        s_n_new = 0;
        s_n_act = 0;
        s_ref[3] = 0;
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            // the ring slot the NEXT frame will count into (its last user finished 3 frames ago)
            const int nxt = (ring + 1) % kActiveRing;
            M.counters[kCtrActive0 + nxt] = 0;
            M.counters[kCtrNew0 + nxt] = 0;
        }
    }
    __syncthreads();

    // ---- pack this CTA's pixel tile into 16-byte texels {valid depth | 0, lambda, rgbx, 0}: the
    //      integrate kernel then needs one gather per voxel instead of four ----
    {
        const int tile = kAllocTile * P.stride;  // pixels per tile side
        const int x0 = blockIdx.x * tile, y0 = blockIdx.y * tile;
        for (int q = tid; q < tile * tile; q += kAllocThreads) {
            const int x = x0 + q % tile, y = y0 + q / tile;
            if (x < P.W && y < P.H) {
                const size_t p = static_cast<size_t>(y) * P.W + x;
                const float d = __ldg(depth + p);
                const uint8_t *c = rgb + 3 * p;
                const uint32_t rgbx = static_cast<uint32_t>(__ldg(c)) | (static_cast<uint32_t>(__ldg(c + 1)) << 8) |
                                      (static_cast<uint32_t>(__ldg(c + 2)) << 16);
                tex[p] = make_float4((d > 0.0f && d < P.depth_trunc) ? d : 0.0f, __ldg(lam + p),
                                     __uint_as_float(rgbx), 0.0f);
            }
        }
    }

    // ---- this thread's depth sample (4 consecutive lanes share one) and its block range ----
    const int sample = tid / kAllocSub, sub = tid % kAllocSub;
    const int j = (blockIdx.x * kAllocTile + (sample & (kAllocTile - 1))) * P.stride;
    const int i = (blockIdx.y * kAllocTile + (sample / kAllocTile)) * P.stride;
    int lo0 = 0, lo1 = 0, lo2 = 0, n1 = 1, n2 = 1, ncand = 0;
    if (j < P.W && i < P.H) {
        const float d = __ldg(depth + static_cast<size_t>(i) * P.W + j);
        if (d > 0.0f && d < P.depth_trunc) {
            const double z = static_cast<double>(d);
            const double x = __ddiv_rn(__dmul_rn(__dsub_rn(static_cast<double>(j), P.cx), z), P.fx);
            const double y = __ddiv_rn(__dmul_rn(__dsub_rn(static_cast<double>(i), P.cy), z), P.fy);
            int lo[3], n[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double pw = __dadd_rn(
                    __dadd_rn(__dadd_rn(__dmul_rn(P.Rwc[3 * a + 0], x), __dmul_rn(P.Rwc[3 * a + 1], y)),
                              __dmul_rn(P.Rwc[3 * a + 2], z)),
                    P.twc[a]);
                const int vlo = voxel_coord(__double2float_rn(__dsub_rn(pw, P.tau_d)), P.inv_vs);
                const int vhi = voxel_coord(__double2float_rn(__dadd_rn(pw, P.tau_d)), P.inv_vs);
                lo[a] = block_coord(vlo);
                n[a] = block_coord(vhi) - lo[a] + 1;
            }
            lo0 = lo[0], lo1 = lo[1], lo2 = lo[2];
            n1 = n[1], n2 = n[2];
            ncand = n[0] * n[1] * n[2];
            if (atomicCAS(&s_ref[3], 0, 1) == 0) {  // one reference key per CTA for unpacking
                s_ref[0] = lo0;
                s_ref[1] = lo1;
                s_ref[2] = lo2;
            }
        }
    }

    // ---- phase A: shared-memory only; lane `sub` of a sample takes candidates sub, sub+4, ... ----
    const int wmax = __reduce_max_sync(0xffffffffu, ncand);
    int dx = 0, dy = 0, dz = sub;  // candidate c = (dx * n1 + dy) * n2 + dz, advanced incrementally
    for (int c = sub; c - sub < wmax; c += kAllocSub, dz += kAllocSub) {
        const bool have = c < ncand;
        int kx = 0, ky = 0, kz = 0;
        unsigned long long pk = (1ull << 63) | static_cast<unsigned long long>(lane);
        if (have) {
            while (dz >= n2) {
                dz -= n2;
                if (++dy >= n1) {
                    dy = 0;
                    ++dx;
                }
            }
            kx = lo0 + dx;
            ky = lo1 + dy;
            kz = lo2 + dz;
            pk = pack_key(kx, ky, kz);
        }
        const unsigned grp = __match_any_sync(0xffffffffu, pk);
        if (have && ((__ffs(grp) - 1) == lane)) {
            uint32_t h = mix32(static_cast<uint32_t>(pk) ^ static_cast<uint32_t>(pk >> 32)) & (kSetSize - 1);
            bool placed = false;
            for (int k = 0; k < 64 && !placed; ++k) {
                const unsigned long long old = atomicCAS(s_set + h, ~0ull, pk);
                placed = (old == ~0ull) || (old == pk);
                h = (h + 1) & (kSetSize - 1);
            }
            // saturated set (> ~400 distinct blocks under one 32x32-pixel tile): go straight to
            // the global table, which de-duplicates anyway
            if (!placed) touch_key(P, T, M, ring, kx, ky, kz, s_new, &s_n_new, s_act, &s_n_act);
        }
    }
    __syncthreads();

    // ---- phase B: every unique key of the CTA, one per thread, all probes in flight ----
    const int rx = s_ref[0], ry = s_ref[1], rz = s_ref[2];
    for (int q = tid; q < kSetSize; q += kAllocThreads) {
        const unsigned long long pk = s_set[q];
        if (pk == ~0ull) continue;
        touch_key(P, T, M, ring, unpack_axis(pk, 0, rx), unpack_axis(pk, 21, ry), unpack_axis(pk, 42, rz),
                  s_new, &s_n_new, s_act, &s_n_act);
    }
    __syncthreads();

    // ---- phase C: one global atomic per CTA for each list; pool indices are contiguous ----
    const uint32_t n_new = min(s_n_new, static_cast<uint32_t>(kListCap));
    const uint32_t n_act = min(s_n_act, static_cast<uint32_t>(kListCap));
    if (tid == 0) {
        s_base_new = n_new ? atomicAdd(M.counters + kCtrPool, n_new) : 0u;
        if (n_new) atomicAdd(M.counters + kCtrNew0 + ring, n_new);
        s_base_act = n_act ? atomicAdd(M.counters + kCtrActive0 + ring, n_act) : 0u;
    }
    __syncthreads();
    for (uint32_t k = tid; k < n_new; k += kAllocThreads) assign_block(T, M, s_new[k], s_base_new + k);
    uint32_t *active_out = M.active_slots + static_cast<size_t>(ring) * M.capacity;
    for (uint32_t k = tid; k < n_act; k += kAllocThreads) {
        const uint32_t g = s_base_act + k;
        if (g < M.capacity) active_out[g] = s_act[k];
    }
}

cudaError_t launch_allocate(const FrameParams &p, const float *depth, const uint8_t *color,
                            const float *lam, float4 *texels, const HashTable &table,
                            const PoolMeta &meta, int ring, cudaStream_t stream) {
    const int gw = (p.W + p.stride - 1) / p.stride;
    const int gh = (p.H + p.stride - 1) / p.stride;
    const dim3 grid((gw + kAllocTile - 1) / kAllocTile, (gh + kAllocTile - 1) / kAllocTile);
    allocate_kernel<<<grid, kAllocThreads, 0, stream>>>(p, depth, color, lam, texels, table, meta, ring);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// projective TSDF + colour update
// ------------------------------------------------------------------------------------------------
//
// One CTA iteration = one touched block: 128 threads x 4 consecutive-x voxels.  A warp reads /
// writes 512 contiguous bytes per plane (LDG.128 / STG.128), so every block-pool transaction is a
// full 128-byte line.  The five plane loads of a block are issued first; the projection of the four
// voxels and their texel gathers (one 16-byte {depth, lambda, rgbx} texel per voxel, packed by
// allocate_kernel and L2-resident) overlap that HBM latency; the next block's table entry is
// prefetched one iteration ahead.  Planes are written back only by threads that updated a voxel.

constexpr int kIntThreads = 128;

__global__ void __launch_bounds__(kIntThreads, 8)
integrate_kernel(const FrameParams P, const float4 *__restrict__ tex, const HashTable T,
                 const PoolMeta M, const int ring) {
    const uint32_t n = min(M.counters[kCtrActive0 + ring], M.capacity);
    const uint32_t *__restrict__ act = M.active_slots + static_cast<size_t>(ring) * M.capacity;
    const int t = threadIdx.x;
    const int lx0 = (t & 1) * 4, ly = (t >> 1) & 7, lz = t >> 4;
    if (blockIdx.x == 0 && t == 0)
        atomicAdd(reinterpret_cast<unsigned long long *>(M.counters + kCtrUpdatesLo),
                  static_cast<unsigned long long>(n));

    uint32_t i = blockIdx.x;
    uint4 e = make_uint4(0u, 0u, 0u, kNoBlock);
    if (i < n) e = T.entries[act[i]];
    while (i < n) {
        const uint32_t i_next = i + gridDim.x;
        uint4 e_next = e;
        if (i_next < n) e_next = T.entries[act[i_next]];  // in flight during this iteration

        if (e.w < M.capacity) {  // (>= capacity: the pool overflowed for this key)
            float *blk = M.pool + static_cast<size_t>(e.w) * kBlockFloats + t * 4;
            float4 q[kPlanes];
#pragma unroll
            for (int c = 0; c < kPlanes; ++c) q[c] = *reinterpret_cast<const float4 *>(blk + c * kVox);

            const int vx0 = static_cast<int>(e.x) * kB + lx0;
            const float cy = __fmul_rn(__fadd_rn(static_cast<float>(static_cast<int>(e.y) * kB + ly), 0.5f), P.vs);
            const float cz = __fmul_rn(__fadd_rn(static_cast<float>(static_cast<int>(e.z) * kB + lz), 0.5f), P.vs);
            // row part of E * c, shared by the voxels of one x-row (contract: fmaf(E0, cx, A))
            const float ax = __fmaf_rn(P.E[1], cy, __fmaf_rn(P.E[2], cz, P.E[3]));
            const float ay = __fmaf_rn(P.E[5], cy, __fmaf_rn(P.E[6], cz, P.E[7]));
            const float az = __fmaf_rn(P.E[9], cy, __fmaf_rn(P.E[10], cz, P.E[11]));

            float pzs[4];
            int pix[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float cx = __fmul_rn(__fadd_rn(static_cast<float>(vx0 + k), 0.5f), P.vs);
                const float px = __fmaf_rn(P.E[0], cx, ax);
                const float py = __fmaf_rn(P.E[4], cx, ay);
                const float pz = __fmaf_rn(P.E[8], cx, az);
                pzs[k] = pz;
                pix[k] = -1;
                if (pz > 0.0f) {
                    const float inv_z = __frcp_rn(pz);
                    const float u_f = __fmaf_rn(__fmul_rn(px, P.fxf), inv_z, P.cxh);
                    const float v_f = __fmaf_rn(__fmul_rn(py, P.fyf), inv_z, P.cyh);
                    if (u_f >= 0.0001f && u_f < P.safe_w && v_f >= 0.0001f && v_f < P.safe_h)
                        pix[k] = __float2int_rz(v_f) * P.W + __float2int_rz(u_f);
                }
            }
            float4 tx[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) tx[k] = pix[k] >= 0 ? __ldg(tex + pix[k]) : make_float4(0.f, 0.f, 0.f, 0.f);

            float *ts = reinterpret_cast<float *>(&q[0]);
            float *w = reinterpret_cast<float *>(&q[1]);
            float *cr = reinterpret_cast<float *>(&q[2]);
            float *cg = reinterpret_cast<float *>(&q[3]);
            float *cb = reinterpret_cast<float *>(&q[4]);
            bool upd = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float d = tx[k].x;  // 0 where the pixel is invalid (allocate_kernel pre-validates)
                const float sdf = __fmul_rn(__fsub_rn(d, pzs[k]), tx[k].y);
                if (d > 0.0f && sdf > -P.tau) {
                    const float tv = fminf(1.0f, __fmul_rn(sdf, P.inv_tau));
                    const uint32_t rgbx = __float_as_uint(tx[k].z);
                    const float w0 = w[k];
                    const float wn = __fadd_rn(w0, 1.0f);
                    const float r = __frcp_rn(wn);
                    ts[k] = __fmul_rn(__fmaf_rn(ts[k], w0, tv), r);
                    cr[k] = __fmul_rn(__fmaf_rn(cr[k], w0, static_cast<float>(rgbx & 0xFFu)), r);
                    cg[k] = __fmul_rn(__fmaf_rn(cg[k], w0, static_cast<float>((rgbx >> 8) & 0xFFu)), r);
                    cb[k] = __fmul_rn(__fmaf_rn(cb[k], w0, static_cast<float>((rgbx >> 16) & 0xFFu)), r);
                    w[k] = wn;
                    upd = true;
                }
            }
            if (upd) {
#pragma unroll
                for (int c = 0; c < kPlanes; ++c) *reinterpret_cast<float4 *>(blk + c * kVox) = q[c];
            }
        }
        e = e_next;
        i = i_next;
    }
}

cudaError_t launch_integrate(const FrameParams &p, const float4 *texels, const HashTable &table,
                             const PoolMeta &meta, int ring, int grid_ctas, cudaStream_t stream) {
    integrate_kernel<<<grid_ctas, kIntThreads, 0, stream>>>(p, texels, table, meta, ring);
    return cudaGetLastError();
}

int integrate_max_resident_ctas_per_sm() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, integrate_kernel, kIntThreads, 0) != cudaSuccess)
        return 8;
    return n > 0 ? n : 1;
}

// lambda(u, v) = sqrt(((u - cx)/fx)^2 + ((v - cy)/fy)^2 + 1): Open3D's depth-to-camera-distance
// multiplier image, recomputed only when the intrinsics or the image size change
__global__ void lambda_kernel(const FrameParams P, float *__restrict__ lam) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= P.W) return;
    const float xx = __fmul_rn(__fsub_rn(static_cast<float>(u), P.cxf), P.inv_fx);
    const float yy = __fmul_rn(__fsub_rn(static_cast<float>(v), P.cyf), P.inv_fy);
    lam[static_cast<size_t>(v) * P.W + u] = __fsqrt_rn(__fmaf_rn(xx, xx, __fmaf_rn(yy, yy, 1.0f)));
}

cudaError_t launch_lambda(const FrameParams &p, float *lam, cudaStream_t stream) {
    lambda_kernel<<<dim3((p.W + 127) / 128, p.H), 128, 0, stream>>>(p, lam);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// small helpers for the parity hooks
// ------------------------------------------------------------------------------------------------

__global__ void block_hashes_kernel(const int4 *__restrict__ keys, uint64_t *__restrict__ hashes,
                                    uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hashes[i] = block_key_hash(keys[i].x, keys[i].y, keys[i].z);
}

cudaError_t launch_block_hashes(const int4 *block_keys, uint64_t *hashes, uint32_t n,
                                cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    block_hashes_kernel<<<(n + 255) / 256, 256, 0, stream>>>(block_keys, hashes, n);
    return cudaGetLastError();
}

__global__ void gather_active_keys_kernel(const HashTable T, const uint32_t *__restrict__ act,
                                          uint32_t n, int4 *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint4 e = T.entries[act[i]];
        out[i] = make_int4(static_cast<int>(e.x), static_cast<int>(e.y), static_cast<int>(e.z),
                           static_cast<int>(e.w));
    }
}

cudaError_t launch_gather_active_keys(const HashTable &table, const uint32_t *active_slots,
                                      uint32_t n, int4 *out, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    gather_active_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(table, active_slots, n, out);
    return cudaGetLastError();
}

// ---- upload (restore / seed) -------------------------------------------------------------------

__global__ void upload_insert_kernel(const int4 *__restrict__ keys, uint32_t n, const HashTable T,
                                     const PoolMeta M, uint32_t *__restrict__ out_idx) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool is_new;
    const uint32_t slot = table_insert(T, keys[i].x, keys[i].y, keys[i].z, &is_new);
    uint32_t idx = kNoBlock;
    if (slot == kEmpty) {
        atomicOr(M.counters + kCtrError, 2u);
    } else if (is_new) {
        idx = atomicAdd(M.counters + kCtrPool, 1u);
        assign_block(T, M, slot, idx);
    } else {
        idx = ld_entry(T.entries + slot).w;
    }
    out_idx[i] = idx;
}

__global__ void __launch_bounds__(128)
upload_copy_kernel(const float *__restrict__ vox, const uint32_t *__restrict__ idx, const PoolMeta M) {
    const uint32_t b = blockIdx.x;
    const uint32_t dst = idx[b];
    if (dst >= M.capacity) return;
    const float4 *src = reinterpret_cast<const float4 *>(vox + static_cast<size_t>(b) * kBlockFloats);
    float4 *out = reinterpret_cast<float4 *>(M.pool + static_cast<size_t>(dst) * kBlockFloats);
    for (int k = threadIdx.x; k < kBlockFloats / 4; k += 128) out[k] = src[k];
}

cudaError_t launch_upload_blocks(const int4 *keys, const float *vox, uint32_t n, uint32_t *scratch_idx,
                                 const HashTable &table, const PoolMeta &meta, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    upload_insert_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keys, n, table, meta, scratch_idx);
    upload_copy_kernel<<<n, 128, 0, stream>>>(vox, scratch_idx, meta);
    return cudaGetLastError();
}

}  // namespace b2v
