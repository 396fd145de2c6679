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
#include <cstdlib>

#include <cuda_fp16.h>

#include "b2v_internal.h"

namespace b2v {

// ------------------------------------------------------------------------------------------------
// allocation
// ------------------------------------------------------------------------------------------------

constexpr int kAllocTile = 8;       // 8 x 8 depth samples per CTA (32 x 32 pixels at stride 4)
constexpr int kAllocThreads = 256;
constexpr int kBoxSet = 128;        // distinct [lo, lo + n) boxes under one tile (power of two)
constexpr int kBoxList = 64;        // ... compacted
constexpr int kKeySet = 1024;       // distinct block keys under one tile (power of two)
constexpr int kListCap = 512;       // CTA-local lists of fresh / first-touched slots
constexpr uint32_t kNoKey = 0xFFFFFFFFu;

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

// Global find-or-insert of one block key + first-touch detection for this frame.  New slots and
// first-touched slots are queued in shared-memory lists (flushed with one atomic per CTA).
struct FrameSlot {   // which frame this CTA works for
    int group_bit;       // >= 0: fused group mode (bit of the membership mask); -1: per-frame mode
    uint32_t frame_id;
};

// owner rank of a block: BlockKeyHash % N.  64-bit division is emulated (~60 instructions); the
// allocate kernels test ~1000 candidate keys per tile, so a power-of-two rank count takes the mask instead
__device__ __forceinline__ bool owned_by_this_rank(const FrameParams &P, int kx, int ky, int kz) {
    const uint64_t h = block_key_hash(kx, ky, kz);
    const uint32_t n = static_cast<uint32_t>(P.shard_count);
    const uint32_t owner = (n & (n - 1u)) == 0u ? static_cast<uint32_t>(h) & (n - 1u)
                                                : static_cast<uint32_t>(h % static_cast<uint64_t>(n));
    return owner == static_cast<uint32_t>(P.shard_rank);
}

__device__ __forceinline__ void touch_key(const FrameParams &P, const FrameSlot &FS, const HashTable &T, const PoolMeta &M,
                                          int ring, int kx, int ky, int kz, uint32_t *s_new,
                                          uint32_t *s_n_new, uint32_t *s_act, uint32_t *s_n_act) {
    if (P.shard_count > 1 && !owned_by_this_rank(P, kx, ky, kz)) return;
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
            atomicAdd(FS.group_bit >= 0 ? M.counters + group_ctr(P.group_buf, kGcNew) : M.counters + kCtrNew0 + ring, 1u);
        }
    }
    bool first;
    if (FS.group_bit >= 0) {  // fused group mode: membership bit; the first frame to touch queues the slot
        uint32_t *mask = M.group_mask + static_cast<size_t>(P.group_buf) * (static_cast<size_t>(T.mask) + 1);
        first = atomicOr(mask + slot, 1u << FS.group_bit) == 0u;
    } else {
        first = atomicExch(T.stamp + slot, FS.frame_id) != FS.frame_id;
    }
    if (first) {
        const uint32_t pos = atomicAdd(s_n_act, 1u);
        if (pos < kListCap) {
            s_act[pos] = slot;
        } else if (FS.group_bit >= 0) {
            const uint32_t g = atomicAdd(M.counters + group_ctr(P.group_buf, kGcUnion), 1u);
            if (g < M.capacity) M.union_slots[static_cast<size_t>(P.group_buf) * M.capacity + g] = slot;
        } else {
            const uint32_t g = atomicAdd(M.counters + kCtrActive0 + ring, 1u);
            if (g < M.capacity) M.active_slots[static_cast<size_t>(ring) * M.capacity + g] = slot;
        }
    }
}

// block key -> 30-bit code relative to the tile's reference key (10 bits per axis); kNoKey if the
// key is further than 511 blocks from the reference on some axis (then it takes the direct path)
__device__ __forceinline__ uint32_t rel_key(int kx, int ky, int kz, const int *ref) {
    const uint32_t rx = static_cast<uint32_t>(kx - ref[0] + 512), ry = static_cast<uint32_t>(ky - ref[1] + 512),
                   rz = static_cast<uint32_t>(kz - ref[2] + 512);
    if ((rx | ry | rz) >= 1024u) return kNoKey;
    return rx | (ry << 10) | (rz << 20);
}

// every 8^3 block of one allocation unit (Open3D volume unit = 2^3 blocks; decision D1: the unit is the block)
__device__ __forceinline__ void touch_unit(const FrameParams &P, const FrameSlot &FS, const HashTable &T, const PoolMeta &M,
                                           int ring, int ux, int uy, int uz, uint32_t *s_new, uint32_t *s_n_new,
                                           uint32_t *s_act, uint32_t *s_n_act) {
    const int S = P.unit_shift, side = (1 << S) - 1;
    for (int sub = 0; sub < (1 << (3 * S)); ++sub)
        touch_key(P, FS, T, M, ring, (ux << S) + (sub & side), (uy << S) + ((sub >> S) & side), (uz << S) + (sub >> (2 * S)),
                  s_new, s_n_new, s_act, s_n_act);
}

// 16-byte texel of the update kernels: {valid depth | 0, lambda, half2(r, g), half2(b, 0)}; the colours are exact in
// binary16 (integers 0..255) and widen to float32 with one instruction each
__device__ __forceinline__ float4 make_texel(float d, float lam, uint8_t r, uint8_t g, uint8_t b) {
    const __half2 rg = __halves2half2(__ushort2half_rn(r), __ushort2half_rn(g));
    const __half2 bx = __halves2half2(__ushort2half_rn(b), __ushort2half_rn(0));
    return make_float4(d, lam, *reinterpret_cast<const float *>(&rg), *reinterpret_cast<const float *>(&bx));
}

// ---- TMA / mbarrier primitives (sm_90+ PTX; SASS: UTMALDG, SYNCS) ----
constexpr int kTmaTile = 32;  // = kAllocTile * 4: the TMA path serves the default stride 4

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 2-D tiled TMA load: global (tensor map, {c0, c1}) -> shared, completion counted on an mbarrier
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1,
                                            unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

// Per frame: pack the frame into texels, find the touched blocks, allocate the new ones.
//   pack    every CTA packs its 32x32-pixel tile into 16-byte {valid depth | 0, lambda, rgbx} texels
//   boxes   one thread per depth sample: back-project (float64), range [lo, lo+n) of allocation UNITS (Open3D volume
//           units of 2^3 blocks; decision D1: blocks) of the [p - tau, p + tau] box; neighbouring samples share
//           boxes, so the DISTINCT boxes of the tile (typically 10-20) are collected in a shared-memory set
//   keys    the distinct boxes are expanded, one candidate unit per thread, into a shared-memory set of distinct
//           unit keys (typically ~30 units = ~240 blocks per tile)
//   probe   every block of every distinct unit probes / inserts into the global table - one block per thread, all
//           probes in flight - and exchanges the slot's frame stamp (first toucher queues the slot); blocks of
//           another rank are dropped here
//   flush   one atomic per CTA hands out contiguous pool indices and active-list positions
template <bool kTma>
__device__ __forceinline__ void allocate_body(const FrameParams &P, const FramePose &pose, const FrameSlot FS,
                                              const float *__restrict__ depth,
                                              const uint8_t *__restrict__ rgb, const float *__restrict__ lam,
                                              float4 *__restrict__ tex, const HashTable &T, const PoolMeta &M,
                                              const int ring, const FrameMaps &maps, const LambdaMap &lmap) {
    // TMA staging buffers of the 32x32-pixel tile (kTma only): depth, lambda (f32) and colour (u8 x3)
    __shared__ alignas(128) float s_td[kTmaTile * kTmaTile];
    __shared__ alignas(128) float s_tl[kTmaTile * kTmaTile];
    __shared__ alignas(128) uint8_t s_tc[kTmaTile * kTmaTile * 3];
    __shared__ alignas(8) unsigned long long s_bar;
    __shared__ unsigned long long s_boxset[kBoxSet];
    __shared__ unsigned long long s_box[kBoxList];
    __shared__ uint32_t s_keyset[kKeySet];
    __shared__ uint32_t s_keys[kListCap];  // the distinct keys, compacted
    __shared__ uint32_t s_new[kListCap];
    __shared__ uint32_t s_act[kListCap];
    __shared__ uint32_t s_n_box, s_n_keys, s_n_new, s_n_act, s_base_new, s_base_act;
    __shared__ int s_ref[4];  // reference key of the tile; s_ref[3]: 0 = unset, 1 = set
    __shared__ uint32_t s_magic[16];   // ceil(2^16 / d): floor(x / d) = (x * magic) >> 16 for x < 4096, d <= 15

    const int tid = threadIdx.x;
    if (tid < 16) s_magic[tid] = tid ? (65536u + tid - 1u) / static_cast<uint32_t>(tid) : 0u;
    for (int i = tid; i < kKeySet; i += kAllocThreads) s_keyset[i] = kNoKey;
    if (tid < kBoxSet) s_boxset[tid] = ~0ull;
    if (tid == 0) {
        s_n_box = 0;
        s_n_keys = 0;
        s_n_new = 0;
        s_n_act = 0;
        s_ref[3] = 0;
        if (blockIdx.x == 0 && blockIdx.y == 0 && FS.group_bit < 0) {
            // the ring slot the NEXT frame will count into (its last user finished 3 frames ago)
            const int nxt = (ring + 1) % kActiveRing;
            M.counters[kCtrActive0 + nxt] = 0;
            M.counters[kCtrNew0 + nxt] = 0;
        }
    }

    if constexpr (kTma) {
        // one thread arms an mbarrier with the tile's byte count and issues three 2-D TMA tile loads;
        // they land in shared memory while the CTA back-projects its depth samples
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            fence_mbar_init();
            mbar_expect_tx(&s_bar, kTmaTile * kTmaTile * (4 + 4 + 3));
            const int x0 = blockIdx.x * kTmaTile, y0 = blockIdx.y * kTmaTile;
            tma_load_2d(s_td, &maps.depth, x0, y0, &s_bar);
            tma_load_2d(s_tl, &lmap.lam, x0, y0, &s_bar);
            tma_load_2d(s_tc, &maps.color, 3 * x0, y0, &s_bar);
        }
    }

    // ---- boxes: thread s < 64 owns depth sample s of the tile ----
    int lo[3] = {0, 0, 0}, n[3] = {0, 0, 0};
    bool have = false;
    if (tid < kAllocTile * kAllocTile) {
        const int j = (blockIdx.x * kAllocTile + (tid & (kAllocTile - 1))) * P.stride;
        const int i = (blockIdx.y * kAllocTile + (tid / kAllocTile)) * P.stride;
        if (j < P.W && i < P.H) {
            const float d = __ldg(depth + static_cast<size_t>(i) * P.W + j);
            if (d > 0.0f && d < P.depth_trunc) {
                const double z = static_cast<double>(d);
                const double x = __ddiv_rn(__dmul_rn(__dsub_rn(static_cast<double>(j), P.cx), z), P.fx);
                const double y = __ddiv_rn(__dmul_rn(__dsub_rn(static_cast<double>(i), P.cy), z), P.fy);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const double pw = __dadd_rn(
                        __dadd_rn(__dadd_rn(__dmul_rn(pose.Rwc[3 * a + 0], x), __dmul_rn(pose.Rwc[3 * a + 1], y)),
                                  __dmul_rn(pose.Rwc[3 * a + 2], z)),
                        pose.twc[a]);
                    if (P.unit_shift > 0) {
                        // Open3D ScalableTSDFVolume::LocateVolumeUnit: floor(p / volume_unit_length) in float64;
                        // every 8^3 block of a touched unit is touched
                        const int ulo = __double2int_rd(__ddiv_rn(__dsub_rn(pw, P.tau_d), P.unit_len));
                        const int uhi = __double2int_rd(__ddiv_rn(__dadd_rn(pw, P.tau_d), P.unit_len));
                        lo[a] = ulo;
                        n[a] = uhi - ulo + 1;
                    } else {  // decision D1: pyslam float32 key arithmetic (voxel_hashing.h:69-75, 139-151)
                        const int vlo = voxel_coord(__double2float_rn(__dsub_rn(pw, P.tau_d)), P.inv_vs);
                        const int vhi = voxel_coord(__double2float_rn(__dadd_rn(pw, P.tau_d)), P.inv_vs);
                        lo[a] = block_coord(vlo);
                        n[a] = block_coord(vhi) - lo[a] + 1;
                    }
                }
                have = true;
            }
        }
    }
    __syncthreads();  // sets initialised
    if (have && atomicCAS(&s_ref[3], 0, 1) == 0) {
        s_ref[0] = lo[0];
        s_ref[1] = lo[1];
        s_ref[2] = lo[2];
    }

    // ---- pack this CTA's pixel tile into texels (independent of the allocation work) ----
    if constexpr (kTma) {
        mbar_wait(&s_bar, 0);  // s_bar was initialised before the first __syncthreads above
        const int x0 = blockIdx.x * kTmaTile, y0 = blockIdx.y * kTmaTile;
#pragma unroll
        for (int k = 0; k < kTmaTile * kTmaTile / kAllocThreads; ++k) {
            const int q = k * kAllocThreads + tid;
            const int x = x0 + (q & (kTmaTile - 1)), y = y0 + q / kTmaTile;
            if (x < P.W && y < P.H) {
                const float d = s_td[q];
                tex[static_cast<size_t>(y) * P.W + x] = make_texel((d > 0.0f && d < P.depth_trunc) ? d : 0.0f, s_tl[q],
                                                                   s_tc[3 * q], s_tc[3 * q + 1], s_tc[3 * q + 2]);
            }
        }
    } else {
        const int tile = kAllocTile * P.stride;  // pixels per tile side
        const int x0 = blockIdx.x * tile, y0 = blockIdx.y * tile;
        for (int q0 = 0; q0 < tile * tile; q0 += 4 * kAllocThreads) {
            float dv[4], lv[4];
            uint8_t cv[4][3];
            size_t pv[4];
            bool ok[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // four independent pixels per thread: loads issued together
                const int q = q0 + k * kAllocThreads + tid;
                const int x = x0 + q % tile, y = y0 + q / tile;
                ok[k] = q < tile * tile && x < P.W && y < P.H;
                pv[k] = ok[k] ? static_cast<size_t>(y) * P.W + x : 0;
                dv[k] = __ldg(depth + pv[k]);
                lv[k] = __ldg(lam + pv[k]);
                const uint8_t *c = rgb + 3 * pv[k];
                cv[k][0] = __ldg(c);
                cv[k][1] = __ldg(c + 1);
                cv[k][2] = __ldg(c + 2);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (ok[k])
                    tex[pv[k]] = make_texel((dv[k] > 0.0f && dv[k] < P.depth_trunc) ? dv[k] : 0.0f, lv[k], cv[k][0],
                                            cv[k][1], cv[k][2]);
        }
    }
    __syncthreads();  // reference key visible

    // ---- distinct boxes of the tile ----
    if (have) {
        const uint32_t r0 = static_cast<uint32_t>(lo[0] - s_ref[0] + 32768), r1 = static_cast<uint32_t>(lo[1] - s_ref[1] + 32768),
                       r2 = static_cast<uint32_t>(lo[2] - s_ref[2] + 32768);
        bool placed = false;
        if ((r0 | r1 | r2) < 65536u && n[0] <= 15 && n[1] <= 15 && n[2] <= 15) {
            const unsigned long long bk = static_cast<unsigned long long>(r0) | (static_cast<unsigned long long>(r1) << 16) |
                                          (static_cast<unsigned long long>(r2) << 32) |
                                          (static_cast<unsigned long long>(n[0] | (n[1] << 4) | (n[2] << 8)) << 48);
            uint32_t h = mix32(static_cast<uint32_t>(bk) ^ static_cast<uint32_t>(bk >> 32)) & (kBoxSet - 1);
            for (int k = 0; k < kBoxSet && !placed; ++k) {
                const unsigned long long old = atomicCAS(s_boxset + h, ~0ull, bk);
                if (old == ~0ull) {
                    const uint32_t pos = atomicAdd(&s_n_box, 1u);
                    if (pos < kBoxList) {
                        s_box[pos] = bk;
                        placed = true;
                    } else {
                        break;  // list full: handle this box directly below
                    }
                } else if (old == bk) {
                    placed = true;
                }
                h = (h + 1) & (kBoxSet - 1);
            }
        }
        if (!placed) {  // far-away or over-sized box, or > 64 distinct boxes: straight to the table
            for (int dx = 0; dx < n[0]; ++dx)
                for (int dy = 0; dy < n[1]; ++dy)
                    for (int dz = 0; dz < n[2]; ++dz)
                        touch_unit(P, FS, T, M, ring, lo[0] + dx, lo[1] + dy, lo[2] + dz, s_new, &s_n_new, s_act, &s_n_act);
        }
    }
    __syncthreads();

    // ---- distinct keys: expand every distinct box, one candidate unit per thread ----
    {
        const uint32_t nbox = min(s_n_box, static_cast<uint32_t>(kBoxList));
        for (uint32_t item = tid; item < nbox * 32u; item += kAllocThreads) {  // 32 lanes per box (27 typical)
            const unsigned long long bk = s_box[item >> 5];
            const uint32_t c = item & 31u;
            const uint32_t n0 = static_cast<uint32_t>(bk >> 48) & 15u, n1 = static_cast<uint32_t>(bk >> 52) & 15u,
                           n2 = static_cast<uint32_t>(bk >> 56) & 15u;
            // boxes with more than 32 blocks loop over the remainder (c, c + 32, ...); the candidate index is split
            // with multiply-shift divisions (a 32-bit division costs ~20 instructions, and a tile has ~1000 candidates)
            const int bx = s_ref[0] + static_cast<int>(static_cast<uint32_t>(bk) & 0xFFFFu) - 32768;
            const int by = s_ref[1] + static_cast<int>(static_cast<uint32_t>(bk >> 16) & 0xFFFFu) - 32768;
            const int bz = s_ref[2] + static_cast<int>(static_cast<uint32_t>(bk >> 32) & 0xFFFFu) - 32768;
            const uint32_t m2 = s_magic[n2], m1 = s_magic[n1], total = n0 * n1 * n2;
            for (uint32_t cc = c; cc < total; cc += 32u) {
                const uint32_t r = (cc * m2) >> 16, dz = cc - r * n2, dx = (r * m1) >> 16, dy = r - dx * n1;
                const int kx = bx + static_cast<int>(dx), ky = by + static_cast<int>(dy), kz = bz + static_cast<int>(dz);
                const uint32_t rk = rel_key(kx, ky, kz, s_ref);
                bool placed = false;
                if (rk != kNoKey) {
                    uint32_t h = mix32(rk) & (kKeySet - 1);
                    for (int k = 0; k < 96 && !placed; ++k) {
                        const uint32_t old = atomicCAS(s_keyset + h, kNoKey, rk);
                        if (old == kNoKey) {  // first sighting in this tile: queue it for the probe phase
                            const uint32_t pos = atomicAdd(&s_n_keys, 1u);
                            if (pos < kListCap) {
                                s_keys[pos] = rk;
                                placed = true;
                            } else {
                                break;  // list full: probe it right away (below)
                            }
                        } else if (old == rk) {
                            placed = true;
                        }
                        h = (h + 1) & (kKeySet - 1);
                    }
                }
                if (!placed) touch_unit(P, FS, T, M, ring, kx, ky, kz, s_new, &s_n_new, s_act, &s_n_act);
            }
        }
    }
    __syncthreads();

    // ---- probe: every block of every distinct unit of the tile, one per thread, all probes in flight ----
    {
        const uint32_t nkeys = min(s_n_keys, static_cast<uint32_t>(kListCap));
        const int S = P.unit_shift, side = (1 << S) - 1;
        for (uint32_t q = tid; q < (nkeys << (3 * S)); q += kAllocThreads) {
            const uint32_t rk = s_keys[q >> (3 * S)];
            const int sub = static_cast<int>(q & ((1u << (3 * S)) - 1u));
            const int ux = s_ref[0] + static_cast<int>(rk & 1023u) - 512, uy = s_ref[1] + static_cast<int>((rk >> 10) & 1023u) - 512,
                      uz = s_ref[2] + static_cast<int>((rk >> 20) & 1023u) - 512;
            touch_key(P, FS, T, M, ring, (ux << S) + (sub & side), (uy << S) + ((sub >> S) & side), (uz << S) + (sub >> (2 * S)),
                      s_new, &s_n_new, s_act, &s_n_act);
        }
    }
    __syncthreads();

    // ---- flush: one global atomic per list and CTA (three threads, three independent round trips) ----
    const uint32_t n_new = min(s_n_new, static_cast<uint32_t>(kListCap));
    const uint32_t n_act = min(s_n_act, static_cast<uint32_t>(kListCap));
    if (tid == 0) s_base_new = n_new ? atomicAdd(M.counters + kCtrPool, n_new) : 0u;
    uint32_t *list_count = FS.group_bit >= 0 ? M.counters + group_ctr(P.group_buf, kGcUnion) : M.counters + kCtrActive0 + ring;
    if (tid == 32) s_base_act = n_act ? atomicAdd(list_count, n_act) : 0u;
    if (tid == 64 && n_new)
        atomicAdd(FS.group_bit >= 0 ? M.counters + group_ctr(P.group_buf, kGcNew) : M.counters + kCtrNew0 + ring, n_new);
    __syncthreads();
    for (uint32_t k = tid; k < n_new; k += kAllocThreads) assign_block(T, M, s_new[k], s_base_new + k);
    uint32_t *active_out = FS.group_bit >= 0 ? M.union_slots + static_cast<size_t>(P.group_buf) * M.capacity
                                            : M.active_slots + static_cast<size_t>(ring) * M.capacity;
    for (uint32_t k = tid; k < n_act; k += kAllocThreads) {
        const uint32_t g = s_base_act + k;
        if (g < M.capacity) active_out[g] = s_act[k];
    }
}

template <bool kTma>
__global__ void __launch_bounds__(kAllocThreads, 4)
allocate_kernel(const FrameParams P, const float *__restrict__ depth, const uint8_t *__restrict__ rgb,
                const float *__restrict__ lam, float4 *__restrict__ tex, const HashTable T,
                const PoolMeta M, const int ring, const __grid_constant__ FrameMaps maps,
                const __grid_constant__ LambdaMap lmap) {
    allocate_body<kTma>(P, P.pose, FrameSlot{-1, P.frame_id}, depth, rgb, lam, tex, T, M, ring, maps, lmap);
}

// blockIdx.z = frame of the group: one launch allocates for up to kMaxGroup frames
template <bool kTma>
__global__ void __launch_bounds__(kAllocThreads, 8)
allocate_group_kernel(const __grid_constant__ GroupAllocArgs A, const float *__restrict__ lam,
                      const HashTable T, const PoolMeta M) {
    const int k = blockIdx.z;
    allocate_body<kTma>(A.P, A.pose[k], FrameSlot{k, A.frame_id0 + static_cast<uint32_t>(k)}, A.depth[k], A.color[k], lam,
                        A.tex[k], T, M, 0, A.maps[k], A.lmap);
}

cudaError_t launch_allocate_group(const GroupAllocArgs &args, const float *lam, const HashTable &table,
                                  const PoolMeta &meta, cudaStream_t stream) {
    const FrameParams &p = args.P;
    const int gw = (p.W + p.stride - 1) / p.stride;
    const int gh = (p.H + p.stride - 1) / p.stride;
    const dim3 grid((gw + kAllocTile - 1) / kAllocTile, (gh + kAllocTile - 1) / kAllocTile, args.count);
    if (args.use_tma && p.stride * kAllocTile == kTmaTile)
        allocate_group_kernel<true><<<grid, kAllocThreads, 0, stream>>>(args, lam, table, meta);
    else
        allocate_group_kernel<false><<<grid, kAllocThreads, 0, stream>>>(args, lam, table, meta);
    return cudaGetLastError();
}

cudaError_t launch_allocate(const FrameParams &p, const float *depth, const uint8_t *color,
                            const float *lam, float4 *texels, const HashTable &table,
                            const PoolMeta &meta, int ring, const FrameMaps *maps, const LambdaMap *lmap,
                            cudaStream_t stream) {
    const int gw = (p.W + p.stride - 1) / p.stride;
    const int gh = (p.H + p.stride - 1) / p.stride;
    const dim3 grid((gw + kAllocTile - 1) / kAllocTile, (gh + kAllocTile - 1) / kAllocTile);
    if (maps != nullptr && lmap != nullptr && p.stride * kAllocTile == kTmaTile) {
        allocate_kernel<true><<<grid, kAllocThreads, 0, stream>>>(p, depth, color, lam, texels, table, meta,
                                                                 ring, *maps, *lmap);
    } else {
        static const FrameMaps dummy{};
        static const LambdaMap ldummy{};
        allocate_kernel<false><<<grid, kAllocThreads, 0, stream>>>(p, depth, color, lam, texels, table, meta,
                                                                  ring, dummy, ldummy);
    }
    return cudaGetLastError();
}

bool tma_tiles_usable(int W, int stride, const void *depth, const void *color, const void *lam) {
    auto aligned = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    return stride * kAllocTile == kTmaTile && (W % 16) == 0 && aligned(depth) && aligned(color) && aligned(lam);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &st) == cudaSuccess &&
            st == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

static bool encode_2d(CUtensorMap *m, CUtensorMapDataType dt, const void *base, uint64_t w_elems,
                      uint64_t h, uint64_t pitch_bytes, uint32_t box_w, uint32_t box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {w_elems, h};
    const cuuint64_t strides[1] = {pitch_bytes};
    const cuuint32_t box[2] = {box_w, box_h};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool encode_frame_maps(FrameMaps *maps, const float *depth, const uint8_t *color, int H, int W, int tile) {
    return encode_2d(&maps->depth, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, depth, W, H, static_cast<uint64_t>(W) * 4, tile, tile) &&
           encode_2d(&maps->color, CU_TENSOR_MAP_DATA_TYPE_UINT8, color, static_cast<uint64_t>(W) * 3, H,
                     static_cast<uint64_t>(W) * 3, 3 * tile, tile);
}

bool encode_lambda_map(LambdaMap *map, const float *lam, int H, int W, int tile) {
    return encode_2d(&map->lam, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, lam, W, H, static_cast<uint64_t>(W) * 4, tile, tile);
}

// ------------------------------------------------------------------------------------------------
// projective TSDF + colour update
// ------------------------------------------------------------------------------------------------
//
// Arithmetic = Open3D's UniformTSDFVolume::IntegrateWithDepthToCameraDistanceMultiplier in Open3D's own
// operation order (contract v3; bit-identical in tsdf and weight to oracle/open3d_order.c):
//   a block is sub-block s = key - (unit << unit_shift) of its volume unit; voxel (x, y, z) of the unit has
//   x = 8 s.x + lx ...;  centre h = (float)((double)(vl/2 + vl*x) + unit * L) with z taken at the unit's z = 0;
//   p = ((E0*h0 + E1*h1) + E2*h2) + E3 (no FMA), then p += vl * E[:,2] once per z step (INCREMENTAL);
//   u_f = p.x*fx / p.z + cx + 0.5 with IEEE divisions; tsdf = (tsdf*w + t) / (w + 1): mul, add, div.
//
// One CTA iteration = one touched block: 128 threads, each owning a RUN OF 4 VOXELS ALONG z (so the incremental
// projection costs three adds per voxel).  A warp's 32 lanes cover lx 0..7 x ly 0..3: every plane access of a
// warp is one full 128-byte line (LDG.32 / STG.32, coalesced).  The plane loads of a block are issued first; the
// projections and texel gathers (one 16-byte {depth, lambda, rgbx} texel per voxel, packed by allocate_kernel,
// L2-resident) overlap that HBM latency.  Planes are written back only by threads that updated a voxel.

constexpr int kIntThreads = 128;
constexpr int kRun = 4;  // voxels per thread, consecutive in z

// frame-independent geometry of a thread's voxel run
struct VoxelRun {
    float h0, h1, h2;  // Open3D voxel-centre coordinates of the column (x, y) and of the UNIT's first z
    int zskip;         // z steps from the unit's z = 0 to the first voxel of the run
};

__device__ __forceinline__ VoxelRun voxel_run(const uint4 e, const int t, const VolumeConsts &V) {
    const int lx = t & 7, ly = (t >> 3) & 7, z0 = (t >> 6) * kRun;
    const int b[3] = {static_cast<int>(e.x), static_cast<int>(e.y), static_cast<int>(e.z)};
    int u[3], sb[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        u[a] = b[a] >> V.unit_shift;  // floor division by the blocks per unit side
        sb[a] = b[a] - (u[a] << V.unit_shift);
    }
    VoxelRun r;
    // float(half_voxel_length_f + voxel_length_f * x + origin_(0)): float product and sum, widened, plus the
    // float64 unit origin index.cast<double>() * volume_unit_length_, narrowed once
    r.h0 = __double2float_rn(__dadd_rn(
        static_cast<double>(__fadd_rn(V.half_vs, __fmul_rn(V.vs, static_cast<float>(sb[0] * kB + lx)))),
        __dmul_rn(static_cast<double>(u[0]), V.unit_len)));
    r.h1 = __double2float_rn(__dadd_rn(
        static_cast<double>(__fadd_rn(V.half_vs, __fmul_rn(V.vs, static_cast<float>(sb[1] * kB + ly)))),
        __dmul_rn(static_cast<double>(u[1]), V.unit_len)));
    r.h2 = __double2float_rn(__dadd_rn(static_cast<double>(V.half_vs), __dmul_rn(static_cast<double>(u[2]), V.unit_len)));
    r.zskip = sb[2] * kB + z0;
    return r;
}

// ---- IEEE division without the compiler's slow-path scaffolding ------------------------------------------------
// div.rn.f32 expands to MUFU.RCP + a Newton chain + FCHK + a call to a slow path (denormal / huge operands), wrapped
// in BSSY / BSYNC: ~14 issue slots and a dozen register moves per division, three divisions per voxel update.  Here
// the operand range is known, so the fast path is written out: one correctly rounded reciprocal shared by the
// quotients of one denominator, and per quotient two residual corrections (Markstein: with y = RN(1/b) and q
// faithful, RN(q + (a - b q) y) = RN(a / b); the first correction makes q faithful).  Valid for normal operands with
// 2^-100 <= |b| <= 2^100 and |a / b| >= 2^-100 (exact residuals); the callers route anything else to __fdiv_rn.
// tests/test_gpu_tsdf.py::test_fast_division_is_ieee checks rcp_rn_fast over ALL 2^23 significands and div_rn_fast
// against __fdiv_rn on 2^30 operand pairs; the parity tests compare the end results.
constexpr float kDivLo = 7.8886090522101181e-31f;   // 2^-100
constexpr float kDivHi = 1.2676506002282294e+30f;   // 2^100
__device__ __noinline__ float div_rn_slow(const float a, const float b) { return __fdiv_rn(a, b); }  // rare operands
__device__ __forceinline__ float rcp_rn_fast(const float b) {  // RN(1 / b), kDivLo <= |b| <= kDivHi
    float y0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(b));
    return __fmaf_rn(y0, __fmaf_rn(-b, y0, 1.0f), y0);
}
__device__ __forceinline__ float div_rn_fast(const float a, const float b, const float y /* = RN(1/b) */) {
    float q = __fmul_rn(a, y);
    q = __fmaf_rn(__fmaf_rn(-q, b, a), y, q);
    return __fmaf_rn(__fmaf_rn(-q, b, a), y, q);
}

// One frame applied to the kRun voxels of a thread.  F lives in kernel-parameter space.  Everything up to the texel
// gather is branch-free (predicated); the update itself is skipped by warps none of whose voxels is in the band.
__device__ __forceinline__ bool apply_frame(const IntFrame &F, const VoxelRun &r, float *ts, float *w, float *cr,
                                            float *cg, float *cb) {
    float p0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(F.E[0], r.h0), __fmul_rn(F.E[1], r.h1)), __fmul_rn(F.E[2], r.h2)), F.E[3]);
    float p1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(F.E[4], r.h0), __fmul_rn(F.E[5], r.h1)), __fmul_rn(F.E[6], r.h2)), F.E[7]);
    float p2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(F.E[8], r.h0), __fmul_rn(F.E[9], r.h1)), __fmul_rn(F.E[10], r.h2)), F.E[11]);
#pragma unroll 1
    for (int s = 0; s < r.zskip; s += kRun) {  // zskip is a multiple of kRun, uniform across the warp
#pragma unroll
        for (int k = 0; k < kRun; ++k) {
            p0 = __fadd_rn(p0, F.Es[0]);
            p1 = __fadd_rn(p1, F.Es[1]);
            p2 = __fadd_rn(p2, F.Es[2]);
        }
    }
    float pz[kRun];
    int pix[kRun];
    bool rare = false;
#pragma unroll
    for (int k = 0; k < kRun; ++k) {
        pz[k] = p2;
        // p2 <= 0 (or NaN): Open3D skips the voxel; outside [2^-100, 2^100] the fast quotient is not exact (`rare`)
        const bool in_range = p2 >= kDivLo && p2 <= kDivHi;
        rare |= p2 > 0.0f && !in_range;
        const float y = rcp_rn_fast(p2);
        // a quotient below 2^-100 in magnitude may be inexact, but then RN(q + c) = RN(c) either way
        const float u_f = __fadd_rn(__fadd_rn(div_rn_fast(__fmul_rn(p0, F.fxf), p2, y), F.cxf), 0.5f);
        const float v_f = __fadd_rn(__fadd_rn(div_rn_fast(__fmul_rn(p1, F.fyf), p2, y), F.cyf), 0.5f);
        const bool inb = in_range && u_f >= 0.0001f && u_f < F.safe_w && v_f >= 0.0001f && v_f < F.safe_h;
        pix[k] = inb ? __float2int_rz(v_f) * F.W + __float2int_rz(u_f) : -1;
        p0 = __fadd_rn(p0, F.Es[0]);
        p1 = __fadd_rn(p1, F.Es[1]);
        p2 = __fadd_rn(p2, F.Es[2]);
    }
    if (rare) {  // a voxel within 1e-30 m of the camera plane (impossible with a rigid pose): exact divisions
#pragma unroll 1
        for (int k = 0; k < kRun; ++k) {
            const float q2 = pz[k];
            if (q2 > 0.0f && !(q2 >= kDivLo && q2 <= kDivHi)) {
                // p.x, p.y of this voxel: replay the chain from the column base (stepping back is not bit-exact)
                float a0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(F.E[0], r.h0), __fmul_rn(F.E[1], r.h1)), __fmul_rn(F.E[2], r.h2)), F.E[3]);
                float a1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(F.E[4], r.h0), __fmul_rn(F.E[5], r.h1)), __fmul_rn(F.E[6], r.h2)), F.E[7]);
                for (int s = 0; s < r.zskip + k; ++s) {
                    a0 = __fadd_rn(a0, F.Es[0]);
                    a1 = __fadd_rn(a1, F.Es[1]);
                }
                const float u_f = __fadd_rn(__fadd_rn(div_rn_slow(__fmul_rn(a0, F.fxf), q2), F.cxf), 0.5f);
                const float v_f = __fadd_rn(__fadd_rn(div_rn_slow(__fmul_rn(a1, F.fyf), q2), F.cyf), 0.5f);
                const bool inb = u_f >= 0.0001f && u_f < F.safe_w && v_f >= 0.0001f && v_f < F.safe_h;
                pix[k] = inb ? __float2int_rz(v_f) * F.W + __float2int_rz(u_f) : -1;
            }
        }
    }
    float4 tx[kRun];
#pragma unroll
    for (int k = 0; k < kRun; ++k) tx[k] = pix[k] >= 0 ? __ldg(F.tex + pix[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
    bool upd = false;
#pragma unroll
    for (int k = 0; k < kRun; ++k) {
        const float d = tx[k].x;  // 0 where the pixel is invalid (allocate_kernel pre-validates)
        const float sdf = __fmul_rn(__fsub_rn(d, pz[k]), tx[k].y);
        if (d > 0.0f && sdf > -F.tau) {
            const float tv = fminf(1.0f, __fmul_rn(sdf, F.inv_tau));
            const float w0 = w[k];
            const float wn = __fadd_rn(w0, 1.0f);
            const float rc = rcp_rn_fast(wn);  // correctly rounded 1 / (w + 1): weights are integers < 2^24
            const float num = __fadd_rn(__fmul_rn(ts[k], w0), tv);
            // (tsdf*w + t) / (w + 1): exact residuals need |num| >= 2^-100 (num = 0 gives +-0 either way)
            ts[k] = (fabsf(num) >= kDivLo || num == 0.0f) ? div_rn_fast(num, wn, rc) : div_rn_slow(num, wn);
            // colour: float32 running mean; the texel carries r, g, b as binary16 (exact for 0..255)
            const __half2 rg = *reinterpret_cast<const __half2 *>(&tx[k].z);
            const __half2 bx = *reinterpret_cast<const __half2 *>(&tx[k].w);
            cr[k] = __fmul_rn(__fmaf_rn(cr[k], w0, __low2float(rg)), rc);
            cg[k] = __fmul_rn(__fmaf_rn(cg[k], w0, __high2float(rg)), rc);
            cb[k] = __fmul_rn(__fmaf_rn(cb[k], w0, __low2float(bx)), rc);
            w[k] = wn;
            upd = true;
        }
    }
    return upd;
}

// plane access of a thread's run: voxel k of the run sits at  base + 64 k  of each 512-float plane
__device__ __forceinline__ int run_base(const int t) { return (t & 63) + 256 * (t >> 6); }

__device__ __forceinline__ void load_block(const float *blk, float q[kPlanes][kRun]) {
#pragma unroll
    for (int c = 0; c < kPlanes; ++c)
#pragma unroll
        for (int k = 0; k < kRun; ++k) q[c][k] = blk[c * kVox + 64 * k];
}
__device__ __forceinline__ void store_block(float *blk, const float q[kPlanes][kRun]) {
#pragma unroll
    for (int c = 0; c < kPlanes; ++c)
#pragma unroll
        for (int k = 0; k < kRun; ++k) blk[c * kVox + 64 * k] = q[c][k];
}

// sign summary of the block for the mesh extraction (PoolMeta::block_flags): one vote per warp, an atomic only when a
// bit is missing (steady state: one 4-byte read per warp and block visit).  Called by converged warps.
__device__ __forceinline__ void note_signs(uint32_t *flag, const float ts[kRun], const float w[kRun]) {
    bool neg = false, pos = false;
#pragma unroll
    for (int k = 0; k < kRun; ++k) {
        neg |= w[k] != 0.0f && ts[k] < 0.0f;
        pos |= w[k] != 0.0f && !(ts[k] < 0.0f);
    }
    const unsigned need = (__any_sync(0xffffffffu, neg) ? 1u : 0u) | (__any_sync(0xffffffffu, pos) ? 2u : 0u);
    if ((threadIdx.x & 31) == 0 && (*flag & need) != need) atomicOr(flag, need);
}

__global__ void __launch_bounds__(kIntThreads, 8)
integrate_kernel(const __grid_constant__ IntFrame F, const __grid_constant__ VolumeConsts V, const HashTable T,
                 const PoolMeta M, const int ring) {
    const uint32_t n = min(M.counters[kCtrActive0 + ring], M.capacity);
    const uint32_t *__restrict__ act = M.active_slots + static_cast<size_t>(ring) * M.capacity;
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && t == 0) {
        atomicAdd(reinterpret_cast<unsigned long long *>(M.counters + kCtrUpdatesLo),
                  static_cast<unsigned long long>(n));
        atomicAdd(reinterpret_cast<unsigned long long *>(M.counters + kCtrVisitsLo),
                  static_cast<unsigned long long>(n));
    }

    uint32_t i = blockIdx.x;
    uint4 e = make_uint4(0u, 0u, 0u, kNoBlock);
    if (i < n) e = T.entries[act[i]];
    while (i < n) {
        const uint32_t i_next = i + gridDim.x;
        uint4 e_next = e;
        if (i_next < n) e_next = T.entries[act[i_next]];  // in flight during this iteration

        if (e.w < M.capacity) {  // (>= capacity: the pool overflowed for this key)
            float *blk = M.pool + static_cast<size_t>(e.w) * kBlockFloats + run_base(t);
            float q[kPlanes][kRun];
            load_block(blk, q);
            const VoxelRun r = voxel_run(e, t, V);
            const bool upd = apply_frame(F, r, q[0], q[1], q[2], q[3], q[4]);
            if (upd) store_block(blk, q);
            if (__any_sync(0xffffffffu, upd)) note_signs(M.block_flags + e.w, q[0], q[1]);
        }
        e = e_next;
        i = i_next;
    }
}

cudaError_t launch_integrate(const FrameParams &p, const VolumeConsts &vc, const HashTable &table,
                             const PoolMeta &meta, int ring, int grid_ctas, cudaStream_t stream) {
    integrate_kernel<<<grid_ctas, kIntThreads, 0, stream>>>(p.I, vc, table, meta, ring);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// fused group update: a block is loaded once, the frames of the group that touch it are applied in
// frame order while it sits in registers, and it is stored once.  Per voxel the arithmetic is the
// same sequence as frame-by-frame integration, so results are bit-identical; HBM traffic per frame
// drops by the group's overlap factor (consecutive keyframes see mostly the same blocks).
// The per-frame constants are read straight from the kernel-parameter (constant) bank: the frame loop is
// unrolled over the 8 slots of the group, so every constant is an immediate-offset uniform operand.
// ------------------------------------------------------------------------------------------------
constexpr int kUnrolledGroup = 8;   // groups up to this size use the fully unrolled frame loop
// kMinCtas: resident CTAs per SM the register allocation is capped for (8 -> 64 registers, 10 -> 48, 12 -> 40)
template <bool kUnrolled, int kMinCtas>
__global__ void __launch_bounds__(kIntThreads, kMinCtas)
integrate_group_kernel(const __grid_constant__ GroupArgs A, const HashTable T, const PoolMeta M,
                       const int gbuf) {
    __shared__ uint32_t s_next;           // work-stealing: next list position of this CTA
    const uint32_t n = min(M.counters[group_ctr(gbuf, kGcUnion)], M.capacity);
    const uint32_t *__restrict__ list = M.union_slots + static_cast<size_t>(gbuf) * M.capacity;
    const uint32_t *__restrict__ mask = M.group_mask + static_cast<size_t>(gbuf) * (static_cast<size_t>(T.mask) + 1);
    uint32_t *cursor = M.counters + group_ctr(gbuf, kGcNext);
    const int t = threadIdx.x;
    if (blockIdx.x == 0 && t == 0)
        atomicAdd(reinterpret_cast<unsigned long long *>(M.counters + kCtrVisitsLo),
                  static_cast<unsigned long long>(n));
    uint32_t my_cnt = 0;  // thread k < 8: blocks touched by frame k, seen by this CTA

    // dynamic work distribution: blocks cost 1..8 frame updates, static striding leaves a long tail
    uint32_t i = blockIdx.x;  // first item is static; later ones come from the shared cursor
    uint4 e = make_uint4(0u, 0u, 0u, kNoBlock);
    uint32_t m = 0;
    if (i < n) {
        const uint32_t slot = list[i];
        e = T.entries[slot];
        m = mask[slot];
    }
    __syncthreads();
    while (i < n) {
        if (t == 0) s_next = atomicAdd(cursor, 1u) + gridDim.x;
        __syncthreads();
        const uint32_t i_next = s_next;
        uint4 e_next = e;
        uint32_t m_next = 0;
        if (i_next < n) {  // in flight during this iteration
            const uint32_t slot = list[i_next];
            e_next = T.entries[slot];
            m_next = mask[slot];
        }
        if (t < kMaxGroup) my_cnt += (m >> t) & 1u;

        if (e.w < M.capacity) {
            float *blk = M.pool + static_cast<size_t>(e.w) * kBlockFloats + run_base(t);
            float q[kPlanes][kRun];
            load_block(blk, q);
            const VoxelRun r = voxel_run(e, t, A.V);
            bool upd = false;
            if constexpr (kUnrolled) {
#pragma unroll
                for (int k = 0; k < kUnrolledGroup; ++k)  // ascending bits = frame order
                    if ((m >> k) & 1u) upd |= apply_frame(A.f[k], r, q[0], q[1], q[2], q[3], q[4]);
            } else {
                for (uint32_t mm = m; mm; mm &= mm - 1u)  // ascending bits = frame order; constants via LDC
                    upd |= apply_frame(A.f[__ffs(mm) - 1], r, q[0], q[1], q[2], q[3], q[4]);
            }
            if (upd) store_block(blk, q);
            if (__any_sync(0xffffffffu, upd)) note_signs(M.block_flags + e.w, q[0], q[1]);
        }
        e = e_next;
        m = m_next;
        i = i_next;
        __syncthreads();  // s_next is rewritten at the top of the next iteration
    }
    if (t < kMaxGroup && my_cnt) {
        atomicAdd(M.counters + group_ctr(gbuf, kGcTouched0) + t, my_cnt);
        atomicAdd(reinterpret_cast<unsigned long long *>(M.counters + kCtrUpdatesLo),
                  static_cast<unsigned long long>(my_cnt));
    }
}

// clears the membership masks of a finished group (its buffer is reused kGroupBufs groups later)
__global__ void group_clear_kernel(const HashTable T, const PoolMeta M, const int gbuf) {
    const uint32_t n = min(M.counters[group_ctr(gbuf, kGcUnion)], M.capacity);
    uint32_t *mask = M.group_mask + static_cast<size_t>(gbuf) * (static_cast<size_t>(T.mask) + 1);
    const uint32_t *list = M.union_slots + static_cast<size_t>(gbuf) * M.capacity;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) mask[list[i]] = 0u;
}

cudaError_t launch_integrate_group(const GroupArgs &args, const HashTable &table, const PoolMeta &meta,
                                   int group_buf, int grid_ctas, cudaStream_t stream) {
    // B2V_UNROLL=1: groups of <= 8 frames use the frame loop unrolled over the 8 slots (constants become immediate
    // constant-bank operands, but the 75 KB of code miss the instruction cache: measured slower, see profiles/)
    static const bool unroll = [] {
        const char *e = std::getenv("B2V_UNROLL");
        return e != nullptr && std::atoi(e) != 0;
    }();
    const int per_sm = grid_ctas / 148;   // B2V_INT_CTAS_PER_SM selects the occupancy variant (default 8)
    if (unroll && args.count <= kUnrolledGroup)
        integrate_group_kernel<true, 8><<<grid_ctas, kIntThreads, 0, stream>>>(args, table, meta, group_buf);
    else if (per_sm >= 11)
        integrate_group_kernel<false, 12><<<grid_ctas, kIntThreads, 0, stream>>>(args, table, meta, group_buf);
    else if (per_sm >= 9)
        integrate_group_kernel<false, 10><<<grid_ctas, kIntThreads, 0, stream>>>(args, table, meta, group_buf);
    else
        integrate_group_kernel<false, 8><<<grid_ctas, kIntThreads, 0, stream>>>(args, table, meta, group_buf);
    group_clear_kernel<<<148, 256, 0, stream>>>(table, meta, group_buf);
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
    const float xx = __fmul_rn(__fsub_rn(static_cast<float>(u), P.I.cxf), P.inv_fx);
    const float yy = __fmul_rn(__fsub_rn(static_cast<float>(v), P.I.cyf), P.inv_fy);
    lam[static_cast<size_t>(v) * P.W + u] =
        __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xx, xx), __fmul_rn(yy, yy)), 1.0f));  // Open3D: sqrtf(xx*xx + yy*yy + 1)
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
    // the upload replaces the block: its sign summary is recomputed, not accumulated
    const float4 f = src[threadIdx.x], w = src[128 + threadIdx.x];
    const float fs[4] = {f.x, f.y, f.z, f.w}, ws[4] = {w.x, w.y, w.z, w.w};
    bool neg = false, pos = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        neg |= ws[k] != 0.0f && fs[k] < 0.0f;
        pos |= ws[k] != 0.0f && !(fs[k] < 0.0f);
    }
    const int any_neg = __syncthreads_or(neg), any_pos = __syncthreads_or(pos);
    if (threadIdx.x == 0) M.block_flags[dst] = (any_neg ? 1u : 0u) | (any_pos ? 2u : 0u);
}

cudaError_t launch_upload_blocks(const int4 *keys, const float *vox, uint32_t n, uint32_t *scratch_idx,
                                 const HashTable &table, const PoolMeta &meta, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    upload_insert_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keys, n, table, meta, scratch_idx);
    upload_copy_kernel<<<n, 128, 0, stream>>>(vox, scratch_idx, meta);
    return cudaGetLastError();
}

// ---- self-test of the division fast path (b2v_selftest_division) ------------------------------------------------
__global__ void selftest_rcp_kernel(unsigned long long *bad) {
    // every significand, at three exponents
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (1u << 23)) return;
    unsigned n = 0;
    for (uint32_t e : {127u, 100u, 140u}) {
        const float b = __uint_as_float((e << 23) | m);
        n += __float_as_uint(rcp_rn_fast(b)) != __float_as_uint(__frcp_rn(b));
    }
    if (n) atomicAdd(bad, static_cast<unsigned long long>(n));
}
__global__ void selftest_div_kernel(unsigned long long *bad, const uint64_t seed, const uint32_t per_thread) {
    uint64_t s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x + 1);
    unsigned n = 0;
    for (uint32_t i = 0; i < per_thread; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;   // xorshift64
        // denominators as the kernels see them: depths (0.01 .. 40 m), integer weights, plus any exponent in range
        const uint32_t mode = static_cast<uint32_t>(s >> 60);
        float b;
        if (mode < 6) b = __uint_as_float(((120u + (static_cast<uint32_t>(s >> 50) % 12u)) << 23) | (static_cast<uint32_t>(s) & 0x7FFFFFu));
        else if (mode < 10) b = static_cast<float>(1u + (static_cast<uint32_t>(s >> 32) % 70000u));
        else b = __uint_as_float(((30u + (static_cast<uint32_t>(s >> 50) % 195u)) << 23) | (static_cast<uint32_t>(s) & 0x7FFFFFu));
        const uint32_t ea = 60u + (static_cast<uint32_t>(s >> 40) % 120u);
        float a = __uint_as_float((static_cast<uint32_t>(s >> 24) & 0x80000000u) | (ea << 23) | (static_cast<uint32_t>(s >> 17) & 0x7FFFFFu));
        if ((s & 0xFFF00000000ull) == 0) a = 0.0f;
        const float want = __fdiv_rn(a, b);
        const float ab = fabsf(b);
        // the fast path's domain: normal divisor in [2^-100, 2^100], quotient neither below 2^-100 nor overflowing
        const bool ok = ab >= kDivLo && ab <= kDivHi && ((fabsf(want) >= kDivLo && fabsf(want) <= 8.5e37f) || a == 0.0f);
        const float got = ok ? div_rn_fast(a, b, rcp_rn_fast(b)) : want;
        n += __float_as_uint(want) != __float_as_uint(got);
    }
    if (n) atomicAdd(bad, static_cast<unsigned long long>(n));
}

cudaError_t launch_selftest_division(unsigned long long *d_bad, uint64_t pairs, cudaStream_t stream) {
    selftest_rcp_kernel<<<(1u << 23) / 256, 256, 0, stream>>>(d_bad);
    const uint32_t per_thread = 1024;
    const uint64_t threads = (pairs + per_thread - 1) / per_thread;
    selftest_div_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(d_bad + 1, 0x1234567ull, per_thread);
    return cudaGetLastError();
}

}  // namespace b2v
