// b2v_prep.cu — per-frame depth preparation on the GPU ("next" row f#1).
//
// filter_shadow_points (pyslam/utilities/depth.py:103-146): remove the ghost points on depth
// discontinuities.  delta_y = |d[r+2,c] - d[r,c]|, delta_x = |d[r,c+2] - d[r,c]|; the threshold is
// 3 * 1.4826 * median(all positive deltas) (float32 arithmetic, like numpy on float32 input); a pixel is
// set to fill_value if any delta it takes part in exceeds the threshold.
//
// The global median is an exact order statistic: a 3-pass radix select over the float bit patterns
// (positive floats order like their bits), entirely on the device.  For an even count numpy averages the
// two middle elements in float32; both ranks are selected in the same passes.
#include "b2v_internal.h"

namespace b2v {

struct SelectState {
    uint32_t prefix[2];  // bits fixed so far, for rank 0 (lower middle) and rank 1 (upper middle)
    uint32_t rank[2];    // residual rank inside the prefix bucket
    uint32_t total;      // number of positive deltas
    float threshold;     // 3 * 1.4826 * median
};

__device__ __forceinline__ bool delta_at(const float *__restrict__ d, int H, int W, int dx, int dy, int64_t i,
                                         float *out) {
    // index space: first the (H - dy) * W vertical deltas, then the H * (W - dx) horizontal ones
    const int64_t ny = dy > 0 ? static_cast<int64_t>(H - dy) * W : 0;
    if (i < ny) {
        *out = fabsf(__fsub_rn(d[i + static_cast<int64_t>(dy) * W], d[i]));
        return true;
    }
    const int64_t j = i - ny;
    const int wx = W - dx;
    if (dx > 0 && j < static_cast<int64_t>(H) * wx) {
        const int r = static_cast<int>(j / wx), c = static_cast<int>(j % wx);
        const int64_t p = static_cast<int64_t>(r) * W + c;
        *out = fabsf(__fsub_rn(d[p + dx], d[p]));
        return true;
    }
    return false;
}

// pass p in {0,1,2}: digit widths 11, 11, 10 bits from the top
__device__ __forceinline__ void digit_layout(int pass, int *shift, uint32_t *nbins, uint32_t *hi_mask) {
    if (pass == 0) {
        *shift = 21;
        *nbins = 2048;
        *hi_mask = 0u;
    } else if (pass == 1) {
        *shift = 10;
        *nbins = 2048;
        *hi_mask = 0xFFE00000u;
    } else {
        *shift = 0;
        *nbins = 1024;
        *hi_mask = 0xFFFFFC00u;
    }
}

__global__ void __launch_bounds__(256)
shadow_hist_kernel(const float *__restrict__ depth, int H, int W, int dx, int dy, int pass,
                   const SelectState *__restrict__ st, uint32_t *__restrict__ hist) {
    // per-CTA histograms in shared memory (the positive deltas of an image crowd into a few bins: global atomics
    // on them serialise), flushed once
    __shared__ uint32_t s_h[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_h[i] = 0u;
    __syncthreads();
    int shift;
    uint32_t nbins, hi_mask;
    digit_layout(pass, &shift, &nbins, &hi_mask);
    const int64_t n = (dy > 0 ? static_cast<int64_t>(H - dy) * W : 0) + (dx > 0 ? static_cast<int64_t>(H) * (W - dx) : 0);
    const uint32_t p0 = pass ? st->prefix[0] : 0u, p1 = pass ? st->prefix[1] : 0u;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float v;
        if (!delta_at(depth, H, W, dx, dy, i, &v) || !(v > 0.0f)) continue;  // delta_values[delta_values > 0]
        const uint32_t bits = __float_as_uint(v);
        const uint32_t digit = (bits >> shift) & (nbins - 1);
        if ((bits & hi_mask) == p0) atomicAdd(s_h + digit, 1u);
        if ((bits & hi_mask) == p1) atomicAdd(s_h + 2048 + digit, 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096; i += blockDim.x)
        if (s_h[i]) atomicAdd(hist + i, s_h[i]);
}

// one block: walk the two histograms, fix the next digit of both ranks, clear the histograms.  The walk is a
// block-wide scan: thread t owns 8 consecutive bins of each histogram.
__global__ void __launch_bounds__(256)
shadow_pick_kernel(int pass, SelectState *st, uint32_t *hist) {
    __shared__ uint32_t s_part[2][256];
    __shared__ uint32_t s_total;
    int shift;
    uint32_t nbins, hi_mask;
    digit_layout(pass, &shift, &nbins, &hi_mask);
    const int t = threadIdx.x;
    uint32_t mine[2][8];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t b = static_cast<uint32_t>(t) * 8u + j;
            mine[k][j] = b < nbins ? hist[2048 * k + b] : 0u;
            sum += mine[k][j];
        }
        s_part[k][t] = sum;
    }
    __syncthreads();
    if (t == 0) {  // 256 partial sums: exclusive prefixes in place (tiny, serial)
        for (int k = 0; k < 2; ++k) {
            uint32_t run = 0;
            for (int i = 0; i < 256; ++i) {
                const uint32_t v = s_part[k][i];
                s_part[k][i] = run;
                run += v;
            }
            if (k == 0) s_total = run;
        }
        if (pass == 0) {
            const uint32_t total = s_total;
            st->total = total;
            st->rank[0] = total ? (total - 1) / 2 : 0;  // numpy median: mean of elements (n-1)//2 and n//2
            st->rank[1] = total / 2;
            st->prefix[0] = st->prefix[1] = 0;
        }
    }
    __syncthreads();
    const uint32_t ranks[2] = {st->rank[0], st->rank[1]};
    __syncthreads();  // every thread has read the ranks before the owner of the crossing bin rewrites them
    // the bin b with  sum(h[0..b)) <= rank < sum(h[0..b]), or the last bin (the serial walk's stopping rule)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint32_t r = ranks[k];
        uint32_t before = s_part[k][t];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t b = static_cast<uint32_t>(t) * 8u + j;
            if (b < nbins) {
                const bool last = b + 1 == nbins;
                if (r >= before && (r < before + mine[k][j] || last)) {
                    st->rank[k] = r - before;
                    st->prefix[k] |= b << shift;
                }
            }
            before += mine[k][j];
        }
    }
    __syncthreads();
    if (t == 0 && pass == 2) {
        const float a = __uint_as_float(st->prefix[0]), b = __uint_as_float(st->prefix[1]);
        // np.median -> mean of the two middle values in float32; then float32(1.4826) * mad; then 3 * sigma
        const float mad = st->total ? __fmul_rn(__fadd_rn(a, b), 0.5f) : __uint_as_float(0x7FC00000u);
        st->threshold = __fmul_rn(3.0f, __fmul_rn(1.4826f, mad));
    }
    for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x) hist[i] = 0;
}

__global__ void shadow_mask_kernel(const float *__restrict__ depth, int H, int W, int dx, int dy, float fill,
                                   const SelectState *__restrict__ st, float *__restrict__ out) {
    const int64_t n = static_cast<int64_t>(H) * W;
    const float thr = st->threshold;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / W), c = static_cast<int>(i % W);
        const float d = depth[i];
        bool m = false;
        if (dy > 0) {
            if (r >= dy) m |= fabsf(__fsub_rn(d, depth[i - static_cast<int64_t>(dy) * W])) > thr;
            if (r < H - dy) m |= fabsf(__fsub_rn(depth[i + static_cast<int64_t>(dy) * W], d)) > thr;
        }
        if (dx > 0) {
            if (c >= dx) m |= fabsf(__fsub_rn(d, depth[i - dx])) > thr;
            if (c < W - dx) m |= fabsf(__fsub_rn(depth[i + dx], d)) > thr;
        }
        out[i] = m ? fill : d;
    }
}

// scratch: sizeof(SelectState) + 4096 * 4 bytes of device memory, zero-initialised by the caller once
cudaError_t launch_filter_shadow_points(const float *depth, int H, int W, int dx, int dy, float fill, float *out,
                                        void *scratch, cudaStream_t stream) {
    SelectState *st = static_cast<SelectState *>(scratch);
    uint32_t *hist = reinterpret_cast<uint32_t *>(static_cast<char *>(scratch) + 64);
    cudaError_t e = cudaMemsetAsync(scratch, 0, 64 + 4096 * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    for (int pass = 0; pass < 3; ++pass) {
        shadow_hist_kernel<<<296, 256, 0, stream>>>(depth, H, W, dx, dy, pass, st, hist);
        shadow_pick_kernel<<<1, 256, 0, stream>>>(pass, st, hist);
    }
    shadow_mask_kernel<<<296, 256, 0, stream>>>(depth, H, W, dx, dy, fill, st, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// undistort / rectify: cv2.remap with the precomputed maps of initUndistortRectifyMap
// (pyslam/dense/volumetric_integrator_base.py:1017-1054: colour INTER_LINEAR, depth and labels
// INTER_NEAREST, constant zero border).  OpenCV's fixed-point arithmetic is reproduced exactly:
//   linear (8-bit): sx = round_half_even(mapx * 32), pixel = sx >> 5, fraction a = sx & 31;
//                   weights (32-ay)(32-ax)*32, ... (sum 32768); out = (sum w*p + 16384) >> 15
//   nearest:        pixel = round_half_even(mapx)
// ------------------------------------------------------------------------------------------------
__global__ void remap_u8c3_linear_kernel(const uint8_t *__restrict__ src, int H, int W,
                                         const float *__restrict__ mapx, const float *__restrict__ mapy,
                                         uint8_t *__restrict__ dst, int swap_rb) {
    const int64_t n = static_cast<int64_t>(H) * W;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int sx = __float2int_rn(__fmul_rn(mapx[i], 32.0f)), sy = __float2int_rn(__fmul_rn(mapy[i], 32.0f));
        const int x0 = sx >> 5, y0 = sy >> 5, ax = sx & 31, ay = sy & 31;
        const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32,
                  w11 = ay * ax * 32;
        int acc[3] = {16384, 16384, 16384};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int x = x0 + (t & 1), y = y0 + (t >> 1);
            const int w = t == 0 ? w00 : (t == 1 ? w01 : (t == 2 ? w10 : w11));
            if (x >= 0 && x < W && y >= 0 && y < H && w) {  // BORDER_CONSTANT, value 0
                const uint8_t *p = src + (static_cast<int64_t>(y) * W + x) * 3;
                acc[0] += w * p[0];
                acc[1] += w * p[1];
                acc[2] += w * p[2];
            }
        }
        uint8_t *o = dst + i * 3;
        o[swap_rb ? 2 : 0] = static_cast<uint8_t>(acc[0] >> 15);
        o[1] = static_cast<uint8_t>(acc[1] >> 15);
        o[swap_rb ? 0 : 2] = static_cast<uint8_t>(acc[2] >> 15);
    }
}

__global__ void remap_b32_nearest_kernel(const uint32_t *__restrict__ src, int H, int W,
                                         const float *__restrict__ mapx, const float *__restrict__ mapy,
                                         uint32_t *__restrict__ dst) {
    const int64_t n = static_cast<int64_t>(H) * W;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int x = __float2int_rn(mapx[i]), y = __float2int_rn(mapy[i]);
        dst[i] = (x >= 0 && x < W && y >= 0 && y < H) ? src[static_cast<int64_t>(y) * W + x] : 0u;
    }
}

// raw 16-bit depth -> metres: float(u16) * scale, one float32 rounding (numpy: depth.astype(float32) * factor)
__global__ void depth_u16_to_f32_kernel(const uint16_t *__restrict__ src, float *__restrict__ dst, const size_t n,
                                        const float scale) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[i] = __fmul_rn(static_cast<float>(src[i]), scale);
}

cudaError_t launch_depth_u16_to_f32(const uint16_t *src, float *dst, size_t n, float scale, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    depth_u16_to_f32_kernel<<<1184, 256, 0, stream>>>(src, dst, n, scale);
    return cudaGetLastError();
}

cudaError_t launch_remap_u8c3_linear(const uint8_t *src, int H, int W, const float *mapx, const float *mapy,
                                     uint8_t *dst, int swap_rb, cudaStream_t stream) {
    remap_u8c3_linear_kernel<<<296, 256, 0, stream>>>(src, H, W, mapx, mapy, dst, swap_rb);
    return cudaGetLastError();
}

cudaError_t launch_remap_b32_nearest(const void *src, int H, int W, const float *mapx, const float *mapy, void *dst,
                                     cudaStream_t stream) {
    remap_b32_nearest_kernel<<<296, 256, 0, stream>>>(static_cast<const uint32_t *>(src), H, W, mapx, mapy,
                                                      static_cast<uint32_t *>(dst));
    return cudaGetLastError();
}

}  // namespace b2v
