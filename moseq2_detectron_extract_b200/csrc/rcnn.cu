// a4: glue kernels of the repo's own Keypoint/Mask R-CNN graph (model/rcnn.py), built to the reference's detectron2
// configuration (ref: model/config.py:21-94 on top of COCO-Keypoints/keypoint_rcnn_R_50_FPN_3x.yaml):
//   * GroupNorm of the FPN lateral / output convolutions (FPN.NORM = 'GN', 32 groups) on channels-last maps, fused with the
//     top-down path: out = (GN(x) + nearest_up2(top)) * scale      (FPN.FUSE_TYPE = 'avg' -> scale 0.5)
//   * ROIAlignV2 (aligned = true, adaptive sampling grid when sampling_ratio = 0) over the pyramid with detectron2's
//     level assignment, written channels-last (R, P, P, C)
//   * the Fast R-CNN output stage for TEST.DETECTIONS_PER_IMAGE = 1 (soft-max, score threshold, arg-max, box decoding)
//   * detectron2's heatmaps_to_keypoints (bicubic arg-max + the pooled soft-max score)
// All arithmetic in float32 whatever the storage type of the maps (bf16 or float32).
#include "common.cuh"
#include <cuda_bf16.h>
#include <limits.h>
#include <math.h>

namespace msq {
namespace {

// ---- 8 consecutive channels of a channels-last map <-> 8 floats ---------------------------------------------------------
template <typename T> struct Ch8;
template <> struct Ch8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(p);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[2 * k] = __uint_as_float(w[k] << 16); v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<const uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct Ch8<float> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float *p, const float (&v)[8]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};

// =====================================================================================================================
// GroupNorm on (n, H, W, C) channels-last maps; channels per group a multiple of 8 (detectron2: 256 / 32 = 8)
// =====================================================================================================================
constexpr int kGnThreads = 512;

// one CTA per (image, slab of pixels): per-thread float partial sums over <= a few hundred values, combined in double
template <typename T>
__global__ void __launch_bounds__(kGnThreads)
gn_partial_kernel(const T *__restrict__ x, int HW, int C, int slabs, double *__restrict__ partial /* (n, slabs, C/8, 2) */) {
    const int img = blockIdx.x / slabs, slab = blockIdx.x - img * slabs;
    const int vecs = C >> 3;                                      // 8-channel vectors per pixel
    const int lanes = kGnThreads / vecs * vecs;                   // threads in use: a whole number of pixels per sweep
    const int px_per_sweep = lanes / vecs;
    const int px0 = (int)((long long)HW * slab / slabs), px1 = (int)((long long)HW * (slab + 1) / slabs);
    extern __shared__ double red[];                               // (kGnThreads, 2)
    double s = 0.0, ss = 0.0;
    if ((int)threadIdx.x < lanes) {
        const int v = threadIdx.x % vecs, p_off = threadIdx.x / vecs;
        const T *base = x + (size_t)img * HW * C + (size_t)v * 8;
        float fs = 0.f, fss = 0.f;
        int since = 0;
        for (int p = px0 + p_off; p < px1; p += px_per_sweep) {
            float val[8];
            Ch8<T>::load(base + (size_t)p * C, val);
#pragma unroll
            for (int k = 0; k < 8; ++k) { fs += val[k]; fss += val[k] * val[k]; }
            if (++since == 32) { s += fs; ss += fss; fs = fss = 0.f; since = 0; }
        }
        s += fs; ss += fss;
    }
    red[2 * threadIdx.x] = s;
    red[2 * threadIdx.x + 1] = ss;
    __syncthreads();
    if ((int)threadIdx.x < vecs) {                                // thread v sums the pixel rows of its vector
        double a = 0.0, b = 0.0;
        for (int t = threadIdx.x; t < lanes; t += vecs) { a += red[2 * t]; b += red[2 * t + 1]; }
        double *o = partial + (((size_t)img * slabs + slab) * vecs + threadIdx.x) * 2;
        o[0] = a; o[1] = b;
    }
}

__global__ void gn_finish_kernel(const double *__restrict__ partial, int n, int slabs, int vecs, int groups, int HW, int C, float eps,
                                 float *__restrict__ stats /* (n, groups, 2): mean, rstd */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * groups) return;
    const int img = i / groups, g = i - img * groups;
    const int vpg = vecs / groups;                                // vectors per group
    double s = 0.0, ss = 0.0;
    for (int sl = 0; sl < slabs; ++sl)
        for (int v = 0; v < vpg; ++v) {
            const double *p = partial + (((size_t)img * slabs + sl) * vecs + g * vpg + v) * 2;
            s += p[0]; ss += p[1];
        }
    const double cnt = (double)HW * (C / groups);
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// out = ((x - mean) * rstd * gamma + beta + up2(top)) * scale
template <typename T>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const T *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ gamma, const float *__restrict__ beta,
                const T *__restrict__ top, int n, int H, int W, int C, int groups, float scale, T *__restrict__ out) {
    const int vecs = C >> 3, cpg = C / groups;
    const size_t total = (size_t)n * H * W * vecs;
    const int Ht = (H + 1) >> 1, Wt = (W + 1) >> 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vecs);
        const size_t pix = i / vecs;
        const int img = (int)(pix / ((size_t)H * W));
        const int c0 = v * 8, g = c0 / cpg;
        const float mean = stats[2 * (img * groups + g)], rstd = stats[2 * (img * groups + g) + 1];
        float val[8], ga[8], be[8];
        Ch8<T>::load(x + pix * C + c0, val);
        Ch8<float>::load(gamma + c0, ga);
        Ch8<float>::load(beta + c0, be);
#pragma unroll
        for (int k = 0; k < 8; ++k) val[k] = (val[k] - mean) * rstd * ga[k] + be[k];
        if (top) {
            const int rem = (int)(pix - (size_t)img * H * W), y = rem / W, xx = rem - y * W;
            float t[8];
            Ch8<T>::load(top + (((size_t)img * Ht + (y >> 1)) * Wt + (xx >> 1)) * C + c0, t);
#pragma unroll
            for (int k = 0; k < 8; ++k) val[k] += t[k];
        }
        if (scale != 1.f) {
#pragma unroll
            for (int k = 0; k < 8; ++k) val[k] *= scale;
        }
        Ch8<T>::store(out + pix * C + c0, val);
    }
}

// =====================================================================================================================
// ROIAlignV2 over the pyramid, channels-last in and out
// =====================================================================================================================
constexpr int kRoiThreads = 256;
constexpr int kRoiMaxLevels = 8;
struct RoiPyramid {
    const void *feat[kRoiMaxLevels];
    int H[kRoiMaxLevels], W[kRoiMaxLevels];
    float scale[kRoiMaxLevels];
    int n_levels;
};

// One axis of torchvision's roi_align_forward_kernel_impl (aligned = true) for pooled bin `p`: the `grid` sample positions
// start + p * bin + (i + .5) * bin / grid, each bilinear between two neighbouring rows (columns).  Bilinear weights are products
// of a row term and a column term and the samples form a product grid, so the pooled value is
//     out[ph][pw] = sum_rows sum_cols WY[ph][row] * WX[pw][col] * F[row][col]
// with per-axis weights that already hold the 1 / grid of the average.  Rows touched by one bin are consecutive: (first, count).
constexpr int kRoiMaxTaps = 12;
struct __align__(16) AxisTaps {
    float w[kRoiMaxTaps];        // first: the weights of a bin's first four taps are one 16-byte shared-memory load
    int first, count;
    int pad[2];
};
static_assert(sizeof(AxisTaps) == 64 && kRoiMaxTaps == 12, "AxisTaps is one 64-byte record");

// acc += wgt * (8 bf16 / fp32 channels held in `raw`)
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
    typedef uint4 type;
    static __device__ __forceinline__ type load(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ void fma(float (&acc)[8], const type &r, float wgt) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[2 * k] = __fmaf_rn(wgt, __uint_as_float(w[k] << 16), acc[2 * k]);
            acc[2 * k + 1] = __fmaf_rn(wgt, __uint_as_float(w[k] & 0xffff0000u), acc[2 * k + 1]);
        }
    }
};
template <> struct Raw8<float> {
    struct type { float4 a, b; };
    static __device__ __forceinline__ type load(const float *p) {
        type r;
        r.a = __ldg(reinterpret_cast<const float4 *>(p)); r.b = __ldg(reinterpret_cast<const float4 *>(p + 4));
        return r;
    }
    static __device__ __forceinline__ void fma(float (&acc)[8], const type &r, float wgt) {
        const float v[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = __fmaf_rn(wgt, v[k], acc[k]);
    }
};

__device__ __forceinline__ bool build_axis_taps(float start, float bin, int p, int grid, int size, AxisTaps &t) {
    t.first = 0; t.count = 0;
    int lo = INT_MAX, hi = -1;
    for (int i = 0; i < grid; ++i) {                                  // extent first
        float y = start + p * bin + ((float)i + .5f) * bin / (float)grid;
        if (y < -1.0f || y > (float)size) continue;
        if (y <= 0.f) y = 0.f;
        int y_low = (int)y, y_high;
        if (y_low >= size - 1) { y_high = y_low = size - 1; } else y_high = y_low + 1;
        lo = min(lo, y_low); hi = max(hi, y_high);
    }
    if (hi < 0) return true;                                          // every sample outside: the bin is 0
    if (hi - lo + 1 > kRoiMaxTaps) return false;
    t.first = lo; t.count = hi - lo + 1;
#pragma unroll
    for (int k = 0; k < kRoiMaxTaps; ++k) t.w[k] = 0.f;
    const float inv = 1.f / (float)grid;
    for (int i = 0; i < grid; ++i) {
        float y = start + p * bin + ((float)i + .5f) * bin / (float)grid;
        if (y < -1.0f || y > (float)size) continue;
        if (y <= 0.f) y = 0.f;
        int y_low = (int)y, y_high;
        if (y_low >= size - 1) { y_high = y_low = size - 1; y = (float)y_low; } else y_high = y_low + 1;
        const float l = y - y_low, h = 1.f - l;
        t.w[y_low - lo] += h * inv;
        t.w[y_high - lo] += l * inv;
    }
    return true;
}

// up to four taps of one feature row (n of them, n >= 1): loads first, then the fused multiply-adds
template <typename T>
__device__ __forceinline__ void tap_quad(float (&acc)[8], const T *row, int C, int n, float wy, const float4 &wx) {
    typename Raw8<T>::type r0, r1, r2, r3;
    r0 = Raw8<T>::load(row);
    if (n > 1) r1 = Raw8<T>::load(row + C);
    if (n > 2) r2 = Raw8<T>::load(row + 2 * C);
    if (n > 3) r3 = Raw8<T>::load(row + 3 * C);
    Raw8<T>::fma(acc, r0, wy * wx.x);
    if (n > 1) Raw8<T>::fma(acc, r1, wy * wx.y);
    if (n > 2) Raw8<T>::fma(acc, r2, wy * wx.z);
    if (n > 3) Raw8<T>::fma(acc, r3, wy * wx.w);
}

template <typename T>
__global__ void __launch_bounds__(kRoiThreads)
roi_align_v2_kernel(RoiPyramid pyr, int C, const float *__restrict__ boxes, int rois_per_image, int P, int sampling_ratio,
                    int min_level, int canonical_level, float canonical_size, T *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char roi_smem[];
    AxisTaps *taps_y = reinterpret_cast<AxisTaps *>(roi_smem);       // [P]
    AxisTaps *taps_x = taps_y + P;                                    // [P]
    __shared__ int s_direct;
    const int r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 b = *reinterpret_cast<const float4 *>(boxes + (size_t)r * 4);
    // detectron2.modeling.poolers.assign_boxes_to_levels
    int lvl = 0;
    if (pyr.n_levels > 1) {
        const float size = sqrtf((b.z - b.x) * (b.w - b.y));
        const float l = floorf((float)canonical_level + log2f(size / canonical_size + 1e-8f));
        const float lo = (float)min_level, hi = (float)(min_level + pyr.n_levels - 1);
        lvl = (int)fminf(fmaxf(l, lo), hi) - min_level;
        if (!(l == l)) lvl = 0;                                       // NaN sizes (degenerate boxes) land on the lowest level
    }
    const int H = pyr.H[lvl], W = pyr.W[lvl];
    const float scale = pyr.scale[lvl];
    const T *feat = static_cast<const T *>(pyr.feat[lvl]) + (size_t)(r / rois_per_image) * H * W * C;
    // torchvision roi_align_forward_kernel_impl, aligned = true
    const float x0 = b.x * scale - 0.5f, y0 = b.y * scale - 0.5f, x1 = b.z * scale - 0.5f, y1 = b.w * scale - 0.5f;
    const float roi_w = x1 - x0, roi_h = y1 - y0;
    const float bin_h = roi_h / (float)P, bin_w = roi_w / (float)P;
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / (float)P);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / (float)P);
    if (threadIdx.x == 0) s_direct = 0;
    __syncthreads();
    if ((int)threadIdx.x < 2 * P) {
        const bool is_y = (int)threadIdx.x < P;
        const int p = is_y ? threadIdx.x : threadIdx.x - P;
        AxisTaps &t = (is_y ? taps_y : taps_x)[p];                   // built in place (dynamic indexing: shared, not local, memory)
        const bool ok = is_y ? build_axis_taps(y0, bin_h, p, gh, H, t) : build_axis_taps(x0, bin_w, p, gw, W, t);
        if (!ok) s_direct = 1;
    }
    __syncthreads();
    const int bins = P * P, vecs = C >> 3;
    T *dst = out + (size_t)r * bins * C;
    if (!s_direct) {
        for (int bin = warp; bin < bins; bin += kRoiThreads / 32) {
            const int ph = bin / P, pw = bin - ph * P;
            const AxisTaps &ty = taps_y[ph], &tx = taps_x[pw];
            const int cy = ty.count, cx = tx.count;                   // warp-uniform
            const int row_stride = W * C;                             // elements; a level of one image is far below 2^31
            for (int v = lane; v < vecs; v += 32) {
                float acc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = 0.f;
                const T *p0 = feat + ((size_t)ty.first * W + tx.first) * C + v * 8;
                if (cx <= 4) {
                    // the common case (sampling grid <= 3 per bin): the x weights live in registers, the taps of a row are
                    // loaded together, then accumulated in the order of the general loop (same rounding)
                    const float4 wx = *reinterpret_cast<const float4 *>(tx.w);
                    for (int j = 0; j < (cx > 0 ? cy : 0); ++j) tap_quad<T>(acc, p0 + j * row_stride, C, cx, ty.w[j], wx);
                } else {
                    for (int j = 0; j < cy; ++j) {
                        const float wy = ty.w[j];
                        const T *row = p0 + j * row_stride;
                        for (int i = 0; i < cx; ++i) Raw8<T>::fma(acc, Raw8<T>::load(row + i * C), wy * tx.w[i]);
                    }
                }
                Ch8<T>::store(dst + (size_t)bin * C + v * 8, acc);
            }
        }
        return;
    }
    // very elongated boxes (more than kRoiMaxTaps rows under one bin): sample by sample, as torchvision loops
    const float count = (float)max(gh * gw, 1);
    for (int bin = warp; bin < bins; bin += kRoiThreads / 32) {
        const int ph = bin / P, pw = bin - ph * P;
        for (int v = lane; v < vecs; v += 32) {
            const T *base = feat + v * 8;
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
            for (int iy = 0; iy < gh; ++iy) {
                float y = y0 + ph * bin_h + ((float)iy + .5f) * bin_h / (float)gh;
                const bool y_out = y < -1.0f || y > (float)H;
                if (y <= 0.f) y = 0.f;
                int y_low = (int)y, y_high;
                if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
                const float ly = y - y_low, hy = 1.f - ly;
                const size_t row_lo = (size_t)y_low * W * C, row_hi = (size_t)y_high * W * C;
                for (int ix = 0; ix < gw; ++ix) {
                    float x = x0 + pw * bin_w + ((float)ix + .5f) * bin_w / (float)gw;
                    if (y_out || x < -1.0f || x > (float)W) continue;
                    if (x <= 0.f) x = 0.f;
                    int x_low = (int)x, x_high;
                    if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
                    const float lx = x - x_low, hx = 1.f - lx;
                    const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                    float v1[8], v2[8], v3[8], v4[8];
                    Ch8<T>::load(base + row_lo + (size_t)x_low * C, v1);
                    Ch8<T>::load(base + row_lo + (size_t)x_high * C, v2);
                    Ch8<T>::load(base + row_hi + (size_t)x_low * C, v3);
                    Ch8<T>::load(base + row_hi + (size_t)x_high * C, v4);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        acc[k] += __fmaf_rn(w4, v4[k], __fmaf_rn(w3, v3[k], __fmaf_rn(w2, v2[k], w1 * v1[k])));
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = acc[k] / count;
            Ch8<T>::store(dst + (size_t)bin * C + v * 8, acc);
        }
    }
}

// =====================================================================================================================
// Fast R-CNN outputs for one detection per image: detectron2 fast_rcnn_inference_single_image with topk_per_image = 1.
// The best-scoring class-0 box above the threshold always survives its NMS, so the result is an arg-max.
// pred (R, stride) float32: [logit_fg, logit_bg, dx, dy, dw, dh, ...]; proposals (n, k, 4); counts (n) valid proposals.
// =====================================================================================================================
__global__ void __launch_bounds__(128)
fastrcnn_top1_kernel(const float *__restrict__ pred, int stride, const float *__restrict__ proposals, const int *__restrict__ counts,
                     int n, int k, float img_h, float img_w, float score_thresh, float wx, float wy, float ww, float wh, float scale_clamp,
                     float *__restrict__ box_out, float *__restrict__ score_out, uint8_t *__restrict__ has_out, int *__restrict__ index_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x * 4 + warp;
    if (img >= n) return;
    const int cnt = counts ? min(counts[img], k) : k;
    float best = -1.f;
    int best_i = INT_MAX;
    for (int i = lane; i < cnt; i += 32) {
        const float *p = pred + ((size_t)img * k + i) * stride;
        const float a = p[0], bg = p[1], m = fmaxf(a, bg);
        const float ea = expf(a - m), eb = expf(bg - m);
        const float s = ea / (ea + eb);
        if (s > score_thresh && s == s && s > best) { best = s; best_i = i; }       // i ascends per lane: first maximum stays
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane != 0) return;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has = best_i != INT_MAX;
    if (has) {
        const float *p = pred + ((size_t)img * k + best_i) * stride;
        const float4 pb = *reinterpret_cast<const float4 *>(proposals + ((size_t)img * k + best_i) * 4);
        const float width = pb.z - pb.x, height = pb.w - pb.y;
        const float cx = pb.x + 0.5f * width, cy = pb.y + 0.5f * height;
        const float dx = p[2] / wx, dy = p[3] / wy, dw = fminf(p[4] / ww, scale_clamp), dh = fminf(p[5] / wh, scale_clamp);
        const float pcx = dx * width + cx, pcy = dy * height + cy, pw = expf(dw) * width, phh = expf(dh) * height;
        o.x = fminf(fmaxf(pcx - 0.5f * pw, 0.f), img_w);
        o.y = fminf(fmaxf(pcy - 0.5f * phh, 0.f), img_h);
        o.z = fminf(fmaxf(pcx + 0.5f * pw, 0.f), img_w);
        o.w = fminf(fmaxf(pcy + 0.5f * phh, 0.f), img_h);
        if (!(isfinite(o.x) && isfinite(o.y) && isfinite(o.z) && isfinite(o.w))) o = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    *reinterpret_cast<float4 *>(box_out + (size_t)img * 4) = o;
    score_out[img] = has ? best : 0.f;
    has_out[img] = has ? 1 : 0;
    if (index_out) index_out[img] = has ? best_i : -1;
}

// =====================================================================================================================
// detectron2.structures.keypoints.heatmaps_to_keypoints for all RoIs: one CTA per (RoI, keypoint)
// =====================================================================================================================
__device__ __forceinline__ float cub1(float x) { return ((-0.75f + 2.f) * x - (-0.75f + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cub2(float x) { return ((-0.75f * x - 5.f * -0.75f) * x + 8.f * -0.75f) * x - 4.f * -0.75f; }
__device__ __forceinline__ void cub_coeffs(float t, float c[4]) {
    c[0] = cub2(t + 1.f); c[1] = cub1(t);
    const float u = 1.f - t;
    c[2] = cub1(u); c[3] = cub2(u + 1.f);
}

constexpr int kKpThreads = 256;

__global__ void __launch_bounds__(kKpThreads)
keypoint_decode_d2_kernel(const float *__restrict__ maps, const float *__restrict__ rois, int K, int Hm, int Wm,
                          float *__restrict__ xyp /* (R, K, 3): x, y, probability */, float *__restrict__ logit /* (R, K) or null */) {
    extern __shared__ float heat[];
    __shared__ float best_v[kKpThreads / 32];
    __shared__ int best_i[kKpThreads / 32];
    __shared__ float part[kKpThreads / 32];
    __shared__ float s_max;
    const int roi = blockIdx.x / K, kp = blockIdx.x - roi * K;
    const float *m = maps + ((size_t)roi * K + kp) * Hm * Wm;
    for (int i = threadIdx.x; i < Hm * Wm; i += kKpThreads) heat[i] = m[i];
    const float x1 = rois[4 * roi], y1 = rois[4 * roi + 1], x2 = rois[4 * roi + 2], y2 = rois[4 * roi + 3];
    const float width = fmaxf(x2 - x1, 1.f), height = fmaxf(y2 - y1, 1.f);
    const int ow = (int)ceilf(width), oh = (int)ceilf(height);
    const float sx = (float)Wm / (float)ow, sy = (float)Hm / (float)oh;
    __syncthreads();
    float bv = -INFINITY;
    int bi = INT_MAX;
    for (int p = threadIdx.x; p < ow * oh; p += kKpThreads) {
        const int oy = p / ow, ox = p - oy * ow;
        const float rx = sx * ((float)ox + 0.5f) - 0.5f, ry = sy * ((float)oy + 0.5f) - 0.5f;
        const float fx = floorf(rx), fy = floorf(ry);
        const int ix = (int)fx, iy = (int)fy;
        float cx[4], cy[4];
        cub_coeffs(rx - fx, cx);
        cub_coeffs(ry - fy, cy);
        float row[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float *rr = heat + min(max(iy - 1 + a, 0), Hm - 1) * Wm;
            row[a] = rr[min(max(ix - 1, 0), Wm - 1)] * cx[0] + rr[min(max(ix, 0), Wm - 1)] * cx[1] +
                     rr[min(max(ix + 1, 0), Wm - 1)] * cx[2] + rr[min(max(ix + 2, 0), Wm - 1)] * cx[3];
        }
        const float v = row[0] * cy[0] + row[1] * cy[1] + row[2] * cy[2] + row[3] * cy[3];
        if (v > bv) { bv = v; bi = p; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { best_v[threadIdx.x >> 5] = bv; best_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wv = 1; wv < kKpThreads / 32; ++wv)
            if (best_v[wv] > bv || (best_v[wv] == bv && best_i[wv] < bi)) { bv = best_v[wv]; bi = best_i[wv]; }
        if (bi == INT_MAX) bi = 0;
        best_i[0] = bi;
        s_max = bv;
    }
    __syncthreads();
    // sum over the POOL-resolution map of exp(map - max of the full-resolution map)
    const float mx = s_max;
    float acc = 0.f;
    for (int i = threadIdx.x; i < Hm * Wm; i += kKpThreads) acc += expf(heat[i] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
        for (int wv = 0; wv < kKpThreads / 32; ++wv) total += part[wv];
        bi = best_i[0];
        const int yi = bi / ow, xi = bi - yi * ow;
        float *o = xyp + ((size_t)roi * K + kp) * 3;
        o[0] = ((float)xi + 0.5f) * (width / (float)ow) + x1;
        o[1] = ((float)yi + 0.5f) * (height / (float)oh) + y1;
        o[2] = 1.f / total;                                       // exp(max - max) / sum(exp(pool - max))
        if (logit) logit[(size_t)roi * K + kp] = mx;
    }
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" size_t msq_group_norm_scratch_bytes(int n, int H, int W, int C) {
    if (n <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
    const int slabs = std::max(1, std::min(64, (H * W + 511) / 512));
    return align_up((size_t)n * slabs * (C / 8) * 2 * sizeof(double), 256) + align_up((size_t)n * C * sizeof(float), 256);
}

extern "C" int msq_group_norm_nhwc(const void *x, int is_bf16, int n, int H, int W, int C, int groups, float eps, const float *gamma,
                                   const float *beta, const void *top, float scale, void *out, void *scratch, size_t scratch_bytes,
                                   void *stream) {
    MSQ_REQUIRE(n >= 0 && H > 0 && W > 0 && C > 0 && groups > 0, MSQ_EINVAL, "msq_group_norm_nhwc: bad sizes");
    MSQ_REQUIRE(C % groups == 0 && (C / groups) % 8 == 0 && C / 8 <= kGnThreads, MSQ_EUNSUPPORTED,
                "msq_group_norm_nhwc: channels per group must be a multiple of 8 and C <= %d (C=%d, groups=%d)", kGnThreads * 8, C, groups);
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(x && gamma && beta && out && scratch, MSQ_EINVAL, "msq_group_norm_nhwc: null pointer");
    MSQ_REQUIRE(scratch_bytes >= msq_group_norm_scratch_bytes(n, H, W, C), MSQ_ENOMEM, "msq_group_norm_nhwc: scratch too small");
    MSQ_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)out % 16 == 0 && (uintptr_t)gamma % 16 == 0 && (uintptr_t)beta % 16 == 0 &&
                (!top || (uintptr_t)top % 16 == 0) && (uintptr_t)scratch % 16 == 0, MSQ_EINVAL, "msq_group_norm_nhwc: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W, vecs = C / 8;
    const int slabs = std::max(1, std::min(64, (HW + 511) / 512));
    double *partial = static_cast<double *>(scratch);
    float *stats = reinterpret_cast<float *>(static_cast<char *>(scratch) + align_up((size_t)n * slabs * vecs * 2 * sizeof(double), 256));
    const size_t smem = (size_t)kGnThreads * 2 * sizeof(double);
    TimedLaunch timed(K_DETECTOR_GLUE, st, 3);
    if (is_bf16) gn_partial_kernel<__nv_bfloat16><<<n * slabs, kGnThreads, smem, st>>>(static_cast<const __nv_bfloat16 *>(x), HW, C, slabs, partial);
    else         gn_partial_kernel<float><<<n * slabs, kGnThreads, smem, st>>>(static_cast<const float *>(x), HW, C, slabs, partial);
    gn_finish_kernel<<<(n * groups + 127) / 128, 128, 0, st>>>(partial, n, slabs, vecs, groups, HW, C, eps, stats);
    const size_t work = (size_t)n * HW * vecs;
    const int blocks = (int)std::min<size_t>((work + 255) / 256, (size_t)sm_count() * 16);
    if (is_bf16) gn_apply_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(x), stats, gamma, beta,
                                                                      static_cast<const __nv_bfloat16 *>(top), n, H, W, C, groups, scale,
                                                                      static_cast<__nv_bfloat16 *>(out));
    else         gn_apply_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float *>(x), stats, gamma, beta, static_cast<const float *>(top),
                                                                n, H, W, C, groups, scale, static_cast<float *>(out));
    MSQ_LAUNCH_OK("group_norm_nhwc");
    return MSQ_OK;
}

extern "C" int msq_roi_align_v2(const void *const *feat_dev, const int *heights, const int *widths, const float *scales, int n_levels,
                                int C, int is_bf16, const float *boxes_dev, int n_rois, int rois_per_image, int P, int sampling_ratio,
                                int min_level, int canonical_level, float canonical_size, void *out_dev, void *stream) {
    MSQ_REQUIRE(n_levels >= 1 && n_levels <= kRoiMaxLevels, MSQ_EINVAL, "msq_roi_align_v2: 1..%d pyramid levels (got %d)", kRoiMaxLevels, n_levels);
    MSQ_REQUIRE(C > 0 && C % 8 == 0, MSQ_EUNSUPPORTED, "msq_roi_align_v2: channel count must be a multiple of 8 (got %d)", C);
    MSQ_REQUIRE(P >= 1 && sampling_ratio >= 0 && rois_per_image >= 1 && n_rois >= 0, MSQ_EINVAL, "msq_roi_align_v2: bad sizes");
    if (n_rois == 0) return MSQ_OK;
    MSQ_REQUIRE(feat_dev && heights && widths && scales && boxes_dev && out_dev, MSQ_EINVAL, "msq_roi_align_v2: null pointer");
    MSQ_REQUIRE((uintptr_t)boxes_dev % 16 == 0 && (uintptr_t)out_dev % 16 == 0, MSQ_EINVAL, "msq_roi_align_v2: boxes / output must be 16-byte aligned");
    RoiPyramid pyr;
    pyr.n_levels = n_levels;
    for (int l = 0; l < n_levels; ++l) {
        MSQ_REQUIRE(feat_dev[l] && heights[l] > 0 && widths[l] > 0 && (uintptr_t)feat_dev[l] % 16 == 0, MSQ_EINVAL,
                    "msq_roi_align_v2: level %d: bad feature map", l);
        pyr.feat[l] = feat_dev[l]; pyr.H[l] = heights[l]; pyr.W[l] = widths[l]; pyr.scale[l] = scales[l];
    }
    cudaStream_t st = (cudaStream_t)stream;
    MSQ_REQUIRE(2 * P <= kRoiThreads && (size_t)2 * P * sizeof(AxisTaps) <= 48 * 1024, MSQ_EUNSUPPORTED, "msq_roi_align_v2: pooled size %d is too large", P);
    const size_t taps_smem = (size_t)2 * P * sizeof(AxisTaps);
    TimedLaunch timed(K_DETECTOR_GLUE, st);
    if (is_bf16)
        roi_align_v2_kernel<__nv_bfloat16><<<n_rois, kRoiThreads, taps_smem, st>>>(pyr, C, boxes_dev, rois_per_image, P, sampling_ratio, min_level,
                                                                          canonical_level, canonical_size, static_cast<__nv_bfloat16 *>(out_dev));
    else
        roi_align_v2_kernel<float><<<n_rois, kRoiThreads, taps_smem, st>>>(pyr, C, boxes_dev, rois_per_image, P, sampling_ratio, min_level,
                                                                  canonical_level, canonical_size, static_cast<float *>(out_dev));
    MSQ_LAUNCH_OK("roi_align_v2");
    return MSQ_OK;
}

extern "C" int msq_fastrcnn_top1(const float *pred_dev, int pred_stride, const float *proposals_dev, const int32_t *counts_dev, int n, int k,
                                 int img_h, int img_w, float score_thresh, const float *weights4_host, float *box_dev, float *score_dev,
                                 uint8_t *has_dev, int32_t *index_dev, void *stream) {
    MSQ_REQUIRE(n >= 0 && k >= 1 && pred_stride >= 6 && img_h > 0 && img_w > 0, MSQ_EINVAL, "msq_fastrcnn_top1: bad sizes");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(pred_dev && proposals_dev && weights4_host && box_dev && score_dev && has_dev, MSQ_EINVAL, "msq_fastrcnn_top1: null pointer");
    MSQ_REQUIRE((uintptr_t)proposals_dev % 16 == 0 && (uintptr_t)box_dev % 16 == 0, MSQ_EINVAL, "msq_fastrcnn_top1: boxes must be 16-byte aligned");
    TimedLaunch timed(K_DETECTOR_GLUE, (cudaStream_t)stream);
    fastrcnn_top1_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(pred_dev, pred_stride, proposals_dev, counts_dev, n, k, (float)img_h,
                                                                        (float)img_w, score_thresh, weights4_host[0], weights4_host[1],
                                                                        weights4_host[2], weights4_host[3], logf(1000.f / 16.f), box_dev,
                                                                        score_dev, has_dev, index_dev);
    MSQ_LAUNCH_OK("fastrcnn_top1");
    return MSQ_OK;
}

// torch.nn.functional.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) of the keypoint head's deconvolution
// output (detectron2 KRCNNConvDeconvUpsampleHead.layers): (R, K, H, W) in any strides, bf16 or fp32 -> (R, K, 2H, 2W) fp32 dense.
// One thread per output pixel (torch's kernel walks all R * K planes inside every thread: 1 ms for 3 MB of output).
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_bilinear_kernel(const T *__restrict__ in, long long sN, long long sC, long long sH, long long sW, int planes, int K, int H, int W,
                           float *__restrict__ out) {
    const int OW = 2 * W, OH = 2 * H;
    const long long total = (long long)planes * OH * OW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW), oy = (int)((i / OW) % OH), pl = (int)(i / ((long long)OW * OH));
        const int r = pl / K, k = pl - r * K;
        // area_pixel_compute_source_index(scale = 0.5, align_corners = false): max(0.5 * (dst + 0.5) - 0.5, 0)
        const float ys = fmaxf(0.5f * ((float)oy + 0.5f) - 0.5f, 0.f), xs = fmaxf(0.5f * ((float)ox + 0.5f) - 0.5f, 0.f);
        const int y1 = (int)ys, x1 = (int)xs;
        const int yp = y1 < H - 1 ? 1 : 0, xp = x1 < W - 1 ? 1 : 0;
        const float ly = ys - (float)y1, hy = 1.f - ly, lx = xs - (float)x1, hx = 1.f - lx;
        const T *p = in + r * sN + k * sC + y1 * sH + x1 * sW;
        const float v00 = (float)p[0], v01 = (float)p[xp * sW], v10 = (float)p[yp * sH], v11 = (float)p[yp * sH + xp * sW];
        out[i] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
    }
}

extern "C" int msq_upsample2x_bilinear(const void *in, int in_is_bf16, long long sN, long long sC, long long sH, long long sW, int n, int K,
                                       int H, int W, float *out, void *stream) {
    MSQ_REQUIRE(n >= 0 && K > 0 && H > 0 && W > 0, MSQ_EINVAL, "msq_upsample2x_bilinear: bad sizes");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(in && out, MSQ_EINVAL, "msq_upsample2x_bilinear: null pointer");
    const long long total = (long long)n * K * 4 * H * W;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
    TimedLaunch timed(K_DETECTOR_GLUE, (cudaStream_t)stream);
    if (in_is_bf16)
        upsample2x_bilinear_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const __nv_bfloat16 *>(in), sN, sC, sH, sW,
                                                                                       n * K, K, H, W, out);
    else
        upsample2x_bilinear_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const float *>(in), sN, sC, sH, sW, n * K, K, H, W, out);
    MSQ_LAUNCH_OK("upsample2x_bilinear");
    return MSQ_OK;
}

extern "C" int msq_keypoints_from_heatmaps_d2(const float *maps, const float *rois, int n_rois, int K, int Hm, int Wm, float *xyp,
                                              float *logit, void *stream) {
    MSQ_REQUIRE(n_rois >= 0 && K > 0 && Hm > 0 && Wm > 0, MSQ_EINVAL, "msq_keypoints_from_heatmaps_d2: bad sizes");
    if (n_rois == 0) return MSQ_OK;
    MSQ_REQUIRE(maps && rois && xyp, MSQ_EINVAL, "msq_keypoints_from_heatmaps_d2: null pointer");
    const size_t smem = (size_t)Hm * Wm * sizeof(float);
    MSQ_REQUIRE(smem <= 200 * 1024, MSQ_EUNSUPPORTED, "msq_keypoints_from_heatmaps_d2: %dx%d heatmaps are too large", Hm, Wm);
    if (smem > 48 * 1024)
        MSQ_CUDA_OK(cudaFuncSetAttribute(keypoint_decode_d2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TimedLaunch timed(K_DETECTOR_GLUE, (cudaStream_t)stream);
    keypoint_decode_d2_kernel<<<n_rois * K, kKpThreads, smem, (cudaStream_t)stream>>>(maps, rois, K, Hm, Wm, xyp, logit);
    MSQ_LAUNCH_OK("keypoint_decode_d2");
    return MSQ_OK;
}
