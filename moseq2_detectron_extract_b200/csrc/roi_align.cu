// a4 glue: multi-level RoIAlign for the R-CNN's RoI heads -- torchvision.ops.MultiScaleRoIAlign in ONE launch.
//   torchvision (ops/poolers.py:_multiscale_roi_align) runs, per pyramid level, torch.where (a host sync), a gather of the
//   RoIs, roi_align with one THREAD per output element on NCHW float32 features (16 scattered 4-byte loads each) and an
//   index_put into the result: 31 ms for the 100 000 box RoIs of a 1000-frame batch, a quarter of the whole model.
// Here one CTA owns one RoI: every warp takes pooled bins in turn, its lanes own 8 consecutive CHANNELS of the
// channels-last feature map (the backbone's native layout under autocast), so each of the 4 x sampling_ratio^2 bilinear taps
// of a bin is one coalesced 16-byte load per lane; the (C, P, P) block of the RoI is assembled in shared memory and
// written out contiguously in the dtype of the features (bf16 under autocast: what the box head's first Linear reads).
// Arithmetic = torchvision's roi_align_forward_kernel_impl<float> (aligned = false) in float32, without fused
// multiply-adds (results agree to an ulp of float32 before the final rounding).
#include "common.cuh"
#include <cuda_bf16.h>

namespace msq {
namespace {

constexpr int kAlignThreads = 256;
constexpr int kMaxLevels = 8;

struct PyramidArg {
    const void *feat[kMaxLevels];      // (n, H_l, W_l, C) channels-last
    int H[kMaxLevels], W[kMaxLevels];
    float scale[kMaxLevels];
    int n_levels;
};

// acc / count, IEEE-exact: a power-of-two count (sampling 1, 2, 4, 8) multiplies by its exact reciprocal
__device__ __forceinline__ float __fdividef_exact(float a, float count) {
    const int c = (int)count;
    return (c & (c - 1)) == 0 ? a * (1.0f / count) : a / count;
}

template <typename T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(p);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[2 * k] = __uint_as_float(w[k] << 16); v[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
    }
    static __device__ __forceinline__ __nv_bfloat16 store(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float *p, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ float store(float v) { return v; }
};

constexpr int kMaxSampling = 4;

template <typename T, int S>                                       // S = sampling ratio (compile time: the sample loops unroll)
__global__ void __launch_bounds__(kAlignThreads)
roi_align_levels_kernel(PyramidArg pyr, int C, const float *__restrict__ rois, const long long *__restrict__ levels, int P,
                        T *__restrict__ out) {
    constexpr int sampling = S;
    extern __shared__ unsigned char smem_raw[];
    T *block = reinterpret_cast<T *>(smem_raw);                    // (C, P*P) of this RoI
    const int r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *roi = rois + (size_t)r * 5;
    const int lvl = pyr.n_levels > 1 ? (int)levels[r] : 0;
    const int H = pyr.H[lvl], W = pyr.W[lvl];
    const float scale = pyr.scale[lvl];
    const T *feat = static_cast<const T *>(pyr.feat[lvl]) + (size_t)(int)roi[0] * H * W * C;
    const float x0 = roi[1] * scale, y0 = roi[2] * scale, x1 = roi[3] * scale, y1 = roi[4] * scale;
    const float roi_w = fmaxf(x1 - x0, 1.f), roi_h = fmaxf(y1 - y0, 1.f);
    const float bin_h = roi_h / (float)P, bin_w = roi_w / (float)P;
    const float count = (float)max(sampling * sampling, 1);
    // the sample offsets inside a bin, (i + .5) * bin / sampling, are the same for every bin: divide once per RoI
    float off_y[S], off_x[S];
#pragma unroll
    for (int i = 0; i < S; ++i) {
        off_y[i] = (i + .5f) * bin_h / (float)sampling;
        off_x[i] = (i + .5f) * bin_w / (float)sampling;
    }
    const int bins = P * P, groups = (C + 255) / 256;
    for (int bin = warp; bin < bins; bin += kAlignThreads / 32) {
        const int ph = bin / P, pw = bin - ph * P;
        const float by = y0 + ph * bin_h, bx = x0 + pw * bin_w;
        for (int g = 0; g < groups; ++g) {
            const int c0 = (g * 32 + lane) * 8;
            const bool act = c0 < C;
            const T *base = feat + c0;
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
            for (int iy = 0; iy < S; ++iy) {
                float y = by + off_y[iy];
                const bool y_out = y < -1.0f || y > (float)H;
                if (y <= 0.f) y = 0.f;
                int y_low = (int)y, y_high;
                if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
                const float ly = y - y_low, hy = 1.f - ly;
                const int row_lo = y_low * W * C, row_hi = y_high * W * C;          // element offsets inside one image's map
#pragma unroll
                for (int ix = 0; ix < S; ++ix) {
                    float x = bx + off_x[ix];
                    if (y_out || x < -1.0f || x > (float)W) continue;               // sample outside: contributes 0
                    if (x <= 0.f) x = 0.f;
                    int x_low = (int)x, x_high;
                    if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
                    const float lx = x - x_low, hx = 1.f - lx;
                    const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                    if (act) {
                        float v1[8], v2[8], v3[8], v4[8];
                        Vec8<T>::load(base + row_lo + x_low * C, v1);
                        Vec8<T>::load(base + row_lo + x_high * C, v2);
                        Vec8<T>::load(base + row_hi + x_low * C, v3);
                        Vec8<T>::load(base + row_hi + x_high * C, v4);
                        // w1*v1 + w2*v2 + w3*v3 + w4*v4 contracted the way nvcc contracts torchvision's expression
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            acc[k] += __fmaf_rn(w4, v4[k], __fmaf_rn(w3, v3[k], __fmaf_rn(w2, v2[k], w1 * v1[k])));
                    }
                }
            }
            if (act) {
#pragma unroll
                for (int k = 0; k < 8; ++k) block[(size_t)(c0 + k) * bins + bin] = Vec8<T>::store(__fdividef_exact(acc[k], count));
            }
        }
    }
    __syncthreads();
    // contiguous (C * P * P) write-out; 16-byte vectors when the RoI's block is 16-byte aligned and sized
    T *dst = out + (size_t)r * C * bins;
    const size_t bytes = (size_t)C * bins * sizeof(T);
    if (bytes % 16 == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(block);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (int i = threadIdx.x; i < (int)(bytes / 16); i += kAlignThreads) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < C * bins; i += kAlignThreads) dst[i] = block[i];
    }
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" int msq_roi_align_levels(const void *const *feat_dev, const int *heights, const int *widths, const float *scales, int n_levels,
                                    int C, int is_bf16, const float *rois_dev, const long long *levels_dev, int n_rois, int P,
                                    int sampling_ratio, void *out_dev, void *stream) {
    MSQ_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels, MSQ_EINVAL, "msq_roi_align_levels: 1..%d pyramid levels (got %d)", kMaxLevels, n_levels);
    MSQ_REQUIRE(C > 0 && C % 8 == 0, MSQ_EUNSUPPORTED, "msq_roi_align_levels: channel count must be a multiple of 8 (got %d)", C);
    MSQ_REQUIRE(P >= 1 && sampling_ratio >= 1 && sampling_ratio <= kMaxSampling, MSQ_EUNSUPPORTED,
                "msq_roi_align_levels: output size must be positive and the sampling ratio in 1..4 (P=%d, sampling=%d)", P, sampling_ratio);
    MSQ_REQUIRE(n_rois >= 0, MSQ_EINVAL, "msq_roi_align_levels: n_rois=%d", n_rois);
    if (n_rois == 0) return MSQ_OK;
    MSQ_REQUIRE(feat_dev && heights && widths && scales && rois_dev && out_dev && (n_levels == 1 || levels_dev), MSQ_EINVAL,
                "msq_roi_align_levels: null pointer");
    PyramidArg pyr;
    pyr.n_levels = n_levels;
    for (int l = 0; l < n_levels; ++l) {
        MSQ_REQUIRE(feat_dev[l] && heights[l] > 0 && widths[l] > 0 && (uintptr_t)feat_dev[l] % 16 == 0, MSQ_EINVAL,
                    "msq_roi_align_levels: level %d: bad feature map", l);
        pyr.feat[l] = feat_dev[l]; pyr.H[l] = heights[l]; pyr.W[l] = widths[l]; pyr.scale[l] = scales[l];
    }
    const size_t smem = (size_t)C * P * P * (is_bf16 ? 2 : 4);
    MSQ_REQUIRE(smem <= 220 * 1024, MSQ_EUNSUPPORTED, "msq_roi_align_levels: C*P*P = %d*%d*%d needs %zu bytes of shared memory (limit 220 KB)", C, P, P, smem);
    MSQ_REQUIRE((uintptr_t)out_dev % 16 == 0, MSQ_EINVAL, "msq_roi_align_levels: output must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    // one instantiation per (dtype, sampling ratio); launched through a type-erased trampoline
    auto launch = [&](auto kernel, auto *typed_out) -> int {
        if (smem > 48 * 1024) MSQ_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TimedLaunch timed(K_DETECTOR_GLUE, st);
        kernel<<<n_rois, kAlignThreads, smem, st>>>(pyr, C, rois_dev, levels_dev, P, typed_out);
        return MSQ_OK;
    };
    int rc = MSQ_OK;
#define MSQ_ALIGN_CASE(S_)                                                                                                    \
    case S_:                                                                                                                  \
        rc = is_bf16 ? launch(roi_align_levels_kernel<__nv_bfloat16, S_>, static_cast<__nv_bfloat16 *>(out_dev))              \
                     : launch(roi_align_levels_kernel<float, S_>, static_cast<float *>(out_dev));                             \
        break;
    switch (sampling_ratio) {
        MSQ_ALIGN_CASE(1) MSQ_ALIGN_CASE(2) MSQ_ALIGN_CASE(3) MSQ_ALIGN_CASE(4)
    }
#undef MSQ_ALIGN_CASE
    if (rc != MSQ_OK) return rc;
    MSQ_LAUNCH_OK("roi_align_levels");
    return MSQ_OK;
}
