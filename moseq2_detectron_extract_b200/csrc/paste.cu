// a4 mask paste: detectron2's paste_masks_in_image as used by detector_postprocess
// (ref model/util.py:45-62 -> detectron2/layers/mask_ops.py::_do_paste_mask; third-party, restated).
// Each output pixel centre (x+0.5, y+0.5) is mapped into the box, normalised to [-1,1], and the MxM
// soft mask is sampled like F.grid_sample(align_corners=False, padding zeros); >= threshold.
// One CTA per (frame, row band); the soft mask (28x28 floats) is staged in shared memory and the
// mask bytes leave as 32-bit words.  float32 arithmetic in torch's operation order, no contraction.
#include "common.cuh"
#include <math.h>

namespace msq {
namespace {

constexpr int kPasteThreads = 256;
constexpr int kPasteRows = 16;

__device__ __forceinline__ float unnormalize(float g, int M) {
    // grid_sampler_unnormalize, align_corners=False: ((g + 1) * size - 1) / 2
    return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), (float)M), 1.0f), 2.0f);
}

__global__ void __launch_bounds__(kPasteThreads)
paste_kernel(const float *__restrict__ soft, const float *__restrict__ boxes, int n, int M, int h, int w,
             float threshold, uint8_t *__restrict__ out) {
    extern __shared__ float tile[];                       // M*M
    const int f = blockIdx.y;
    const int row0 = blockIdx.x * kPasteRows;
    for (int i = threadIdx.x; i < M * M; i += kPasteThreads) tile[i] = soft[(size_t)f * M * M + i];
    __syncthreads();
    const float x0 = boxes[4 * f], y0 = boxes[4 * f + 1], x1 = boxes[4 * f + 2], y1 = boxes[4 * f + 3];
    const float bw = __fsub_rn(x1, x0), bh = __fsub_rn(y1, y0);
    const int rows = min(kPasteRows, h - row0);
    uint8_t *dst = out + (size_t)f * h * w;
    for (int i = threadIdx.x; i < rows * w; i += kPasteThreads) {
        const int y = row0 + i / w, x = i % w;
        // img = (arange + 0.5 - x0) / (x1 - x0) * 2 - 1
        const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__fadd_rn((float)y, 0.5f), y0), bh), 2.0f), 1.0f);
        const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(__fadd_rn((float)x, 0.5f), x0), bw), 2.0f), 1.0f);
        const float py = unnormalize(gy, M), px = unnormalize(gx, M);
        const float fy = floorf(py), fx = floorf(px);
        const int iy = (int)fy, ix = (int)fx;
        const float wy1 = __fsub_rn(py, fy), wx1 = __fsub_rn(px, fx);
        const float wy0 = __fsub_rn(__fadd_rn(fy, 1.0f), py), wx0 = __fsub_rn(__fadd_rn(fx, 1.0f), px);   // torch: (i_se - i)
        float v = 0.0f;
        // torch accumulates nw, ne, sw, se in this order
        if (iy >= 0 && iy < M && ix >= 0 && ix < M) v = __fadd_rn(v, __fmul_rn(tile[iy * M + ix], __fmul_rn(wx0, wy0)));
        if (iy >= 0 && iy < M && ix + 1 >= 0 && ix + 1 < M) v = __fadd_rn(v, __fmul_rn(tile[iy * M + ix + 1], __fmul_rn(wx1, wy0)));
        if (iy + 1 >= 0 && iy + 1 < M && ix >= 0 && ix < M) v = __fadd_rn(v, __fmul_rn(tile[(iy + 1) * M + ix], __fmul_rn(wx0, wy1)));
        if (iy + 1 >= 0 && iy + 1 < M && ix + 1 >= 0 && ix + 1 < M) v = __fadd_rn(v, __fmul_rn(tile[(iy + 1) * M + ix + 1], __fmul_rn(wx1, wy1)));
        dst[(size_t)y * w + x] = (v >= threshold) ? 1 : 0;
    }
}

}  // namespace
}  // namespace msq

extern "C" int msq_paste_masks(const float *soft, const float *boxes, int n, int M, int h, int w, float threshold,
                               uint8_t *out, void *stream) {
    MSQ_REQUIRE(soft && boxes && out, MSQ_EINVAL, "msq_paste_masks: null pointer");
    MSQ_REQUIRE(n >= 0 && M > 0 && M <= 96 && h > 0 && w > 0, MSQ_EINVAL, "msq_paste_masks: bad sizes n=%d M=%d h=%d w=%d", n, M, h, w);
    if (n == 0) return MSQ_OK;
    dim3 grid((h + msq::kPasteRows - 1) / msq::kPasteRows, n);
    msq::TimedLaunch timed(msq::K_PASTE, (cudaStream_t)stream);
    msq::paste_kernel<<<grid, msq::kPasteThreads, (size_t)M * M * sizeof(float), (cudaStream_t)stream>>>(
        soft, boxes, n, M, h, w, threshold, out);
    MSQ_LAUNCH_OK("paste_masks");
    return MSQ_OK;
}
