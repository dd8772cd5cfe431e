// a13 crop_and_rotate_frame: rotate about the centroid and crop to (crop_h, crop_w), bit-exact with
// cv2.warpAffine(8UC1, INTER_LINEAR, BORDER_CONSTANT 0).   ref proc/proc.py:305-335
//
// The reference pads the frame by the crop size, slices [int(c-crop/2), int(c+crop/2)) and warps the
// slice with getRotationMatrix2D((crop/2, crop/2), angle, 1).  OpenCV inverts the matrix in float64
// and evaluates it in fixed point: coordinates with 10 fractional bits, rounded to 5 interpolation
// bits, bilinear weights scaled to 2^15 (SURVEY.md section 7 trap 4; oracle/extract_oracle.py
// crop_rotate_np is the same arithmetic and is pinned against cv2).  Here that is a gather straight
// from the un-padded frame: one CTA per (frame, plane), one thread per output pixel.
#include "common.cuh"
#include <math.h>

namespace msq {
namespace {

struct WarpCoeffs {
    double i00, i01, b1, i10, i11, b2;   // inverse affine map (dst -> src sub-image)
    int ox, oy, sw, sh;                  // sub-image origin in frame px and its size
    int valid;
};

__device__ void make_coeffs(double cx, double cy, double angle_deg, int cw, int ch, int W, int H, WarpCoeffs &k) {
    k.valid = 0;
    if (angle_deg != angle_deg || cx != cx || cy != cy) return;     // NaN -> zeros (proc.py:317-318)
    if (cx < 0 || cy < 0) return;                                   // proc.py:320-322
    const int hx = cw / 2, hy = ch / 2;
    k.ox = (int)(cx - (double)hx);                                   // Python int(): truncation
    k.oy = (int)(cy - (double)hy);
    k.sw = (int)(cx + (double)hx) - k.ox;
    k.sh = (int)(cy + (double)hy) - k.oy;
    k.sw = min(k.sw, W + cw - k.ox);                                 // NumPy slicing clips at the padded canvas
    k.sh = min(k.sh, H + ch - k.oy);
    if (k.sw <= 0 || k.sh <= 0) return;
    // cv::getRotationMatrix2D(center=(hx,hy), angle, 1)
    const double rad = angle_deg * 0.017453292519943295;            // angle *= CV_PI/180
    const double alpha = cos(rad), beta = sin(rad);
    const double m00 = alpha, m01 = beta, m02 = (1 - alpha) * (double)hx - beta * (double)hy;
    const double m10 = -beta, m11 = alpha, m12 = beta * (double)hx + (1 - alpha) * (double)hy;
    // cv::warpAffine without WARP_INVERSE_MAP inverts M
    double det = m00 * m11 - m01 * m10;
    det = det != 0 ? 1.0 / det : 0.0;
    const double a11 = m11 * det, a22 = m00 * det;
    k.i00 = a11;
    k.i01 = m01 * (-det);
    k.i10 = m10 * (-det);
    k.i11 = a22;
    k.b1 = -k.i00 * m02 - k.i01 * m12;
    k.b2 = -k.i10 * m02 - k.i11 * m12;
    k.valid = 1;
}

__device__ __forceinline__ int sat_short(int v) { return max(-32768, min(32767, v)); }

__global__ void __launch_bounds__(256)
crop_rotate_kernel(const uint8_t *__restrict__ src0, const uint8_t *__restrict__ src1, int n, int H, int W,
                   const double *__restrict__ centroid, const double *__restrict__ angle_deg, int cw, int ch,
                   uint8_t *__restrict__ out0, uint8_t *__restrict__ out1) {
    __shared__ WarpCoeffs k;
    const int f = blockIdx.x;
    const uint8_t *src = (blockIdx.y == 0 ? src0 : src1) + (size_t)f * H * W;
    uint8_t *dst = (blockIdx.y == 0 ? out0 : out1) + (size_t)f * cw * ch;
    if (threadIdx.x == 0) make_coeffs(centroid[2 * f], centroid[2 * f + 1], angle_deg[f], cw, ch, W, H, k);
    __syncthreads();
    const int total = cw * ch;
    if (!k.valid) {
        for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = 0;
        return;
    }
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int y = i / cw, x = i - y * cw;
        // AB_BITS = 10, INTER_BITS = 5, round_delta = 16
        const int adelta = __double2int_rn(k.i00 * (double)x * 1024.0);
        const int bdelta = __double2int_rn(k.i10 * (double)x * 1024.0);
        const int X0 = __double2int_rn((k.i01 * (double)y + k.b1) * 1024.0) + 16;
        const int Y0 = __double2int_rn((k.i11 * (double)y + k.b2) * 1024.0) + 16;
        const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
        const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
        const int fx = X & 31, fy = Y & 31;
        int acc = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int dx = t & 1, dy = t >> 1;
            const int px = sx + dx, py = sy + dy;
            const int wgt = (dx ? fx : 32 - fx) * (dy ? fy : 32 - fy) * 32;     // sums to 2^15
            int v = 0;
            if (px >= 0 && px < k.sw && py >= 0 && py < k.sh) {
                const int gx = px + k.ox, gy = py + k.oy;
                if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = __ldg(src + (size_t)gy * W + gx);
            }
            acc += v * wgt;
        }
        dst[i] = (uint8_t)((acc + 16384) >> 15);
    }
}

}  // namespace

int launch_crop_rotate(const uint8_t *src, const uint8_t *src2, int n, int h, int w, const double *centroid,
                       const double *angle_deg, int cw, int ch, uint8_t *out, uint8_t *out2, cudaStream_t st) {
    dim3 grid(n, (src2 && out2) ? 2 : 1);
    TimedLaunch timed(K_CROP, st);
    crop_rotate_kernel<<<grid, 256, 0, st>>>(src, src2, n, h, w, centroid, angle_deg, cw, ch, out, out2);
    MSQ_LAUNCH_OK("crop_rotate");
    return MSQ_OK;
}

}  // namespace msq

extern "C" int msq_crop_rotate(const uint8_t *src, const uint8_t *src2, int n, int h, int w, const double *centroid,
                               const double *angle_deg, int cw, int ch, uint8_t *out, uint8_t *out2, void *stream) {
    MSQ_REQUIRE(src && out && centroid && angle_deg, MSQ_EINVAL, "msq_crop_rotate: null pointer");
    MSQ_REQUIRE((src2 == nullptr) == (out2 == nullptr), MSQ_EINVAL, "msq_crop_rotate: src2/out2 must both be set or both be null");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0 && cw > 0 && ch > 0, MSQ_EINVAL, "msq_crop_rotate: bad sizes");
    if (n == 0) return MSQ_OK;
    return msq::launch_crop_rotate(src, src2, n, h, w, centroid, angle_deg, cw, ch, out, out2, (cudaStream_t)stream);
}
