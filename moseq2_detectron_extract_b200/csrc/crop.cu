// a13 crop_and_rotate_frame: rotate about the centroid and crop to (crop_h, crop_w), bit-exact with
// cv2.warpAffine(8UC1, INTER_LINEAR, BORDER_CONSTANT 0).   ref proc/proc.py:305-335
//
// The reference pads the frame by the crop size, slices [int(c-crop/2), int(c+crop/2)) and warps the
// slice with getRotationMatrix2D((crop/2, crop/2), angle, 1).  OpenCV inverts the matrix in float64
// and evaluates it in fixed point: coordinates with 10 fractional bits, rounded to 5 interpolation
// bits, bilinear weights scaled to 2^15 (SURVEY.md section 7 trap 4; the test-side
// numpy restatement of the same arithmetic is pinned against cv2).  Here that is a gather straight
// from the un-padded frame: one CTA per frame warps both planes, one thread per output pixel.
#include "common.cuh"
#include <limits.h>
#include <math.h>

namespace msq {
namespace {

struct __align__(16) WarpCoeffs {
    double i00, i01, b1, i10, i11, b2;   // inverse affine map (dst -> src sub-image)
    int ox, oy, sw, sh;                  // sub-image origin in frame px and its size (sw == 0: zero crop)
};
static_assert(sizeof(WarpCoeffs) == 64, "one coefficient record per frame, 64 bytes");

__device__ void make_coeffs(double cx, double cy, double angle_deg, int cw, int ch, int W, int H, WarpCoeffs &k) {
    k.sw = k.sh = 0;
    k.ox = k.oy = 0;
    k.i00 = k.i01 = k.b1 = k.i10 = k.i11 = k.b2 = 0.0;
    if (angle_deg != angle_deg || cx != cx || cy != cy) return;     // NaN -> zeros (proc.py:317-318)
    if (cx < 0 || cy < 0) return;                                   // proc.py:320-322
    const int hx = cw / 2, hy = ch / 2;
    const int ox = (int)(cx - (double)hx);                           // Python int(): truncation
    const int oy = (int)(cy - (double)hy);
    int sw = (int)(cx + (double)hx) - ox;
    int sh = (int)(cy + (double)hy) - oy;
    sw = min(sw, W + cw - ox);                                       // NumPy slicing clips at the padded canvas
    sh = min(sh, H + ch - oy);
    if (sw <= 0 || sh <= 0) return;
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
    k.ox = ox; k.oy = oy; k.sw = sw; k.sh = sh;
}

__device__ __forceinline__ int sat_short(int v) { return max(-32768, min(32767, v)); }

// One CTA per frame: the float64 rotation coefficients, then -- like OpenCV's WarpAffineInvoker -- the fixed-point
// column terms adelta[x], bdelta[x] and row terms X0[y], Y0[y] (AB_BITS 10, round_delta 16), so that the per-pixel
// kernel is integer-only.  Table layout per frame: [adelta(cw) | bdelta(cw) | X0(ch) | Y0(ch)] int32.
__global__ void __launch_bounds__(128)
crop_coeffs_kernel(const double *__restrict__ centroid, const double *__restrict__ angle_deg, int n, int cw, int ch,
                   int W, int H, WarpCoeffs *__restrict__ coeffs, int *__restrict__ tables) {
    __shared__ WarpCoeffs k;
    const int f = blockIdx.x;
    if (threadIdx.x == 0) {
        make_coeffs(centroid[2 * f], centroid[2 * f + 1], angle_deg[f], cw, ch, W, H, k);
        coeffs[f] = k;
    }
    __syncthreads();
    if (k.sw == 0) return;
    int *t = tables + (size_t)f * 2 * (cw + ch);
    for (int x = threadIdx.x; x < cw; x += blockDim.x) {
        t[x] = __double2int_rn(k.i00 * (double)x * 1024.0);
        t[cw + x] = __double2int_rn(k.i10 * (double)x * 1024.0);
    }
    for (int y = threadIdx.x; y < ch; y += blockDim.x) {
        t[2 * cw + y] = __double2int_rn((k.i01 * (double)y + k.b1) * 1024.0) + 16;
        t[2 * cw + ch + y] = __double2int_rn((k.i11 * (double)y + k.b2) * 1024.0) + 16;
    }
}

constexpr int kCropThreads = 256;

// one thread per output pixel, both planes (frame + mask) share the transform.  Per pixel: OpenCV's
// fixed-point coordinates (AB_BITS 10, INTER_BITS 5, round_delta 16) and four gathered bytes per plane.
__global__ void __launch_bounds__(kCropThreads)
crop_rotate_kernel(const uint8_t *__restrict__ src0, const uint8_t *__restrict__ src1, int H, int W,
                   const WarpCoeffs *__restrict__ coeffs, const int *__restrict__ tables, int cw, int ch,
                   uint8_t *__restrict__ out0, uint8_t *__restrict__ out1) {
    // a CTA covers a 16x16 output tile; each warp an 8x4 patch, so that under rotation the 32 lanes of a
    // gather touch a compact source footprint (a 32x1 line would hit up to 32 different source rows)
    const int f = blockIdx.x;                                  // frames on grid.x (2^31 - 1), tiles on grid.y
    const int tiles_x = (cw + 15) >> 4;
    const int tile_y = blockIdx.y / tiles_x, tile_x = blockIdx.y - tile_y * tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x = (tile_x << 4) + ((warp & 1) << 3) + (lane & 7);
    const int y = (tile_y << 4) + ((warp >> 1) << 2) + (lane >> 3);
    if (x >= cw || y >= ch) return;
    const size_t crop_off = (size_t)f * cw * ch + (size_t)y * cw + x;
    const int4 geo = __ldg(reinterpret_cast<const int4 *>(&coeffs[f].ox));
    if (geo.z == 0) {
        out0[crop_off] = 0;
        if (out1) out1[crop_off] = 0;
        return;
    }
    const int *t = tables + (size_t)f * 2 * (cw + ch);
    const int adelta = __ldg(t + x), bdelta = __ldg(t + cw + x);
    const int X0 = __ldg(t + 2 * cw + y), Y0 = __ldg(t + 2 * cw + ch + y);
    const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
    const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
    const int fx = X & 31, fy = Y & 31;
    // a tap (px,py) of the sub-image is readable iff 0 <= px < sw and its frame pixel gx = px + ox lies in [0, W)
    // (same in y): one interval per axis, evaluated once per axis instead of once per tap
    const int lo_x = max(0, -geo.x), hi_x = min(geo.z, W - geo.x);
    const int lo_y = max(0, -geo.y), hi_y = min(geo.w, H - geo.y);
    const bool okx0 = sx >= lo_x && sx < hi_x, okx1 = sx + 1 >= lo_x && sx + 1 < hi_x;
    const bool oky0 = sy >= lo_y && sy < hi_y, oky1 = sy + 1 >= lo_y && sy + 1 < hi_y;
    // clamp the gather address into the frame so that every tap can be loaded unconditionally
    const int gx0 = min(max(sx + geo.x, 0), W - 1), gx1 = min(max(sx + 1 + geo.x, 0), W - 1);
    const int gy0 = min(max(sy + geo.y, 0), H - 1), gy1 = min(max(sy + 1 + geo.y, 0), H - 1);
    const int w00 = (okx0 && oky0) ? (32 - fx) * (32 - fy) : 0, w01 = (okx1 && oky0) ? fx * (32 - fy) : 0;
    const int w10 = (okx0 && oky1) ? (32 - fx) * fy : 0, w11 = (okx1 && oky1) ? fx * fy : 0;
    const int o00 = gy0 * W + gx0, o01 = gy0 * W + gx1, o10 = gy1 * W + gx0, o11 = gy1 * W + gx1;
    const uint8_t *p0 = src0 + (size_t)f * H * W;
    // OpenCV: (sum of w*32*p + 2^14) >> 15 with weights summing to 2^15  ==  (sum of w*p + 2^9) >> 10
    const int acc0 = (int)__ldg(p0 + o00) * w00 + (int)__ldg(p0 + o01) * w01 + (int)__ldg(p0 + o10) * w10 + (int)__ldg(p0 + o11) * w11;
    out0[crop_off] = (uint8_t)((acc0 + 512) >> 10);
    if (src1) {
        const uint8_t *p1 = src1 + (size_t)f * H * W;
        const int acc1 = (int)__ldg(p1 + o00) * w00 + (int)__ldg(p1 + o01) * w01 + (int)__ldg(p1 + o10) * w10 + (int)__ldg(p1 + o11) * w11;
        out1[crop_off] = (uint8_t)((acc1 + 512) >> 10);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Fast path: one CTA per frame, the source footprint of the crop staged in shared memory.
// The rotated crop reads a box of at most ~hypot(cw, ch) + 2 pixels a side around the centroid (26 KB for two planes of
// an 80x80 crop).  Loading that box with coalesced 32-bit words and gathering the four taps from shared memory replaces
// eight scattered one-byte global loads per output pixel (the old kernel was L1-tag bound), and zero-filling everything
// the reference cannot read (outside the frame, outside the [int(c-crop/2), int(c+crop/2)) slice) while staging removes
// every bounds test from the gather: an unreadable tap is a zero pixel instead of a zero weight -- the same product.
// Coefficients and OpenCV's fixed-point tables are computed by the CTA itself (no scratch traffic, one launch).
// ---------------------------------------------------------------------------------------------------------------
struct StagedGeom {
    int pitch_max, rows_max;      // staging capacity per plane (bytes per row, rows)
};

template <bool kTwo>        // kTwo: a second plane (the mask) shares the transform
__global__ void __launch_bounds__(kCropThreads, 5)
crop_rotate_staged_kernel(const uint8_t *__restrict__ src0, const uint8_t *__restrict__ src1, int H, int W,
                          const double *__restrict__ centroid, const double *__restrict__ angle_deg, int cw, int ch,
                          StagedGeom G, uint8_t *__restrict__ out0, uint8_t *__restrict__ out1) {
    extern __shared__ __align__(16) unsigned char crop_smem[];
    __shared__ WarpCoeffs k;
    __shared__ int box[4];                                       // gx_lo (4-aligned), gy_lo, pitch, rows; pitch 0 = global path
    int *adelta = reinterpret_cast<int *>(crop_smem), *bdelta = adelta + cw, *X0 = bdelta + cw, *Y0 = X0 + ch;
    uint8_t *plane0 = reinterpret_cast<uint8_t *>(Y0 + ch);
    uint8_t *plane1 = plane0 + (size_t)G.pitch_max * G.rows_max;
    const int f = blockIdx.x;
    const size_t crop_base = (size_t)f * cw * ch;
    if (threadIdx.x == 0) make_coeffs(centroid[2 * f], centroid[2 * f + 1], angle_deg[f], cw, ch, W, H, k);
    __syncthreads();
    if (k.sw == 0) {                                             // NaN / negative centre / empty slice: a zero crop
        for (int i = threadIdx.x; i < (cw * ch) / 4; i += kCropThreads) {
            reinterpret_cast<uint32_t *>(out0 + crop_base)[i] = 0u;
            if (out1) reinterpret_cast<uint32_t *>(out1 + crop_base)[i] = 0u;
        }
        return;
    }
    for (int x = threadIdx.x; x < cw; x += kCropThreads) {
        adelta[x] = __double2int_rn(k.i00 * (double)x * 1024.0);
        bdelta[x] = __double2int_rn(k.i10 * (double)x * 1024.0);
    }
    for (int y = threadIdx.x; y < ch; y += kCropThreads) {
        X0[y] = __double2int_rn((k.i01 * (double)y + k.b1) * 1024.0) + 16;
        Y0[y] = __double2int_rn((k.i11 * (double)y + k.b2) * 1024.0) + 16;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // adelta/bdelta are monotone in x and X0/Y0 in y (rounded linear maps): the extreme taps are at the corners
        int x_lo = INT_MAX, x_hi = INT_MIN, y_lo = INT_MAX, y_hi = INT_MIN;
        for (int c = 0; c < 4; ++c) {
            const int x = (c & 1) ? cw - 1 : 0, y = (c & 2) ? ch - 1 : 0;
            const int sx = (X0[y] + adelta[x]) >> 10, sy = (Y0[y] + bdelta[x]) >> 10;
            x_lo = min(x_lo, sx); x_hi = max(x_hi, sx + 1);
            y_lo = min(y_lo, sy); y_hi = max(y_hi, sy + 1);
        }
        const int gx_lo = (x_lo + k.ox) & ~3, gx_hi = x_hi + k.ox;   // frame coordinates, left edge on a 32-bit word
        // an odd number of 32-bit words per staged row: lanes that walk down a column of the box (crops rotated by ~90 degrees)
        // then fall on 32 different banks instead of 16
        const int pitch = ((((gx_hi - gx_lo + 1) + 3) >> 2) | 1) << 2, rows = y_hi - y_lo + 1;
        const bool fits = pitch <= G.pitch_max && rows <= G.rows_max && x_lo > -32000 && x_hi < 32000 && y_lo > -32000 && y_hi < 32000;
        box[0] = gx_lo; box[1] = y_lo + k.oy; box[2] = fits ? pitch : 0; box[3] = rows;
    }
    __syncthreads();
    const int gx_lo = box[0], gy_lo = box[1], pitch = box[2], rows = box[3];
    if (pitch == 0) {
        // not expected for a rotation about the crop centre; kept for exactness: gather from global memory per tap
        const int lo_x = max(0, -k.ox), hi_x = min(k.sw, W - k.ox), lo_y = max(0, -k.oy), hi_y = min(k.sh, H - k.oy);
        for (int i = threadIdx.x; i < cw * ch; i += kCropThreads) {
            const int y = i / cw, x = i - y * cw;
            const int X = (X0[y] + adelta[x]) >> 5, Y = (Y0[y] + bdelta[x]) >> 5;
            const int sx = sat_short(X >> 5), sy = sat_short(Y >> 5), fx = X & 31, fy = Y & 31;
            int acc0 = 0, acc1 = 0;
            for (int t = 0; t < 4; ++t) {
                const int px = sx + (t & 1), py = sy + (t >> 1);
                if (px < lo_x || px >= hi_x || py < lo_y || py >= hi_y) continue;
                const int wgt = ((t & 1) ? fx : 32 - fx) * ((t >> 1) ? fy : 32 - fy);
                const size_t o = (size_t)f * H * W + (size_t)(py + k.oy) * W + px + k.ox;
                acc0 += (int)__ldg(src0 + o) * wgt;
                if (src1) acc1 += (int)__ldg(src1 + o) * wgt;
            }
            out0[crop_base + i] = (uint8_t)((acc0 + 512) >> 10);
            if (out1) out1[crop_base + i] = (uint8_t)((acc1 + 512) >> 10);
        }
        return;
    }
    // ---- stage: frame pixels readable by the reference (inside the frame AND inside its slice), zero elsewhere ----
    const int rx0 = max(0, k.ox), rx1 = min(k.ox + k.sw, W), ry0 = max(0, k.oy), ry1 = min(k.oy + k.sh, H);
    const int words = pitch >> 2;
    const uint8_t *f0 = src0 + (size_t)f * H * W;
    const uint8_t *f1 = src1 ? src1 + (size_t)f * H * W : nullptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = kCropThreads / 32;
    // a warp stages whole rows, lane = 32-bit word.  Which bytes of a lane's word are readable depends on the column only, so the
    // byte mask and the lane's pointer are set up once per column block; a row then costs one warp-uniform test, one offset and
    // a predicated load per plane (the first version redid the 64-bit addressing and the byte masks for every word: 39 % of the
    // kernel's instructions, ncu source view).
    uint32_t *s0 = reinterpret_cast<uint32_t *>(plane0), *s1 = reinterpret_cast<uint32_t *>(plane1);
    for (int cb = 0; cb < words; cb += 32) {
        const int c = cb + lane;
        const int gx = gx_lo + 4 * c;
        uint32_t keep = 0u;                                          // bytes gx..gx+3 (little endian) the reference can read
        if (c < words && gx + 3 >= rx0 && gx < rx1) {
            keep = 0xffffffffu;
            if (gx < rx0) keep &= 0xffffffffu << (8 * (rx0 - gx));
            if (gx + 4 > rx1) keep &= 0xffffffffu >> (8 * (gx + 4 - rx1));
        }
        const uint8_t *p0 = f0 + gx, *p1 = f1 ? f1 + gx : nullptr;  // dereferenced only where keep != 0
        // four rows per trip: all their loads are in flight before the first store needs its value (a warp has ~15 rows; one row
        // per trip exposed one global-memory latency per row)
        constexpr int kRowsPerTrip = 8;
        for (int r0 = warp; r0 < rows; r0 += kWarps * kRowsPerTrip) {
            uint32_t v0[kRowsPerTrip], v1[kRowsPerTrip];
#pragma unroll
            for (int q = 0; q < kRowsPerTrip; ++q) {
                const int r = r0 + q * kWarps, gy = gy_lo + r;
                v0[q] = v1[q] = 0u;
                if (keep != 0u && r < rows && gy >= ry0 && gy < ry1) {
                    const int off = gy * W;
                    v0[q] = __ldg(reinterpret_cast<const uint32_t *>(p0 + off));
                    if (kTwo) v1[q] = __ldg(reinterpret_cast<const uint32_t *>(p1 + off));
                }
            }
#pragma unroll
            for (int q = 0; q < kRowsPerTrip; ++q) {
                const int r = r0 + q * kWarps;
                if (c < words && r < rows) {
                    s0[r * words + c] = v0[q] & keep;
                    if (kTwo) s1[r * words + c] = v1[q] & keep;
                }
            }
        }
    }
    __syncthreads();
    // ---- gather: neighbouring lanes take neighbouring output pixels of a row.  (4 pixels per lane with one 32-bit store was
    // measured first: lanes 4 px apart step through the staged box in multiples of 4 words along a rotated line, which folds the
    // 32 lanes onto 8 banks -- 60 % of the shared-memory wavefronts conflicted.)  A thread keeps its column: with cw <= 256 the
    // CTA is kCropThreads / cw row groups of cw threads (80-px crops: 3 x 80 threads, 27 rows each, instead of 32-lane column
    // strips whose third strip is half empty), the column terms stay in registers and the output pointers advance by a constant.
    const int sh_x = k.ox - gx_lo, sh_y = k.oy - gy_lo;
    const int groups = cw <= kCropThreads ? kCropThreads / cw : 0;
    if (groups > 0) {
        const int t = threadIdx.x, g = t / cw, x = t - g * cw;
        if (g >= groups) return;
        const int ad = adelta[x], bd = bdelta[x];
        const size_t step = (size_t)groups * cw;
        uint8_t *o0 = out0 + crop_base + (size_t)g * cw + x;
        uint8_t *o1 = kTwo ? out1 + crop_base + (size_t)g * cw + x : nullptr;
        for (int y = g; y < ch; y += groups, o0 += step, o1 += step) {
            const int X = (X0[y] + ad) >> 5, Y = (Y0[y] + bd) >> 5;
            const int fx = X & 31, fy = Y & 31, gx = 32 - fx, gy = 32 - fy;
            const int idx = ((Y >> 5) + sh_y) * pitch + (X >> 5) + sh_x;
            {
                const int top = (int)plane0[idx] * gx + (int)plane0[idx + 1] * fx;
                const int bot = (int)plane0[idx + pitch] * gx + (int)plane0[idx + pitch + 1] * fx;
                *o0 = (uint8_t)((top * gy + bot * fy + 512) >> 10);
            }
            if (kTwo) {
                const int top = (int)plane1[idx] * gx + (int)plane1[idx + 1] * fx;
                const int bot = (int)plane1[idx + pitch] * gx + (int)plane1[idx + pitch + 1] * fx;
                *o1 = (uint8_t)((top * gy + bot * fy + 512) >> 10);
            }
        }
        return;
    }
    for (int x = lane; x < cw; x += 32) {                       // crops wider than the CTA: column strips
        const int ad = adelta[x], bd = bdelta[x];
        for (int y = warp; y < ch; y += kWarps) {
            const int X = (X0[y] + ad) >> 5, Y = (Y0[y] + bd) >> 5;
            const int fx = X & 31, fy = Y & 31;
            const int idx = ((Y >> 5) + sh_y) * pitch + (X >> 5) + sh_x;
            const size_t o = crop_base + (size_t)y * cw + x;
            {
                const int top = (int)plane0[idx] * (32 - fx) + (int)plane0[idx + 1] * fx;
                const int bot = (int)plane0[idx + pitch] * (32 - fx) + (int)plane0[idx + pitch + 1] * fx;
                out0[o] = (uint8_t)((top * (32 - fy) + bot * fy + 512) >> 10);
            }
            if (kTwo) {
                const int top = (int)plane1[idx] * (32 - fx) + (int)plane1[idx + 1] * fx;
                const int bot = (int)plane1[idx + pitch] * (32 - fx) + (int)plane1[idx + pitch + 1] * fx;
                out1[o] = (uint8_t)((top * (32 - fy) + bot * fy + 512) >> 10);
            }
        }
    }
}

}  // namespace

int launch_crop_rotate(const uint8_t *src, const uint8_t *src2, int n, int h, int w, const double *centroid,
                       const double *angle_deg, int cw, int ch, uint8_t *out, uint8_t *out2, void *scratch,
                       cudaStream_t st) {
    const bool two = src2 && out2;
    // fast path: 32-bit staging loads and stores need 4-byte friendly shapes and bases
    const int side = (int)ceil(hypot((double)cw, (double)ch)) + 4;
    StagedGeom G;
    G.pitch_max = ((side + 4 + 3) & ~3) + 4;
    G.rows_max = side;
    const size_t smem = (size_t)2 * (cw + ch) * sizeof(int) + (size_t)(two ? 2 : 1) * G.pitch_max * G.rows_max;
    const bool aligned = (w % 4 == 0) && (cw % 4 == 0) && ((uintptr_t)src % 4 == 0) && ((uintptr_t)out % 4 == 0) &&
                         (!two || (((uintptr_t)src2 % 4 == 0) && ((uintptr_t)out2 % 4 == 0))) && (((size_t)h * w) % 4 == 0);
    if (aligned && smem <= 160 * 1024) {
        auto kernel = two ? crop_rotate_staged_kernel<true> : crop_rotate_staged_kernel<false>;
        if (smem > 48 * 1024) MSQ_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TimedLaunch timed(K_CROP, st);
        kernel<<<n, kCropThreads, smem, st>>>(src, two ? src2 : nullptr, h, w, centroid, angle_deg, cw, ch, G, out, two ? out2 : nullptr);
        MSQ_LAUNCH_OK("crop_rotate (staged)");
        return MSQ_OK;
    }
    WarpCoeffs *coeffs = reinterpret_cast<WarpCoeffs *>(scratch);
    int *tables = reinterpret_cast<int *>(reinterpret_cast<char *>(scratch) + (size_t)n * sizeof(WarpCoeffs));
    TimedLaunch timed(K_CROP, st, 2);       // the coefficient kernel is part of the crop step
    crop_coeffs_kernel<<<n, 128, 0, st>>>(centroid, angle_deg, n, cw, ch, w, h, coeffs, tables);
    MSQ_LAUNCH_OK("crop_coeffs");
    dim3 grid(n, ((cw + 15) / 16) * ((ch + 15) / 16));
    crop_rotate_kernel<<<grid, kCropThreads, 0, st>>>(src, two ? src2 : nullptr, h, w, coeffs, tables, cw, ch, out,
                                                      two ? out2 : nullptr);
    MSQ_LAUNCH_OK("crop_rotate");
    return MSQ_OK;
}

}  // namespace msq

// per frame: one 64-byte coefficient record + the 2*(cw+ch) int32 fixed-point tables (sized for crops up to 512x512)
extern "C" size_t msq_crop_scratch_bytes(int n) { return (size_t)(n > 0 ? n : 0) * (64 + 2 * (512 + 512) * sizeof(int)); }

extern "C" int msq_crop_rotate(const uint8_t *src, const uint8_t *src2, int n, int h, int w, const double *centroid,
                               const double *angle_deg, int cw, int ch, uint8_t *out, uint8_t *out2, void *scratch,
                               size_t scratch_bytes, void *stream) {
    MSQ_REQUIRE(n == 0 || (src && out && centroid && angle_deg), MSQ_EINVAL, "msq_crop_rotate: null pointer");
    MSQ_REQUIRE((src2 == nullptr) == (out2 == nullptr), MSQ_EINVAL, "msq_crop_rotate: src2/out2 must both be set or both be null");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0 && cw > 0 && ch > 0, MSQ_EINVAL, "msq_crop_rotate: bad sizes");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(cw <= 512 && ch <= 512, MSQ_EUNSUPPORTED, "msq_crop_rotate: crop %dx%d exceeds 512x512", cw, ch);
    MSQ_REQUIRE(scratch && (uintptr_t)scratch % 16 == 0 && scratch_bytes >= msq_crop_scratch_bytes(n), MSQ_ENOMEM,
                "msq_crop_rotate: scratch must be 16-byte aligned and >= %zu bytes", msq_crop_scratch_bytes(n));
    return msq::launch_crop_rotate(src, src2, n, h, w, centroid, angle_deg, cw, ch, out, out2, scratch, (cudaStream_t)stream);
}
