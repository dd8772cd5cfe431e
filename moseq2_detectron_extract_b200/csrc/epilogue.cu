// Per-frame / per-chunk post-processing after the moment features:
//   a8-a10  degrees + clamp, keypoint flip votes, chunk-local iterative angle filter
//           ref proc/proc.py:720-724, 827-839, 851-889, 600-654
//   a12     compute_scalars         ref proc/scalars.py:36-120
//   a11     keypoints_to_dict       ref proc/keypoints.py:93-165, proc/util.py:29-61
// All floating point is float64 in NumPy's operation order (compiled with --fmad=false).
#include "common.cuh"
#include "angles.cuh"
#include <algorithm>
#include <math.h>

namespace msq {
namespace {

// ---------------------------------------------------------------------------------------------
// iterative_filter_angles on one chunk held in shared memory (ref proc/proc.py:600-654).
// All threads of the CTA call this; returns the buffer holding the final angles.
// ---------------------------------------------------------------------------------------------
constexpr int kAngleThreads = 1024;

__device__ double *filter_angles_in_smem(double *buf0, double *buf1, int len, int window, double tolerance,
                                         int max_iters, int *passes_out) {
    double *last = buf0, *cur = buf1;
    const int win = min(window, len);
    const double lo_thr = 180.0 - tolerance, hi_thr = 180.0 + tolerance;
    int passes = 0;
    for (;;) {
        ++passes;
        int not_close = 0;
        for (int i = threadIdx.x; i < len; i += kAngleThreads) {
            const double v0 = last[i];
            double med;
            if (win == 3) {
                // trailing median of the non-NaN values among last[i-2..i] (bottleneck.move_median, min_count=1)
                double w[3];
                int cnt = 0;
                if (v0 == v0) w[cnt++] = v0;
                if (i >= 1) { const double v1 = last[i - 1]; if (v1 == v1) w[cnt++] = v1; }
                if (i >= 2) { const double v2 = last[i - 2]; if (v2 == v2) w[cnt++] = v2; }
                if (cnt == 0) med = nan_f64();
                else if (cnt == 1) med = w[0];
                else if (cnt == 2) med = (w[0] + w[1]) / 2;
                else med = fmax(fmin(w[0], w[1]), fmin(fmax(w[0], w[1]), w[2]));
            } else {
                // general window: selection by insertion into a small sorted list (window <= 15)
                double w[15];
                int cnt = 0;
                for (int k = 0; k < win && i - k >= 0; ++k) {
                    const double v = last[i - k];
                    if (v != v) continue;
                    int j = cnt++;
                    while (j > 0 && w[j - 1] > v) { w[j] = w[j - 1]; --j; }
                    w[j] = v;
                }
                if (cnt == 0) med = nan_f64();
                else med = (cnt & 1) ? w[cnt / 2] : (w[cnt / 2 - 1] + w[cnt / 2]) / 2;
            }
            const double d = v0 - med;
            const double ad = fabs(d);
            double o = v0;
            if (ad > lo_thr && ad < hi_thr) o = v0 + (-180.0 * (d > 0 ? 1.0 : (d < 0 ? -1.0 : 0.0)));
            cur[i] = o;
            // np.allclose(cur, last): |cur-last| <= 1e-8 + 1e-5*|last|  (NaN is never close)
            if (!(fabs(o - v0) <= 1e-8 + 1e-5 * fabs(v0))) not_close = 1;
        }
        const int any_far = __syncthreads_or(not_close);
        // ref proc/proc.py:641-652: stop when converged, or after max_iters + 1 passes
        if (!any_far || passes > max_iters) break;
        double *tmp = last; last = cur; cur = tmp;
    }
    if (passes_out) *passes_out = passes;
    return cur;
}

__device__ __forceinline__ bool is_filter_flip(double after, double before) {
    return fabs(fabs(after - before) - 180.0) <= 1e-8 + 1e-5 * 180.0;    // np.isclose(|cur-angles|, 180)
}

// degrees + clamp (a8) and keypoint flips (a9) are per-frame float64 work (a dozen double-precision sin / cos per frame): one
// thread per frame over as many SMs as it takes -- inside the per-chunk filter kernel below the same work sat on the FP64
// units of ONE SM per chunk (38 us for six 1000-frame chunks, on the critical path of the chunk call).
__global__ void __launch_bounds__(128)
angle_votes_kernel(const double *__restrict__ orientation, const double *__restrict__ axis, const double *__restrict__ centroid,
                   const float *__restrict__ kpts, int n, double *__restrict__ angle_out, uint8_t *__restrict__ flips_out,
                   double *__restrict__ conf_out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const double length = np_max2(axis[2 * f], axis[2 * f + 1]);
    double a = -(orientation[f] * k180OverPi);               // ref proc/proc.py:722-724
    a = (a < 0) ? 360 + a : a;
    a = fmod(a, 360.0);                                      // a >= 0 here, same as numpy's floor-mod
    double conf;
    const bool flip = keypoint_flip_vote(kpts + (size_t)f * 24, centroid[2 * f], centroid[2 * f + 1], a, length, &conf);
    if (conf_out) conf_out[f] = conf;
    flips_out[f] = flip ? 1 : 0;
    if (flip) a += 180;                                      // ref proc/proc.py:834 (no clamp)
    angle_out[f] = a;                                        // pre-filter angle, needed for the isclose test
}

// one CTA per chunk: the iterative filter (a10) on the angles angle_votes_kernel left in angle_out
__global__ void __launch_bounds__(kAngleThreads)
angles_flips_kernel(int n, int chunk, double *__restrict__ angle_out, uint8_t *__restrict__ flips_out, int *__restrict__ passes_out) {
    extern __shared__ __align__(16) double sm_angles[];       // [2][len] ping-pong
    const int begin = blockIdx.x * chunk;
    const int len = min(chunk, n - begin);
    double *buf0 = sm_angles, *buf1 = sm_angles + len;
    for (int i = threadIdx.x; i < len; i += kAngleThreads) buf0[i] = angle_out[begin + i];
    __syncthreads();
    int passes;
    const double *cur = filter_angles_in_smem(buf0, buf1, len, 3, 60.0, 1000, &passes);
    for (int i = threadIdx.x; i < len; i += kAngleThreads) {
        const int f = begin + i;
        const double after = cur[i];
        flips_out[f] = (uint8_t)((flips_out[f] != 0) != is_filter_flip(after, angle_out[f]));   // ref proc/proc.py:839
        angle_out[f] = after;
    }
    if (passes_out && threadIdx.x == 0) passes_out[blockIdx.x] = passes;
}

template <class KP>
__global__ void __launch_bounds__(128)
flips_kernel(const KP *__restrict__ kpts, const double *__restrict__ centroid, const double *__restrict__ angles,
             const double *__restrict__ lengths, int n, uint8_t *__restrict__ flips, double *__restrict__ conf) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    double c;
    const bool flip = keypoint_flip_vote(kpts + (size_t)f * 24, centroid[2 * f], centroid[2 * f + 1], angles[f], lengths[f], &c);
    flips[f] = flip ? 1 : 0;
    if (conf) conf[f] = c;
}

__global__ void __launch_bounds__(kAngleThreads)
filter_kernel(const double *__restrict__ angles, int n, int chunk, int window, double tolerance, int max_iters,
              double *__restrict__ out, uint8_t *__restrict__ flips, int *__restrict__ passes_out) {
    extern __shared__ __align__(16) double sm_angles[];
    const int begin = blockIdx.x * chunk;
    const int len = min(chunk, n - begin);
    double *buf0 = sm_angles, *buf1 = sm_angles + len;
    for (int i = threadIdx.x; i < len; i += kAngleThreads) buf0[i] = angles[begin + i];
    __syncthreads();
    int passes;
    const double *cur = filter_angles_in_smem(buf0, buf1, len, window, tolerance, max_iters, &passes);
    for (int i = threadIdx.x; i < len; i += kAngleThreads) {
        const double before = angles[begin + i], after = cur[i];
        if (flips) flips[begin + i] = is_filter_flip(after, before) ? 1 : 0;
        out[begin + i] = after;
    }
    if (passes_out && threadIdx.x == 0) passes_out[blockIdx.x] = passes;
}

// ---------------------------------------------------------------------------------------------
// per-frame masked area / height sums of chunk*mask (one CTA per frame, 128-bit loads)
// ---------------------------------------------------------------------------------------------
constexpr int kSumThreads = 256;

__global__ void __launch_bounds__(kSumThreads)
masked_sums_kernel(const uint8_t *__restrict__ chunk, const uint8_t *__restrict__ mask, int n, size_t plane,
                   double min_h, double max_h, int vec_ok, int2 *__restrict__ out) {
    __shared__ int s_cnt[kSumThreads / 32], s_sum[kSumThreads / 32];
    // value passes if min_h < v < max_h; on integers: v in [lo, hi]
    const int lo = (int)floor(min_h) + 1, hi = (int)ceil(max_h) - 1;
    const ByteTest t_lo = make_byte_test(lo), t_hi1 = make_byte_test(hi + 1);      // value >= lo, value >= hi + 1
    for (int f = blockIdx.x; f < n; f += gridDim.x) {
        const uint8_t *c = chunk + (size_t)f * plane, *m = mask ? mask + (size_t)f * plane : nullptr;
        int cnt = 0, sum = 0;
        size_t done = 0;
        if (vec_ok) {
            const size_t nvec = plane / 16;
            // SWAR over 4 pixels per word: value = chunk * mask (mod 256), in range iff lo <= value <= hi, count by popcount,
            // sum by a 4-way byte dot product.  (ncu on the per-pixel form: ALU pipe 80 % busy, DRAM 54 %.)
            // A masked-out pixel has value 0, which is out of range whenever lo > 0 (min_height >= 0): 16 pixels without a
            // mask byte contribute nothing and their frame bytes are never requested -- the instance mask covers a few per
            // cent of the frame, so this removes almost half of the kernel's traffic.
            const bool skip_empty = m != nullptr && lo > 0;
            for (size_t i = threadIdx.x; i < nvec; i += kSumThreads) {
                const uint4 mv = m ? ldg_stream_u4(m + i * 16) : make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
                if (skip_empty && ((mv.x | mv.y) | (mv.z | mv.w)) == 0u) continue;
                const uint4 cv = ldg_stream_u4(c + i * 16);
                const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w}, mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t v;
                    if ((mw[q] & 0xfefefefeu) == 0u) {
                        v = cw[q] & (mw[q] * 0xffu);                     // mask bytes 0/1 -> 0x00/0xff, no carries
                    } else {                                              // arbitrary uint8 masks multiply and wrap
                        v = 0u;
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            v |= ((((cw[q] >> (8 * b)) & 0xffu) * ((mw[q] >> (8 * b)) & 0xffu)) & 0xffu) << (8 * b);
                    }
                    const uint32_t in = bytes_ge(v, t_lo) & ~bytes_ge(v, t_hi1);
                    cnt += __popc(in);
                    sum = __dp4a(v & ((in >> 7) * 0xffu), 0x01010101u, (unsigned)sum);
                }
            }
            done = nvec * 16;
        }
        for (size_t i = done + threadIdx.x; i < plane; i += kSumThreads) {
            const int v = ((int)c[i] * (m ? (int)m[i] : 1)) & 0xff;
            const bool in = (v >= lo) && (v <= hi);
            cnt += in;
            sum += in ? v : 0;
        }
        cnt = warp_sum(cnt);
        sum = warp_sum(sum);
        if ((threadIdx.x & 31) == 0) { s_cnt[threadIdx.x >> 5] = cnt; s_sum[threadIdx.x >> 5] = sum; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tc = 0, ts = 0;
            for (int k = 0; k < kSumThreads / 32; ++k) { tc += s_cnt[k]; ts += s_sum[k]; }
            out[f] = make_int2(tc, ts);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// one thread per frame: 17 scalars + 96 keypoint columns
// ---------------------------------------------------------------------------------------------
struct MmScale { double fw, fh, depth; };

__device__ __forceinline__ void px_to_mm(double x, double y, const MmScale &k, double &mx, double &my) {
    mx = k.depth * (x - 256) / k.fw;          // ref proc/util.py:47-59, always 512x424 / 70.6x60 deg
    my = k.depth * (y - 212) / k.fh;
}

__device__ __forceinline__ float mean_height(int2 cs) {
    return cs.x > 0 ? (float)((double)cs.y / (double)cs.x) : 0.0f;
}

__device__ __forceinline__ int clip_floor_index(double v, int dim) {
    // np.clip(np.floor(v).astype(int), 0, dim-1); NaN / out-of-int64 values become INT64_MIN on x86 -> 0
    if (!(fabs(v) < 9.2e18)) return 0;
    const long long i = (long long)floor(v);
    return (int)(i < 0 ? 0 : (i > dim - 1 ? dim - 1 : i));
}

enum ScalarRow {
    S_CX_PX = 0, S_CY_PX, S_V2D_PX, S_V3D_PX, S_WIDTH_PX, S_LENGTH_PX, S_AREA_PX, S_CX_MM, S_CY_MM, S_V2D_MM,
    S_V3D_MM, S_WIDTH_MM, S_LENGTH_MM, S_AREA_MM, S_HEIGHT, S_ANGLE, S_VTHETA
};

template <class KP>
__global__ void __launch_bounds__(128)
scalars_keypoints_kernel(const uint8_t *__restrict__ cleaned, const double *__restrict__ centroid,
                         const double *__restrict__ angle_deg, const double *__restrict__ axis,
                         const KP *__restrict__ kpts, const int2 *__restrict__ sums, int n, int h, int w,
                         int chunk, MmScale mm, double *__restrict__ scalars, double *__restrict__ kcols) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const int p = (f % chunk == 0) ? f : f - 1;              // first frame of a chunk differences with itself
    const size_t N = (size_t)n;

    const double cx = centroid[2 * f], cy = centroid[2 * f + 1];
    const double pcx = centroid[2 * p], pcy = centroid[2 * p + 1];
    double mx, my, sx, sy, pmx, pmy;
    px_to_mm(cx, cy, mm, mx, my);
    px_to_mm(cx + 1, cy + 1, mm, sx, sy);
    px_to_mm(pcx, pcy, mm, pmx, pmy);
    const double step_x = fabs(sx - mx), step_y = fabs(sy - my);
    const double width = np_min2(axis[2 * f], axis[2 * f + 1]);
    const double length = np_max2(axis[2 * f], axis[2 * f + 1]);
    const int2 cs = sums[f];
    const float height = mean_height(cs), pheight = mean_height(sums[p]);

    if (scalars) {
        scalars[S_CX_PX * N + f] = cx;
        scalars[S_CY_PX * N + f] = cy;
        scalars[S_CX_MM * N + f] = mx;
        scalars[S_CY_MM * N + f] = my;
        scalars[S_WIDTH_PX * N + f] = width;
        scalars[S_LENGTH_PX * N + f] = length;
        scalars[S_AREA_PX * N + f] = (double)cs.x;
        scalars[S_WIDTH_MM * N + f] = width * step_y;
        scalars[S_LENGTH_MM * N + f] = length * step_x;
        scalars[S_AREA_MM * N + f] = (double)cs.x * ((step_x + step_y) / 2.0);
        scalars[S_HEIGHT * N + f] = (double)height;
        scalars[S_ANGLE * N + f] = angle_deg[f] * kPiOver180;
        const double vx = cx - pcx, vy = cy - pcy;
        const float vzf = __fsub_rn(height, pheight);
        const double vz2 = (double)__fmul_rn(vzf, vzf);
        scalars[S_V2D_PX * N + f] = hypot(vx, vy);
        scalars[S_V3D_PX * N + f] = sqrt(vx * vx + vy * vy + vz2);
        const double wx = mx - pmx, wy = my - pmy;
        scalars[S_V2D_MM * N + f] = hypot(wx, wy);
        scalars[S_V3D_MM * N + f] = sqrt(wx * wx + wy * wy + vz2);
        scalars[S_VTHETA * N + f] = atan2(wy, wx);
    }

    if (kcols) {
        const double t = (-angle_deg[f]) * kPiOver180;
        const double c = cos(t), s = sin(t);
        const uint8_t *cl = cleaned + (size_t)f * h * w;
#pragma unroll 1
        for (int k = 0; k < MSQ_NUM_KEYPOINTS; ++k) {
            const double x = (double)kpts[(f * 8 + k) * 3], y = (double)kpts[(f * 8 + k) * 3 + 1];
            const double score = (double)kpts[(f * 8 + k) * 3 + 2];
            const double z = (double)cl[(size_t)clip_floor_index(y, h) * w + clip_floor_index(x, w)];
            double xm, ym, rx, ry, rmx, rmy;
            px_to_mm(x, y, mm, xm, ym);
            rotate_about(x, y, cx, cy, c, s, rx, ry);
            rotate_about(xm, ym, mx, my, c, s, rmx, rmy);
            double *o = kcols + (size_t)(k * 12) * N + f;
            o[0 * N] = x;  o[1 * N] = y;  o[2 * N] = score;  o[3 * N] = xm;  o[4 * N] = ym;  o[5 * N] = z;
            o[6 * N] = rx - cx;  o[7 * N] = ry - cy;  o[8 * N] = score;
            o[9 * N] = rmx - mx;  o[10 * N] = rmy - my;  o[11 * N] = z;
        }
    }
}

const char *const kScalarNames[MSQ_NUM_SCALARS] = {
    "centroid_x_px", "centroid_y_px", "velocity_2d_px", "velocity_3d_px", "width_px", "length_px", "area_px",
    "centroid_x_mm", "centroid_y_mm", "velocity_2d_mm", "velocity_3d_mm", "width_mm", "length_mm", "area_mm",
    "height_ave_mm", "angle", "velocity_theta"};
const char *const kKeypointNames[MSQ_NUM_KEYPOINTS] = {"Nose", "Left Ear", "Right Ear", "Neck", "Left Hip",
                                                       "Right Hip", "TailBase", "TailTip"};
const char *const kKeypointFields[12] = {"reference/%s_x_px", "reference/%s_y_px", "reference/%s_score",
                                         "reference/%s_x_mm", "reference/%s_y_mm", "reference/%s_z_mm",
                                         "rotated/%s_x_px", "rotated/%s_y_px", "rotated/%s_score",
                                         "rotated/%s_x_mm", "rotated/%s_y_mm", "rotated/%s_z_mm"};

}  // namespace

int launch_angles_and_flips(const double *orientation, const double *axis, const double *centroid, const float *kpts,
                            int n, int chunk, double *angle_out, uint8_t *flips, double *conf, int32_t *passes,
                            cudaStream_t st) {
    const int chunks = (n + chunk - 1) / chunk;
    const size_t smem = (size_t)2 * std::min(chunk, n) * sizeof(double);
    MSQ_REQUIRE(smem <= 200 * 1024, MSQ_EUNSUPPORTED, "angles_and_flips: chunk of %d frames is too large", chunk);
    if (smem > 48 * 1024)
        MSQ_CUDA_OK(cudaFuncSetAttribute(angles_flips_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TimedLaunch timed(K_ANGLES, st, 2);
    angle_votes_kernel<<<(n + 127) / 128, 128, 0, st>>>(orientation, axis, centroid, kpts, n, angle_out, flips, conf);
    angles_flips_kernel<<<chunks, kAngleThreads, smem, st>>>(n, chunk, angle_out, flips, passes);
    MSQ_LAUNCH_OK("angles_and_flips");
    return MSQ_OK;
}

int launch_masked_sums(const uint8_t *chunk_frames, const uint8_t *mask, int n, int h, int w, double min_h, double max_h,
                       int2 *sums_scratch, cudaStream_t st) {
    const size_t plane = (size_t)h * w;
    const int vec_ok = (plane % 16 == 0) && ((uintptr_t)chunk_frames % 16 == 0) && ((uintptr_t)mask % 16 == 0);   // NULL mask is "aligned"
    TimedLaunch timed(K_MASKED_SUMS, st);
    masked_sums_kernel<<<std::min(n, sm_count() * 8), kSumThreads, 0, st>>>(chunk_frames, mask, n, plane, min_h, max_h,
                                                                          vec_ok, sums_scratch);
    MSQ_LAUNCH_OK("masked_sums");
    return MSQ_OK;
}

int launch_scalars_and_keypoints(const uint8_t *chunk_frames, const uint8_t *mask, const uint8_t *cleaned,
                                 const double *centroid, const double *angle_deg, const double *axis,
                                 const void *kpts, bool kpts_f64, int n, int h, int w, int chunk, double min_h,
                                 double max_h, double true_depth, double *scalars, double *kcols, int2 *sums_scratch,
                                 cudaStream_t st, bool sums_done) {
    if (!sums_done) {
        const int rc = launch_masked_sums(chunk_frames, mask, n, h, w, min_h, max_h, sums_scratch, st);
        if (rc != MSQ_OK) return rc;
    }
    MmScale mm;
    // ref proc/util.py:53-54: f = resolution / (2 * deg2rad(fov / 2)), same float64 operation order
    mm.fw = 512 / (2 * ((70.6 / 2) * kPiOver180));
    mm.fh = 424 / (2 * ((60.0 / 2) * kPiOver180));
    mm.depth = true_depth;
    TimedLaunch timed(K_SCALARS_KPTS, st);
    if (kpts_f64)
        scalars_keypoints_kernel<double><<<(n + 127) / 128, 128, 0, st>>>(cleaned, centroid, angle_deg, axis,
                                                                         static_cast<const double *>(kpts), sums_scratch, n, h,
                                                                         w, chunk, mm, scalars, kcols);
    else
        scalars_keypoints_kernel<float><<<(n + 127) / 128, 128, 0, st>>>(cleaned, centroid, angle_deg, axis,
                                                                        static_cast<const float *>(kpts), sums_scratch, n, h,
                                                                        w, chunk, mm, scalars, kcols);
    MSQ_LAUNCH_OK("scalars_and_keypoints");
    return MSQ_OK;
}

}  // namespace msq

using namespace msq;

extern "C" int msq_angles_and_flips(const double *orientation, const double *axis, const double *centroid,
                                    const float *kpts, int n, int chunk, double *angle_out, uint8_t *flips,
                                    double *conf, int32_t *passes, void *stream) {
    MSQ_REQUIRE(orientation && axis && centroid && kpts && angle_out && flips, MSQ_EINVAL,
                "msq_angles_and_flips: null pointer");
    MSQ_REQUIRE(n >= 0 && chunk > 0, MSQ_EINVAL, "msq_angles_and_flips: bad sizes n=%d chunk=%d", n, chunk);
    if (n == 0) return MSQ_OK;
    return launch_angles_and_flips(orientation, axis, centroid, kpts, n, chunk, angle_out, flips, conf, passes,
                                   (cudaStream_t)stream);
}

static int flips_from_keypoints_any(const void *kpts, bool f64, const double *centroid, const double *angles,
                                   const double *lengths, int n, uint8_t *flips, double *conf, void *stream) {
    MSQ_REQUIRE(kpts && centroid && angles && lengths && flips, MSQ_EINVAL, "msq_flips_from_keypoints: null pointer");
    MSQ_REQUIRE(n >= 0, MSQ_EINVAL, "msq_flips_from_keypoints: bad n=%d", n);
    if (n == 0) return MSQ_OK;
    TimedLaunch timed(K_ANGLES, (cudaStream_t)stream);
    if (f64)
        flips_kernel<double><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(static_cast<const double *>(kpts), centroid, angles, lengths, n, flips, conf);
    else
        flips_kernel<float><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(static_cast<const float *>(kpts), centroid, angles, lengths, n, flips, conf);
    MSQ_LAUNCH_OK("flips_from_keypoints");
    return MSQ_OK;
}
extern "C" int msq_flips_from_keypoints(const float *kpts, const double *centroid, const double *angles,
                                        const double *lengths, int n, uint8_t *flips, double *conf, void *stream) {
    return flips_from_keypoints_any(kpts, false, centroid, angles, lengths, n, flips, conf, stream);
}
extern "C" int msq_flips_from_keypoints_f64(const double *kpts, const double *centroid, const double *angles,
                                            const double *lengths, int n, uint8_t *flips, double *conf, void *stream) {
    return flips_from_keypoints_any(kpts, true, centroid, angles, lengths, n, flips, conf, stream);
}

extern "C" int msq_iterative_filter_angles(const double *angles, int n, int chunk, int window, double tolerance,
                                           int max_iters, double *out, uint8_t *flips, int32_t *passes, void *stream) {
    MSQ_REQUIRE(angles && out, MSQ_EINVAL, "msq_iterative_filter_angles: null pointer");
    MSQ_REQUIRE(n >= 0 && chunk > 0 && window >= 1 && window <= 15 && max_iters >= 0, MSQ_EINVAL,
                "msq_iterative_filter_angles: bad arguments n=%d chunk=%d window=%d max_iters=%d", n, chunk, window, max_iters);
    if (n == 0) return MSQ_OK;
    const size_t smem = (size_t)2 * std::min(chunk, n) * sizeof(double);
    MSQ_REQUIRE(smem <= 200 * 1024, MSQ_EUNSUPPORTED, "iterative_filter_angles: chunk of %d frames is too large", chunk);
    if (smem > 48 * 1024)
        MSQ_CUDA_OK(cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TimedLaunch timed(K_ANGLES, (cudaStream_t)stream);
    filter_kernel<<<(n + chunk - 1) / chunk, kAngleThreads, smem, (cudaStream_t)stream>>>(angles, n, chunk, window,
                                                                                         tolerance, max_iters, out, flips, passes);
    MSQ_LAUNCH_OK("iterative_filter_angles");
    return MSQ_OK;
}

extern "C" size_t msq_scalars_scratch_bytes(int n) { return (size_t)(n > 0 ? n : 0) * sizeof(int2); }

static int scalars_and_keypoints_any(const uint8_t *chunk_frames, const uint8_t *mask, const uint8_t *cleaned,
                                    const double *centroid, const double *angle_deg, const double *axis, const void *kpts,
                                    bool kpts_f64, int n, int h, int w, int chunk, double min_h, double max_h,
                                    double true_depth, double *scalars, double *kcols, void *scratch, size_t scratch_bytes,
                                    void *stream) {
    MSQ_REQUIRE(chunk_frames && cleaned && centroid && angle_deg && axis && kpts, MSQ_EINVAL,
                "msq_scalars_and_keypoints: null input pointer");        // mask may be NULL (= all ones)
    MSQ_REQUIRE(scalars || kcols, MSQ_EINVAL, "msq_scalars_and_keypoints: both outputs are null");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0 && chunk > 0, MSQ_EINVAL, "msq_scalars_and_keypoints: bad sizes");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(scratch && scratch_bytes >= msq_scalars_scratch_bytes(n) && (uintptr_t)scratch % 8 == 0, MSQ_ENOMEM,
                "msq_scalars_and_keypoints: scratch must be 8-byte aligned and >= %zu bytes", msq_scalars_scratch_bytes(n));
    return launch_scalars_and_keypoints(chunk_frames, mask, cleaned, centroid, angle_deg, axis, kpts, kpts_f64, n, h, w, chunk,
                                        min_h, max_h, true_depth, scalars, kcols, reinterpret_cast<int2 *>(scratch),
                                        (cudaStream_t)stream);
}
extern "C" int msq_scalars_and_keypoints(const uint8_t *chunk_frames, const uint8_t *mask, const uint8_t *cleaned,
                                         const double *centroid, const double *angle_deg, const double *axis,
                                         const float *kpts, int n, int h, int w, int chunk, double min_h,
                                         double max_h, double true_depth, double *scalars, double *kcols,
                                         void *scratch, size_t scratch_bytes, void *stream) {
    return scalars_and_keypoints_any(chunk_frames, mask, cleaned, centroid, angle_deg, axis, kpts, false, n, h, w, chunk, min_h,
                                     max_h, true_depth, scalars, kcols, scratch, scratch_bytes, stream);
}
extern "C" int msq_scalars_and_keypoints_f64(const uint8_t *chunk_frames, const uint8_t *mask, const uint8_t *cleaned,
                                             const double *centroid, const double *angle_deg, const double *axis,
                                             const double *kpts, int n, int h, int w, int chunk, double min_h,
                                             double max_h, double true_depth, double *scalars, double *kcols,
                                             void *scratch, size_t scratch_bytes, void *stream) {
    return scalars_and_keypoints_any(chunk_frames, mask, cleaned, centroid, angle_deg, axis, kpts, true, n, h, w, chunk, min_h,
                                     max_h, true_depth, scalars, kcols, scratch, scratch_bytes, stream);
}

extern "C" const char *msq_scalar_name(int i) {
    return (i >= 0 && i < MSQ_NUM_SCALARS) ? kScalarNames[i] : nullptr;
}

extern "C" const char *msq_keypoint_col_name(int i) {
    static thread_local char buf[64];
    if (i < 0 || i >= MSQ_NUM_KPT_COLS) return nullptr;
    snprintf(buf, sizeof(buf), kKeypointFields[i % 12], kKeypointNames[i / 12]);
    return buf;
}
