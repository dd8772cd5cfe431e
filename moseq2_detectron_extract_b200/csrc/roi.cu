// f4 (session setup): get_roi -- RANSAC plane fit of the background image, the connected regions lying on the plane,
// their ranking features and the dilated / eroded / hole-filled mask of every region.
//   ref proc/roi.py:14-103 (get_roi), :106-130 (plane_fit3), :133-212 (plane_ransac)
// The reference runs 1000 NumPy passes over the image for the RANSAC scores, skimage.measure.label / regionprops for
// the regions, and cv2.dilate + scipy binary_fill_holes once per region.  Here:
//   * ransac_score_kernel      one CTA per candidate triple: plane from 3 points (float64, the reference's operation
//                              order), then inlier count and distance sum over all usable pixels;
//   * plane_distance_kernel    |ax + by + cz + d| of every pixel and the "on the plane" bitmap;
//   * label_*_kernel           8-connected labelling by union-find with atomicMin (the root of a region is its
//                              raster-first pixel, so numbering the roots in index order reproduces skimage's labels);
//   * region_props_kernel      area, bounding box and farthest pixel from the image centre (exact integers: 4 d^2);
//   * region_roi_kernel        one CTA per region: the region as bit rows in shared memory, dilation / erosion by an
//                              arbitrary structuring element as OR / AND of shifted rows, hole filling as a 4-connected
//                              flood of the background from the image border (bitrows.cuh), then the byte mask.
// The random triples themselves are drawn on the host from np.random in the reference's order (proc/roi.py of this
// package), so a seeded run picks the same plane as the reference.
#include "common.cuh"
#include "bitrows.cuh"
#include <algorithm>
#include <limits.h>
#include <math.h>

namespace msq {
namespace {

constexpr int kRansacThreads = 256;
constexpr int kScanBlock = 1024;
constexpr int kRoiThreads = 256;

__device__ __forceinline__ double plane_distance(double x, double y, double z, const double pl[4]) {
    return fabs(((x * pl[0] + y * pl[1]) + z * pl[2]) + pl[3]);
}

__global__ void __launch_bounds__(kRansacThreads)
ransac_score_kernel(const int *__restrict__ idx, int npts, const double *__restrict__ depth, int W, const long long *__restrict__ sel,
                    double tol, double *__restrict__ planes, int *__restrict__ ninliers, double *__restrict__ sumdist) {
    __shared__ double pl[4];
    __shared__ int s_cnt[kRansacThreads / 32];
    __shared__ double s_sum[kRansacThreads / 32];
    const int c = blockIdx.x, t = threadIdx.x;
    if (t == 0) {
        double p[3][3];
        for (int k = 0; k < 3; ++k) {
            const int pix = idx[sel[c * 3 + k]];
            const int y = pix / W;
            p[k][0] = (double)(pix - y * W); p[k][1] = (double)y; p[k][2] = depth[pix];
        }
        // plane_fit3: normal = (p1 - p0) x (p2 - p0), normalised; d = -p0 . normal; degenerate triples give NaN
        const double ax = p[1][0] - p[0][0], ay = p[1][1] - p[0][1], az = p[1][2] - p[0][2];
        const double bx = p[2][0] - p[0][0], by = p[2][1] - p[0][1], bz = p[2][2] - p[0][2];
        double n0 = ay * bz - az * by, n1 = az * bx - ax * bz, n2 = ax * by - ay * bx;
        const double denom = (n0 * n0 + n1 * n1) + n2 * n2;
        if (denom < 2.220446049250313e-16) {                     // np.spacing(1)
            pl[0] = pl[1] = pl[2] = pl[3] = nan("");
        } else {
            const double s = sqrt(denom);
            n0 /= s; n1 /= s; n2 /= s;
            pl[0] = n0; pl[1] = n1; pl[2] = n2;
            pl[3] = ((-p[0][0]) * n0 + (-p[0][1]) * n1) + (-p[0][2]) * n2;
        }
        for (int k = 0; k < 4; ++k) planes[c * 4 + k] = pl[k];
    }
    __syncthreads();
    if (isnan(pl[0])) {
        if (t == 0) { ninliers[c] = 0; sumdist[c] = nan(""); }
        return;
    }
    const double q[4] = {pl[0], pl[1], pl[2], pl[3]};
    int cnt = 0;
    double sum = 0.0;
    for (int i = t; i < npts; i += kRansacThreads) {
        const int pix = idx[i];
        const int y = pix / W;
        const double d = plane_distance((double)(pix - y * W), (double)y, depth[pix], q);
        cnt += d < tol;
        sum += d;
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        sum += __shfl_down_sync(0xffffffffu, sum, o);
    }
    if ((t & 31) == 0) { s_cnt[t >> 5] = cnt; s_sum[t >> 5] = sum; }
    __syncthreads();
    if (t == 0) {
        for (int k = 1; k < kRansacThreads / 32; ++k) { cnt += s_cnt[k]; sum += s_sum[k]; }
        ninliers[c] = cnt;
        sumdist[c] = sum;
    }
}

struct PlaneArg { double v[4]; };

__global__ void __launch_bounds__(256)
plane_distance_kernel(const double *__restrict__ depth, int H, int W, PlaneArg pl, double tol, const uint8_t *__restrict__ valid,
                      double *__restrict__ dist, uint8_t *__restrict__ on_plane) {
    const int total = H * W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const int y = p / W;
        const double d = plane_distance((double)(p - y * W), (double)y, depth[p], pl.v);
        if (dist) dist[p] = d;
        if (on_plane) on_plane[p] = (uint8_t)((!valid || valid[p]) && d < tol);
    }
}

// ---- gradient pre-mask: |Sobel_x| < thr and |Sobel_y| < thr (get_roi's gradient_filter, proc/roi.py:29-35) ----------------
struct SobelTaps { double deriv[31]; double smooth[31]; int n_deriv, n_smooth; };

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// cv2.Sobel(src, CV_64F, 1, 0, ksize) and (0, 1): separable correlation, rows first, BORDER_REFLECT_101
__device__ double separable_at(const double *__restrict__ src, int H, int W, int y, int x, const double *kx, int nx, const double *ky, int ny) {
    double acc = 0.0;
    for (int i = 0; i < ny; ++i) {
        const double *row = src + (size_t)reflect101(y + i - ny / 2, H) * W;
        double r = 0.0;
        for (int j = 0; j < nx; ++j) r += kx[j] * row[reflect101(x + j - nx / 2, W)];
        acc += ky[i] * r;
    }
    return acc;
}

__global__ void __launch_bounds__(256)
sobel_mask_kernel(const double *__restrict__ depth, int H, int W, SobelTaps taps, double threshold, uint8_t *__restrict__ mask) {
    const int total = H * W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const int y = p / W, x = p - y * W;
        const double gx = fabs(separable_at(depth, H, W, y, x, taps.deriv, taps.n_deriv, taps.smooth, taps.n_smooth));
        const double gy = fabs(separable_at(depth, H, W, y, x, taps.smooth, taps.n_smooth, taps.deriv, taps.n_deriv));
        mask[p] = (uint8_t)(gx < threshold && gy < threshold);
    }
}

// ---- 8-connected labelling -------------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int *L, int a) {
    const volatile int *V = L;
    int p;
    while ((p = V[a]) != a) a = p;
    return a;
}
__device__ __forceinline__ void uf_union(int *L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[b], a);                      // hang the larger root under the smaller one
        if (old == b) return;
        b = old;                                                  // somebody re-rooted b meanwhile: retry from there
    }
}

__global__ void __launch_bounds__(256) label_init_kernel(const uint8_t *__restrict__ bin, int total, int *__restrict__ L) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) L[p] = bin[p] ? p : -1;
}

__global__ void __launch_bounds__(256) label_merge_kernel(const uint8_t *__restrict__ bin, int H, int W, int *L) {
    const int total = H * W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        if (!bin[p]) continue;
        const int y = p / W, x = p - y * W;
        if (x > 0 && bin[p - 1]) uf_union(L, p, p - 1);
        if (y > 0) {
            if (bin[p - W]) uf_union(L, p, p - W);
            if (x > 0 && bin[p - W - 1]) uf_union(L, p, p - W - 1);
            if (x + 1 < W && bin[p - W + 1]) uf_union(L, p, p - W + 1);
        }
    }
}

// L[p] <- root(p); counts the roots of every kScanBlock-pixel slab
__global__ void __launch_bounds__(kScanBlock) label_flatten_kernel(int *L, int total, int *__restrict__ slab_roots) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int p = blockIdx.x * kScanBlock + threadIdx.x;
    bool root = false;
    if (p < total && L[p] >= 0) {
        const int r = uf_find(L, p);
        root = (r == p);
        if (!root) L[p] = r;
    }
    const uint32_t b = __ballot_sync(0xffffffffu, root);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&cnt, __popc(b));
    __syncthreads();
    if (threadIdx.x == 0) slab_roots[blockIdx.x] = cnt;
}

// ids[root] = 1 + number of roots before it in raster order
__global__ void __launch_bounds__(kScanBlock)
label_number_kernel(const int *__restrict__ L, int total, const int *__restrict__ slab_roots, int *__restrict__ ids, int *__restrict__ n_regions) {
    __shared__ int warp_sum[kScanBlock / 32];
    __shared__ int base;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int before = 0;
    for (int s = t; s < (int)blockIdx.x; s += kScanBlock) before += slab_roots[s];
    for (int o = 16; o > 0; o >>= 1) before += __shfl_down_sync(0xffffffffu, before, o);
    if (lane == 0) warp_sum[warp] = before;
    __syncthreads();
    if (t == 0) {
        int s = 0;
        for (int k = 0; k < kScanBlock / 32; ++k) s += warp_sum[k];
        base = s;
    }
    __syncthreads();
    const int p = blockIdx.x * kScanBlock + t;
    const bool root = p < total && L[p] == p;
    const uint32_t b = __ballot_sync(0xffffffffu, root);
    __syncthreads();
    if (lane == 0) warp_sum[warp] = __popc(b);
    __syncthreads();
    int prior = base;
    for (int k = 0; k < warp; ++k) prior += warp_sum[k];
    if (root) ids[p] = prior + __popc(b & ((1u << lane) - 1u)) + 1;
    if (blockIdx.x == gridDim.x - 1 && t == kScanBlock - 1) n_regions[0] = prior + __popc(b);
}

__global__ void __launch_bounds__(256)
label_assign_kernel(const int *__restrict__ L, const int *__restrict__ ids, int total, int *__restrict__ labels) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const int r = L[p];
        labels[p] = r >= 0 ? ids[r] : 0;
    }
}

// ---- regionprops: area, bbox, farthest pixel from the image centre ------------------------------------------------
__global__ void __launch_bounds__(256)
region_props_init_kernel(int n, int *__restrict__ area, int *__restrict__ bbox, unsigned *__restrict__ maxd4) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        area[r] = 0; maxd4[r] = 0u;
        bbox[4 * r] = INT_MAX; bbox[4 * r + 1] = INT_MAX; bbox[4 * r + 2] = -1; bbox[4 * r + 3] = -1;
    }
}

__global__ void __launch_bounds__(256)
region_props_kernel(const int *__restrict__ labels, int H, int W, int n, int *area, int *bbox, unsigned *maxd4) {
    const int total = H * W;
    const int rounds = (total + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (int it = 0; it < rounds; ++it) {
        const int p = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        const int l = p < total ? labels[p] : 0;
        const bool valid = l > 0 && l <= n;
        const uint32_t active = __ballot_sync(0xffffffffu, valid);
        if (!valid) continue;
        // neighbouring pixels mostly share a label: one atomic per (warp, label) instead of one per pixel
        const uint32_t peers = __match_any_sync(active, l);
        const int y = p / W, x = p - y * W;
        const int dy = 2 * y - H, dx = 2 * x - W;                  // 2 * (coords - shape / 2), exact for odd sizes too
        const unsigned d4 = (unsigned)(dy * dy + dx * dx);
        const int cnt = __popc(peers);
        const int ymin = __reduce_min_sync(peers, y), ymax = __reduce_max_sync(peers, y);
        const int xmin = __reduce_min_sync(peers, x), xmax = __reduce_max_sync(peers, x);
        const unsigned dmax = __reduce_max_sync(peers, d4);
        if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
            const int r = l - 1;
            atomicAdd(&area[r], cnt);
            atomicMin(&bbox[4 * r], ymin); atomicMin(&bbox[4 * r + 1], xmin);
            atomicMax(&bbox[4 * r + 2], ymax); atomicMax(&bbox[4 * r + 3], xmax);
            atomicMax(&maxd4[r], dmax);
        }
    }
}

// ---- per-region mask: dilate / erode / fill holes --------------------------------------------------------------------
__device__ __forceinline__ uint32_t word_mask(int k, int W) {
    const int left = W - 32 * k;
    return left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << left) - 1u));
}

// word k of the row shifted so that bit x of the result is bit x + dx of `row` (0 outside the image);
// `complement` reads the row inverted (inside the image only)
__device__ __forceinline__ uint32_t shifted_word(const uint32_t *row, int k, int lpr, int W, int dx, bool complement) {
    const int t = 32 * k + dx;
    const int q = t >> 5, o = t & 31;
    uint32_t w0 = 0u, w1 = 0u;
    if (q >= 0 && q < lpr) w0 = complement ? (~row[q] & word_mask(q, W)) : row[q];
    if (q + 1 >= 0 && q + 1 < lpr) w1 = complement ? (~row[q + 1] & word_mask(q + 1, W)) : row[q + 1];
    return __funnelshift_r(w0, w1, o);
}

// cv2.dilate / cv2.erode of a binary image with an arbitrary structuring element anchored at its centre and OpenCV's
// default border (pixels outside the image never win): dst(x,y) = max|min over se(i,j) != 0 of src(x + j - ax, y + i - ay)
__device__ void morph_bits(const uint32_t *src, uint32_t *dst, int H, int W, int lpr, const uint8_t *__restrict__ se, int kh, int kw,
                           bool erode) {
    const int ay = kh / 2, ax = kw / 2;
    for (int i = threadIdx.x; i < H * lpr; i += kRoiThreads) {
        const int r = i / lpr, k = i - r * lpr;
        uint32_t acc = 0u;
        for (int si = 0; si < kh; ++si) {
            const int rr = r + si - ay;
            if (rr < 0 || rr >= H) continue;
            const uint32_t *row = src + rr * lpr;
            for (int sj = 0; sj < kw; ++sj)
                if (__ldg(se + si * kw + sj)) acc |= shifted_word(row, k, lpr, W, sj - ax, erode);
        }
        dst[i] = (erode ? ~acc : acc) & word_mask(k, W);
    }
}

__global__ void __launch_bounds__(kRoiThreads)
region_roi_kernel(const int *__restrict__ labels, int H, int W, const int *__restrict__ order, const uint8_t *__restrict__ se_d, int dh, int dw,
                  const uint8_t *__restrict__ se_e, int eh, int ew, int fill_holes, uint8_t *__restrict__ rois, int *__restrict__ bboxes) {
    extern __shared__ uint32_t planes[];
    __shared__ int box[4];
    const int lpr = (W + 31) >> 5;
    uint32_t *cur = planes, *other = planes + H * lpr;
    const int id = order[blockIdx.x] + 1;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) { box[0] = INT_MAX; box[1] = INT_MAX; box[2] = -1; box[3] = -1; }

    for (int r = warp; r < H; r += kRoiThreads / 32)
        for (int k = 0; k < lpr; ++k) {
            const int x = 32 * k + lane;
            const uint32_t word = __ballot_sync(0xffffffffu, x < W && labels[(size_t)r * W + x] == id);
            if (lane == 0) cur[r * lpr + k] = word;
        }
    __syncthreads();
    if (se_d) {
        morph_bits(cur, other, H, W, lpr, se_d, dh, dw, false);
        __syncthreads();
        uint32_t *s = cur; cur = other; other = s;
    }
    if (se_e) {
        morph_bits(cur, other, H, W, lpr, se_e, eh, ew, true);
        __syncthreads();
        uint32_t *s = cur; cur = other; other = s;
    }
    if (fill_holes) {
        // scipy.ndimage.binary_fill_holes: background pixels that are not 4-connected to the outside become foreground.
        // `other` collects the background reached from the border; alternating down / up sweeps until nothing changes.
        if (warp == 0) {
            const bool act = lane < lpr;
            const uint32_t wm = word_mask(lane, W);
            const uint32_t edge = (lane == 0 ? 1u : 0u) | (lane == ((W - 1) >> 5) ? (1u << ((W - 1) & 31)) : 0u);
            for (int r = 0; r < H; ++r)
                if (act) other[r * lpr + lane] = 0u;
            __syncwarp();
            bool down = true;
            for (int sweep = 0;; ++sweep) {
                bool changed = false;
                uint32_t prev = wm;                                 // the row outside the image counts as reached
                const int r_begin = down ? 0 : H - 1, r_end = down ? H : -1, dr = down ? 1 : -1;
                for (int r = r_begin; r != r_end; r += dr) {
                    const uint32_t open = act ? (~cur[r * lpr + lane] & wm) : 0u;
                    const uint32_t old = act ? other[r * lpr + lane] : 0u;
                    const uint32_t now = fill_row((prev | old | edge) & open, open, lane);
                    if (act) other[r * lpr + lane] = now;
                    changed |= (now != old);
                    prev = now;
                }
                if (!__any_sync(0xffffffffu, changed) && sweep > 0) break;
                down = !down;
            }
        }
        __syncthreads();
        for (int i = t; i < H * lpr; i += kRoiThreads) cur[i] = ~other[i] & word_mask(i % lpr, W);
        __syncthreads();
    }
    // byte mask + bounding box of the result
    int ymin = INT_MAX, ymax = -1, xmin = INT_MAX, xmax = -1;
    for (int i = t; i < H * lpr; i += kRoiThreads) {
        const uint32_t w = cur[i];
        if (!w) continue;
        const int r = i / lpr, k = i - r * lpr;
        ymin = min(ymin, r); ymax = max(ymax, r);
        xmin = min(xmin, 32 * k + __ffs(w) - 1); xmax = max(xmax, 32 * k + 31 - __clz(w));
    }
    if (ymax >= 0) { atomicMin(&box[0], ymin); atomicMin(&box[1], xmin); atomicMax(&box[2], ymax); atomicMax(&box[3], xmax); }
    uint8_t *out = rois + (size_t)blockIdx.x * H * W;
    for (int p = t; p < H * W; p += kRoiThreads) {
        const int y = p / W, x = p - y * W;
        out[p] = (uint8_t)((cur[y * lpr + (x >> 5)] >> (x & 31)) & 1u);
    }
    __syncthreads();
    if (t < 4) bboxes[blockIdx.x * 4 + t] = box[2] >= 0 ? box[t] : -1;
}

int grid_for(int total) { return std::max(1, std::min((total + 255) / 256, sm_count() * 8)); }

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" int msq_plane_ransac_score(const int *idx_dev, int npts, const double *depth_dev, int H, int W, const long long *sel_dev,
                                      int iters, double tol, double *planes_dev, int *ninliers_dev, double *sumdist_dev, void *stream) {
    MSQ_REQUIRE(H > 0 && W > 0 && (long long)H * W < INT_MAX, MSQ_EINVAL, "msq_plane_ransac_score: bad image size %dx%d", H, W);
    MSQ_REQUIRE(npts >= 1 && iters >= 0, MSQ_EINVAL, "msq_plane_ransac_score: need at least one usable pixel (npts=%d, iters=%d)", npts, iters);
    MSQ_REQUIRE(idx_dev && depth_dev && sel_dev && planes_dev && ninliers_dev && sumdist_dev, MSQ_EINVAL, "msq_plane_ransac_score: null pointer");
    if (iters == 0) return MSQ_OK;
    TimedLaunch timed(K_ROI, (cudaStream_t)stream);
    ransac_score_kernel<<<iters, kRansacThreads, 0, (cudaStream_t)stream>>>(idx_dev, npts, depth_dev, W, sel_dev, tol, planes_dev,
                                                                            ninliers_dev, sumdist_dev);
    MSQ_LAUNCH_OK("ransac_score");
    return MSQ_OK;
}

extern "C" int msq_plane_distance(const double *depth_dev, int H, int W, const double *plane_host, double tol, const uint8_t *valid_dev,
                                  double *dist_dev, uint8_t *on_plane_dev, void *stream) {
    MSQ_REQUIRE(H > 0 && W > 0 && (long long)H * W < INT_MAX, MSQ_EINVAL, "msq_plane_distance: bad image size %dx%d", H, W);
    MSQ_REQUIRE(depth_dev && plane_host && (dist_dev || on_plane_dev), MSQ_EINVAL, "msq_plane_distance: null pointer");
    PlaneArg pl;
    for (int k = 0; k < 4; ++k) pl.v[k] = plane_host[k];
    TimedLaunch timed(K_ROI, (cudaStream_t)stream);
    plane_distance_kernel<<<grid_for(H * W), 256, 0, (cudaStream_t)stream>>>(depth_dev, H, W, pl, tol, valid_dev, dist_dev, on_plane_dev);
    MSQ_LAUNCH_OK("plane_distance");
    return MSQ_OK;
}

extern "C" int msq_sobel_gradient_mask(const double *depth_dev, int H, int W, const double *deriv_host, int n_deriv,
                                       const double *smooth_host, int n_smooth, double threshold, uint8_t *mask_dev, void *stream) {
    MSQ_REQUIRE(H > 0 && W > 0 && (long long)H * W < INT_MAX, MSQ_EINVAL, "msq_sobel_gradient_mask: bad image size %dx%d", H, W);
    MSQ_REQUIRE(depth_dev && deriv_host && smooth_host && mask_dev, MSQ_EINVAL, "msq_sobel_gradient_mask: null pointer");
    MSQ_REQUIRE(n_deriv >= 1 && n_deriv <= 31 && n_smooth >= 1 && n_smooth <= 31 && (n_deriv & 1) && (n_smooth & 1), MSQ_EINVAL,
                "msq_sobel_gradient_mask: kernels must have an odd number of taps <= 31 (got %d, %d)", n_deriv, n_smooth);
    SobelTaps taps;
    taps.n_deriv = n_deriv; taps.n_smooth = n_smooth;
    for (int k = 0; k < 31; ++k) { taps.deriv[k] = k < n_deriv ? deriv_host[k] : 0.0; taps.smooth[k] = k < n_smooth ? smooth_host[k] : 0.0; }
    TimedLaunch timed(K_ROI, (cudaStream_t)stream);
    sobel_mask_kernel<<<grid_for(H * W), 256, 0, (cudaStream_t)stream>>>(depth_dev, H, W, taps, threshold, mask_dev);
    MSQ_LAUNCH_OK("sobel_mask");
    return MSQ_OK;
}

extern "C" size_t msq_label_scratch_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    const size_t total = (size_t)H * W;
    return align_up(total * 4, 16) * 2 + align_up(((total + kScanBlock - 1) / kScanBlock) * 4, 16);
}

extern "C" int msq_label_regions(const uint8_t *bin_dev, int H, int W, int *labels_dev, int *n_regions_dev, void *scratch_dev,
                                 size_t scratch_bytes, void *stream) {
    MSQ_REQUIRE(H > 0 && W > 0 && (long long)H * W < INT_MAX, MSQ_EINVAL, "msq_label_regions: bad image size %dx%d", H, W);
    MSQ_REQUIRE(bin_dev && labels_dev && n_regions_dev, MSQ_EINVAL, "msq_label_regions: null pointer");
    MSQ_REQUIRE(scratch_dev && scratch_bytes >= msq_label_scratch_bytes(H, W) && (uintptr_t)scratch_dev % 16 == 0, MSQ_ENOMEM,
                "msq_label_regions: scratch must hold %zu bytes, 16-byte aligned", msq_label_scratch_bytes(H, W));
    cudaStream_t st = (cudaStream_t)stream;
    const int total = H * W, slabs = (total + kScanBlock - 1) / kScanBlock;
    int *L = static_cast<int *>(scratch_dev);
    int *ids = reinterpret_cast<int *>(static_cast<char *>(scratch_dev) + align_up((size_t)total * 4, 16));
    int *slab_roots = reinterpret_cast<int *>(static_cast<char *>(scratch_dev) + 2 * align_up((size_t)total * 4, 16));
    TimedLaunch timed(K_ROI, st);
    label_init_kernel<<<grid_for(total), 256, 0, st>>>(bin_dev, total, L);
    label_merge_kernel<<<grid_for(total), 256, 0, st>>>(bin_dev, H, W, L);
    label_flatten_kernel<<<slabs, kScanBlock, 0, st>>>(L, total, slab_roots);
    label_number_kernel<<<slabs, kScanBlock, 0, st>>>(L, total, slab_roots, ids, n_regions_dev);
    label_assign_kernel<<<grid_for(total), 256, 0, st>>>(L, ids, total, labels_dev);
    MSQ_LAUNCH_OK("label_regions");
    return MSQ_OK;
}

extern "C" int msq_region_props(const int *labels_dev, int H, int W, int n_regions, int *area_dev, int *bbox_dev, unsigned *maxd4_dev,
                                void *stream) {
    MSQ_REQUIRE(H > 0 && W > 0 && H <= 16384 && W <= 16384, MSQ_EINVAL, "msq_region_props: bad image size %dx%d", H, W);
    MSQ_REQUIRE(n_regions >= 0, MSQ_EINVAL, "msq_region_props: n_regions=%d", n_regions);
    if (n_regions == 0) return MSQ_OK;
    MSQ_REQUIRE(labels_dev && area_dev && bbox_dev && maxd4_dev, MSQ_EINVAL, "msq_region_props: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    TimedLaunch timed(K_ROI, st);
    region_props_init_kernel<<<grid_for(n_regions), 256, 0, st>>>(n_regions, area_dev, bbox_dev, maxd4_dev);
    region_props_kernel<<<grid_for(H * W), 256, 0, st>>>(labels_dev, H, W, n_regions, area_dev, bbox_dev, maxd4_dev);
    MSQ_LAUNCH_OK("region_props");
    return MSQ_OK;
}

extern "C" int msq_region_rois(const int *labels_dev, int H, int W, const int *order_dev, int n_out, const uint8_t *se_dilate_dev, int dh,
                               int dw, const uint8_t *se_erode_dev, int eh, int ew, int fill_holes, uint8_t *rois_dev, int *bboxes_dev,
                               void *stream) {
    MSQ_REQUIRE(H > 0 && W > 0 && W <= 1024, MSQ_EUNSUPPORTED, "msq_region_rois: images up to 1024 pixels wide (got %dx%d)", H, W);
    MSQ_REQUIRE(n_out >= 0, MSQ_EINVAL, "msq_region_rois: n_out=%d", n_out);
    if (n_out == 0) return MSQ_OK;
    MSQ_REQUIRE(labels_dev && order_dev && rois_dev && bboxes_dev, MSQ_EINVAL, "msq_region_rois: null pointer");
    MSQ_REQUIRE(!se_dilate_dev || (dh > 0 && dw > 0), MSQ_EINVAL, "msq_region_rois: bad dilation element %dx%d", dh, dw);
    MSQ_REQUIRE(!se_erode_dev || (eh > 0 && ew > 0), MSQ_EINVAL, "msq_region_rois: bad erosion element %dx%d", eh, ew);
    const size_t smem = (size_t)2 * H * ((W + 31) / 32) * 4;
    MSQ_REQUIRE(smem <= 220 * 1024, MSQ_EUNSUPPORTED, "msq_region_rois: %dx%d needs %zu bytes of shared memory (limit 220 KB)", H, W, smem);
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        MSQ_CUDA_OK(cudaFuncSetAttribute(region_roi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    TimedLaunch timed(K_ROI, (cudaStream_t)stream);
    region_roi_kernel<<<n_out, kRoiThreads, smem, (cudaStream_t)stream>>>(labels_dev, H, W, order_dev, se_dilate_dev, dh, dw, se_erode_dev,
                                                                          eh, ew, fill_holes, rois_dev, bboxes_dev);
    MSQ_LAUNCH_OK("region_roi");
    return MSQ_OK;
}
