// a4 (R-CNN glue): segmented greedy NMS for the proposal filtering of a batch of images.
// torchvision's RegionProposalNetwork.filter_proposals runs one NMS per image (a kernel that builds an N x N suppression
// mask, a copy of it to the host and a sequential scan there): ~0.6 ms of launch / synchronisation latency per frame.
// Here ONE launch serves the whole batch: a warp per image walks that image's candidates in score order and keeps a box
// unless a box it kept earlier overlaps it by more than the threshold -- the same greedy rule -- and stops after
// `max_keep` boxes (the RPN only wants the first 100 survivors, which are found within the first few hundred candidates).
// Boxes arrive sorted by score and already shifted per pyramid level (torchvision's "coordinate trick", so that levels
// never suppress each other and the float32 arithmetic is the one torchvision does); IoU as in torchvision's devIoU.
#include "common.cuh"
#include <cuda_bf16.h>
#include <algorithm>
#include <limits.h>
#include <math.h>

namespace msq {
namespace {

constexpr int kNmsWarps = 4;

__device__ __forceinline__ bool iou_above(float4 a, float4 b, float thr) {
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z), top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    const float inter = width * height;
    const float sa = (a.z - a.x) * (a.w - a.y), sb = (b.z - b.x) * (b.w - b.y);
    return inter / (sa + sb - inter) > thr;
}

__global__ void __launch_bounds__(kNmsWarps * 32)
nms_sorted_kernel(const float4 *__restrict__ boxes, const uint8_t *__restrict__ valid, int n, int K, float thr, int max_keep,
                  int *__restrict__ keep, int *__restrict__ count) {
    extern __shared__ __align__(16) float4 kept_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x * kNmsWarps + warp;
    if (img >= n) return;
    float4 *kept = kept_all + (size_t)warp * max_keep;
    const float4 *b = boxes + (size_t)img * K;
    const uint8_t *v = valid + (size_t)img * K;
    int *out = keep + (size_t)img * max_keep;
    int kc = 0;
    for (int i = 0; i < K && kc < max_keep; ++i) {
        if (!v[i]) continue;                                   // (warp-uniform: every lane reads the same flag)
        const float4 box = __ldg(b + i);
        bool hit = false;
        for (int j = lane; j < kc; j += 32) hit |= iou_above(kept[j], box, thr);
        if (!__any_sync(0xffffffffu, hit)) {
            if (lane == 0) { kept[kc] = box; out[kc] = i; }
            ++kc;
            __syncwarp();
        }
    }
    for (int j = kc + lane; j < max_keep; j += 32) out[j] = -1;
    if (lane == 0) count[img] = kc;
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" int msq_nms_sorted(const float *boxes, const uint8_t *valid, int n, int K, float iou_threshold, int max_keep,
                              int32_t *keep, int32_t *count, void *stream) {
    MSQ_REQUIRE(n >= 0 && K >= 0 && max_keep > 0, MSQ_EINVAL, "msq_nms_sorted: bad sizes n=%d K=%d max_keep=%d", n, K, max_keep);
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(boxes && valid && keep && count, MSQ_EINVAL, "msq_nms_sorted: null pointer");
    MSQ_REQUIRE((uintptr_t)boxes % 16 == 0, MSQ_EINVAL, "msq_nms_sorted: boxes must be 16-byte aligned");
    const size_t smem = (size_t)kNmsWarps * max_keep * sizeof(float4);
    MSQ_REQUIRE(smem <= 200 * 1024, MSQ_EUNSUPPORTED, "msq_nms_sorted: max_keep %d is too large", max_keep);
    if (smem > 48 * 1024)
        MSQ_CUDA_OK(cudaFuncSetAttribute(nms_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TimedLaunch timed(K_DETECTOR_GLUE, (cudaStream_t)stream);
    nms_sorted_kernel<<<(n + kNmsWarps - 1) / kNmsWarps, kNmsWarps * 32, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(boxes), valid, n, K, iou_threshold, max_keep, keep, count);
    MSQ_LAUNCH_OK("nms_sorted");
    return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// The same greedy rule for LONG keep lists (detectron2's RPN keeps the first 1000 survivors of ~3000 candidates per image:
// the walk above would compare every candidate with up to 1000 kept boxes, 41 ms per 250 images).  Two kernels:
//   nms_mask_kernel   every (64 candidates) x (64 later candidates) block of the K x K overlap matrix in parallel, one bit
//                     per pair (upper triangle only);
//   nms_scan_kernel   a warp per image walks the candidates in score order: candidate i survives when its bit is clear in
//                     the running `removed` set (K / 64 words spread over the lanes), then ORs row i into the set.  Rows
//                     do not depend on the walk, so the rows of the next candidates are prefetched into L2.
// Results are identical to nms_sorted_kernel (greedy NMS either way).
// ---------------------------------------------------------------------------------------------------------------
namespace msq {
namespace {

__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4 *__restrict__ boxes, const uint8_t *__restrict__ valid, int K, int words, float thr,
                unsigned long long *__restrict__ mask /* (n, K, words) */) {
    const int img = blockIdx.z, rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;                                           // only later candidates can be suppressed
    __shared__ float4 col[64];
    __shared__ uint8_t col_ok[64];
    const float4 *b = boxes + (size_t)img * K;
    const uint8_t *v = valid + (size_t)img * K;
    const int c = cb * 64 + threadIdx.x;
    col[threadIdx.x] = c < K ? b[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    col_ok[threadIdx.x] = c < K ? v[c] : 0;
    __syncthreads();
    const int r = rb * 64 + threadIdx.x;
    if (r >= K) return;
    unsigned long long bits = 0ull;
    if (v[r]) {
        const float4 box = b[r];
        const int start = cb == rb ? threadIdx.x + 1 : 0;
        for (int j = start; j < 64; ++j)
            if (col_ok[j] && iou_above(box, col[j], thr)) bits |= 1ull << j;
    }
    mask[((size_t)img * K + r) * words + cb] = bits;
}

__global__ void __launch_bounds__(kNmsWarps * 32)
nms_scan_kernel(const unsigned long long *__restrict__ mask, const uint8_t *__restrict__ valid, int n, int K, int words, int max_keep,
                int *__restrict__ keep, int *__restrict__ count) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x * kNmsWarps + warp;
    if (img >= n) return;
    const unsigned long long *m = mask + (size_t)img * K * words;
    const uint8_t *v = valid + (size_t)img * K;
    int *out = keep + (size_t)img * max_keep;
    // lane l owns words l, l + 32, l + 64 (K <= 6144); row words left of the diagonal were never written: masked below
    unsigned long long rem0 = 0ull, rem1 = 0ull, rem2 = 0ull;
    int kc = 0;
    for (int i = 0; i < K && kc < max_keep; ++i) {
        const int w = i >> 6;
        const unsigned long long word = w < 32 ? rem0 : (w < 64 ? rem1 : rem2);
        const unsigned long long mine = __shfl_sync(0xffffffffu, word, w & 31);
        if (i + 6 < K && lane < words) asm volatile("prefetch.global.L2 [%0];" :: "l"(m + (size_t)(i + 6) * words + lane));
        if (!v[i] || ((mine >> (i & 63)) & 1ull)) continue;
        if (lane == 0) out[kc] = i;
        ++kc;
        const unsigned long long *row = m + (size_t)i * words;
        if (lane >= w && lane < words) rem0 |= row[lane];
        if (lane + 32 >= w && lane + 32 < words) rem1 |= row[lane + 32];
        if (lane + 64 >= w && lane + 64 < words) rem2 |= row[lane + 64];
    }
    for (int j = kc + lane; j < max_keep; j += 32) out[j] = -1;
    if (lane == 0) count[img] = kc;
}

}  // namespace
}  // namespace msq

extern "C" size_t msq_nms_scratch_bytes(int n, int K) {
    if (n <= 0 || K <= 0) return 0;
    return (size_t)n * K * ((K + 63) / 64) * sizeof(unsigned long long);
}

extern "C" int msq_nms_sorted_long(const float *boxes, const uint8_t *valid, int n, int K, float iou_threshold, int max_keep,
                                   int32_t *keep, int32_t *count, void *scratch, size_t scratch_bytes, void *stream) {
    MSQ_REQUIRE(n >= 0 && K >= 0 && max_keep > 0, MSQ_EINVAL, "msq_nms_sorted_long: bad sizes n=%d K=%d max_keep=%d", n, K, max_keep);
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(boxes && valid && keep && count && scratch, MSQ_EINVAL, "msq_nms_sorted_long: null pointer");
    MSQ_REQUIRE((uintptr_t)boxes % 16 == 0 && (uintptr_t)scratch % 8 == 0, MSQ_EINVAL, "msq_nms_sorted_long: boxes must be 16-byte, scratch 8-byte aligned");
    MSQ_REQUIRE(K <= 6144, MSQ_EUNSUPPORTED, "msq_nms_sorted_long: at most 6144 candidates per image (got %d)", K);
    MSQ_REQUIRE(scratch_bytes >= msq_nms_scratch_bytes(n, K), MSQ_ENOMEM, "msq_nms_sorted_long: scratch too small");
    const int words = (K + 63) / 64;
    cudaStream_t st = (cudaStream_t)stream;
    TimedLaunch timed(K_DETECTOR_GLUE, st, 2);
    nms_mask_kernel<<<dim3(words, words, n), 64, 0, st>>>(reinterpret_cast<const float4 *>(boxes), valid, K, words, iou_threshold,
                                                          static_cast<unsigned long long *>(scratch));
    nms_scan_kernel<<<(n + kNmsWarps - 1) / kNmsWarps, kNmsWarps * 32, 0, st>>>(static_cast<const unsigned long long *>(scratch), valid, n, K,
                                                                                 words, max_keep, keep, count);
    MSQ_LAUNCH_OK("nms_sorted_long");
    return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Long keep lists when the candidates carry their pyramid level (valid[i] = 1 + level, as msq_rpn_select writes it): levels
// never suppress each other, so the greedy walk decomposes into one walk per (image, level) and the overlap bits are only
// needed for pairs of the SAME level -- with 3 x 1000 + 192 + 48 candidates that is 1.5 M pairs per image instead of 5.2 M,
// and five short walks in parallel instead of one of 3240 steps.  Four launches:
//   nms_level_lists_kernel   per (image, level) the positions of its candidates in the score-sorted list, in order;
//   nms_level_mask_kernel    64 x 64 overlap bits per level (upper triangle), boxes looked up through the lists;
//   nms_level_scan_kernel    a warp per (image, level) walks its list and flags the suppressed candidates;
//   nms_collect_kernel       a warp per image keeps the first max_keep unflagged candidates of the score-sorted list.
// Same result as nms_sorted_kernel / msq_nms_sorted_long on the shifted boxes (cross-level overlaps are zero there).
// ---------------------------------------------------------------------------------------------------------------
namespace msq {
namespace {

constexpr int kNmsMaxLevels = 8;

__global__ void __launch_bounds__(32 * kNmsMaxLevels)
nms_level_lists_kernel(const uint8_t *__restrict__ valid, int K, int L, int Kl, int *__restrict__ lists /* (n, L, Kl) */,
                       int *__restrict__ counts /* (n, L) */, uint8_t *__restrict__ removed /* (n, K), zeroed */) {
    const int img = blockIdx.x, l = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (l >= L) return;
    const uint8_t *v = valid + (size_t)img * K;
    int *list = lists + ((size_t)img * L + l) * Kl;
    int cnt = 0;
    for (int base = 0; base < K; base += 32) {
        const int i = base + lane;
        const bool mine = i < K && v[i] == (uint8_t)(1 + l);
        const unsigned b = __ballot_sync(0xffffffffu, mine);
        if (mine) {
            const int pos = cnt + __popc(b & ((1u << lane) - 1u));
            if (pos < Kl) list[pos] = i;
            else removed[(size_t)img * K + i] = 1;               // more candidates than the caller announced: never kept
        }
        cnt += __popc(b);
    }
    if (lane == 0) counts[img * L + l] = min(cnt, Kl);
}

__global__ void __launch_bounds__(64)
nms_level_mask_kernel(const float4 *__restrict__ boxes, int K, int L, int Kl, int words, float thr, const int *__restrict__ lists,
                      const int *__restrict__ counts, unsigned long long *__restrict__ mask /* (n, L, Kl, words) */) {
    const int il = blockIdx.z, rb = blockIdx.y, cb = blockIdx.x;
    const int cnt = counts[il];
    if (cb < rb || cb * 64 >= cnt) return;                         // (rb <= cb, so rb * 64 < cnt as well)
    const int img = il / L;
    const float4 *b = boxes + (size_t)img * K;
    const int *list = lists + (size_t)il * Kl;
    __shared__ float4 col[64];
    const int c = cb * 64 + threadIdx.x;
    col[threadIdx.x] = c < cnt ? b[list[c]] : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int r = rb * 64 + threadIdx.x;
    if (r >= cnt) return;
    const float4 box = b[list[r]];
    const int last = min(64, cnt - cb * 64);
    unsigned long long bits = 0ull;
    for (int j = cb == rb ? threadIdx.x + 1 : 0; j < last; ++j)
        if (iou_above(box, col[j], thr)) bits |= 1ull << j;
    mask[((size_t)il * Kl + r) * words + cb] = bits;
}

__global__ void __launch_bounds__(kNmsWarps * 32)
nms_level_scan_kernel(const unsigned long long *__restrict__ mask, int n_il, int K, int L, int Kl, int words, int max_keep,
                      const int *__restrict__ lists, const int *__restrict__ counts, uint8_t *__restrict__ removed) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int il = blockIdx.x * kNmsWarps + warp;
    if (il >= n_il) return;
    const int img = il / L, cnt = counts[il];
    const unsigned long long *m = mask + (size_t)il * Kl * words;
    const int *list = lists + (size_t)il * Kl;
    uint8_t *rem_out = removed + (size_t)img * K;
    // lane w owns word w of the running "suppressed" set (words <= 32); row words left of the diagonal were never written
    unsigned long long rem = 0ull;
    int kc = 0;
    for (int i = 0; i < cnt; ++i) {
        const int w = i >> 6;
        const unsigned long long mine = __shfl_sync(0xffffffffu, rem, w);
        if (i + 6 < cnt && lane < words) asm volatile("prefetch.global.L2 [%0];" :: "l"(m + (size_t)(i + 6) * words + lane));
        // a level's survivors beyond its first max_keep can never be among the image's first max_keep: flag them too
        if (((mine >> (i & 63)) & 1ull) || kc >= max_keep) {
            if (lane == 0) rem_out[list[i]] = 1;
            continue;
        }
        ++kc;
        if (lane >= w && lane < words) rem |= m[(size_t)i * words + lane];
    }
}

__global__ void __launch_bounds__(kNmsWarps * 32)
nms_collect_kernel(const uint8_t *__restrict__ valid, const uint8_t *__restrict__ removed, int n, int K, int max_keep,
                   int *__restrict__ keep, int *__restrict__ count) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x * kNmsWarps + warp;
    if (img >= n) return;
    const uint8_t *v = valid + (size_t)img * K, *r = removed + (size_t)img * K;
    int *out = keep + (size_t)img * max_keep;
    int kc = 0;
    for (int base = 0; base < K && kc < max_keep; base += 32) {
        const int i = base + lane;
        const bool ok = i < K && v[i] != 0 && r[i] == 0;
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        const int pos = kc + __popc(b & ((1u << lane) - 1u));
        if (ok && pos < max_keep) out[pos] = i;
        kc += __popc(b);
    }
    kc = min(kc, max_keep);
    for (int j = kc + lane; j < max_keep; j += 32) out[j] = -1;
    if (lane == 0) count[img] = kc;
}

struct LevelNmsLayout {
    size_t lists, counts, removed, mask, total;
    int words;
};
LevelNmsLayout level_nms_layout(int n, int K, int L, int Kl) {
    LevelNmsLayout o;
    o.words = (Kl + 63) / 64;
    o.lists = 0;
    o.counts = align_up(o.lists + (size_t)n * L * Kl * sizeof(int), 256);
    o.removed = align_up(o.counts + (size_t)n * L * sizeof(int), 256);
    o.mask = align_up(o.removed + (size_t)n * K, 256);
    o.total = o.mask + (size_t)n * L * Kl * o.words * sizeof(unsigned long long);
    return o;
}

}  // namespace
}  // namespace msq

extern "C" size_t msq_nms_levels_scratch_bytes(int n, int K, int n_levels, int max_per_level) {
    if (n <= 0 || K <= 0 || n_levels <= 0 || max_per_level <= 0) return 0;
    return msq::level_nms_layout(n, K, n_levels, std::min(K, max_per_level)).total;
}

extern "C" int msq_nms_levels_long(const float *boxes, const uint8_t *level_valid, int n, int K, int n_levels, int max_per_level,
                                   float iou_threshold, int max_keep, int32_t *keep, int32_t *count, void *scratch, size_t scratch_bytes,
                                   void *stream) {
    using namespace msq;
    MSQ_REQUIRE(n >= 0 && K >= 0 && max_keep > 0 && max_per_level > 0, MSQ_EINVAL, "msq_nms_levels_long: bad sizes n=%d K=%d max_keep=%d", n, K, max_keep);
    MSQ_REQUIRE(n_levels >= 1 && n_levels <= kNmsMaxLevels, MSQ_EINVAL, "msq_nms_levels_long: 1..%d levels (got %d)", kNmsMaxLevels, n_levels);
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(boxes && level_valid && keep && count && scratch, MSQ_EINVAL, "msq_nms_levels_long: null pointer");
    MSQ_REQUIRE((uintptr_t)boxes % 16 == 0 && (uintptr_t)scratch % 256 == 0, MSQ_EINVAL, "msq_nms_levels_long: boxes must be 16-byte, scratch 256-byte aligned");
    const int Kl = std::min(K, max_per_level);
    MSQ_REQUIRE(Kl <= 2048, MSQ_EUNSUPPORTED, "msq_nms_levels_long: at most 2048 candidates per level (got %d)", Kl);
    const LevelNmsLayout lay = level_nms_layout(n, K, n_levels, Kl);
    MSQ_REQUIRE(scratch_bytes >= lay.total, MSQ_ENOMEM, "msq_nms_levels_long: scratch too small (%zu < %zu)", scratch_bytes, lay.total);
    char *base = static_cast<char *>(scratch);
    int *lists = reinterpret_cast<int *>(base + lay.lists), *counts = reinterpret_cast<int *>(base + lay.counts);
    uint8_t *removed = reinterpret_cast<uint8_t *>(base + lay.removed);
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(base + lay.mask);
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 0) {
        MSQ_CUDA_OK(cudaMemsetAsync(keep, 0xff, (size_t)n * max_keep * sizeof(int32_t), st));
        MSQ_CUDA_OK(cudaMemsetAsync(count, 0, (size_t)n * sizeof(int32_t), st));
        return MSQ_OK;
    }
    MSQ_CUDA_OK(cudaMemsetAsync(removed, 0, (size_t)n * K, st));
    TimedLaunch timed(K_DETECTOR_GLUE, st, 4);
    nms_level_lists_kernel<<<n, 32 * kNmsMaxLevels, 0, st>>>(level_valid, K, n_levels, Kl, lists, counts, removed);
    nms_level_mask_kernel<<<dim3(lay.words, lay.words, n * n_levels), 64, 0, st>>>(reinterpret_cast<const float4 *>(boxes), K, n_levels, Kl,
                                                                                 lay.words, iou_threshold, lists, counts, mask);
    const int n_il = n * n_levels;
    nms_level_scan_kernel<<<(n_il + kNmsWarps - 1) / kNmsWarps, kNmsWarps * 32, 0, st>>>(mask, n_il, K, n_levels, Kl, lay.words, max_keep,
                                                                                        lists, counts, removed);
    nms_collect_kernel<<<(n + kNmsWarps - 1) / kNmsWarps, kNmsWarps * 32, 0, st>>>(level_valid, removed, n, K, max_keep, keep, count);
    MSQ_LAUNCH_OK("nms_levels_long");
    return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Keypoint decoding for a whole batch: torchvision's heatmaps_to_keypoints (the reference's detectron2 keypoint head does
// the same) resizes every RoI's K heatmaps to the RoI's size with bicubic interpolation and takes the arg-max -- one
// F.interpolate + arg-max + two host synchronisations per RoI.  Here one CTA per (RoI, keypoint) holds the heatmap in
// shared memory, evaluates PyTorch's bicubic kernel (A = -0.75, align_corners = False, taps clamped to the map) at every
// pixel of the virtual resized map and reduces to the first maximum.  With `round_bf16` the values are rounded to bfloat16
// before they are compared, which is what the resize returns under autocast.
// ---------------------------------------------------------------------------------------------------------------
namespace msq {
namespace {

__device__ __forceinline__ float cubic1(float x) { return ((-0.75f + 2.f) * x - (-0.75f + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x) { return ((-0.75f * x - 5.f * -0.75f) * x + 8.f * -0.75f) * x - 4.f * -0.75f; }
__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
    c[0] = cubic2(t + 1.f);
    c[1] = cubic1(t);
    const float u = 1.f - t;
    c[2] = cubic1(u);
    c[3] = cubic2(u + 1.f);
}
__device__ __forceinline__ float round_to_bf16(float v) {
    uint32_t u = __float_as_uint(v);
    if ((u & 0x7f800000u) == 0x7f800000u) return v;                    // inf / nan unchanged
    u += 0x7fffu + ((u >> 16) & 1u);                                    // round to nearest even
    return __uint_as_float(u & 0xffff0000u);
}

constexpr int kKpThreads = 256;

__global__ void __launch_bounds__(kKpThreads)
keypoint_decode_kernel(const float *__restrict__ maps, const float *__restrict__ rois, int K, int Hm, int Wm, int round_bf16,
                       float *__restrict__ xyv, float *__restrict__ scores) {
    extern __shared__ float heat[];
    __shared__ float best_v[kKpThreads / 32];
    __shared__ int best_i[kKpThreads / 32];
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
        cubic_coeffs(rx - fx, cx);
        cubic_coeffs(ry - fy, cy);
        float row[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float *r = heat + min(max(iy - 1 + a, 0), Hm - 1) * Wm;
            row[a] = r[min(max(ix - 1, 0), Wm - 1)] * cx[0] + r[min(max(ix, 0), Wm - 1)] * cx[1] +
                     r[min(max(ix + 1, 0), Wm - 1)] * cx[2] + r[min(max(ix + 2, 0), Wm - 1)] * cx[3];
        }
        float v = row[0] * cy[0] + row[1] * cy[1] + row[2] * cy[2] + row[3] * cy[3];
        if (round_bf16) v = round_to_bf16(v);
        if (v > bv) { bv = v; bi = p; }                                 // p ascends per thread: the first maximum stays
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
        if (bi == INT_MAX) bi = 0;                                      // all NaN: torch returns index 0 as well
        const int yi = bi / ow, xi = bi - yi * ow;
        float *o = xyv + ((size_t)roi * K + kp) * 3;
        o[0] = ((float)xi + 0.5f) * (width / (float)ow) + x1;
        o[1] = ((float)yi + 0.5f) * (height / (float)oh) + y1;
        o[2] = 1.f;
        scores[(size_t)roi * K + kp] = bv;
    }
}

}  // namespace
}  // namespace msq

extern "C" int msq_keypoints_from_heatmaps(const float *maps, const float *rois, int n_rois, int K, int Hm, int Wm, int round_bf16,
                                           float *xyv, float *scores, void *stream) {
    MSQ_REQUIRE(n_rois >= 0 && K > 0 && Hm > 0 && Wm > 0, MSQ_EINVAL, "msq_keypoints_from_heatmaps: bad sizes");
    if (n_rois == 0) return MSQ_OK;
    MSQ_REQUIRE(maps && rois && xyv && scores, MSQ_EINVAL, "msq_keypoints_from_heatmaps: null pointer");
    const size_t smem = (size_t)Hm * Wm * sizeof(float);
    MSQ_REQUIRE(smem <= 200 * 1024, MSQ_EUNSUPPORTED, "msq_keypoints_from_heatmaps: %dx%d heatmaps are too large", Hm, Wm);
    if (smem > 48 * 1024)
        MSQ_CUDA_OK(cudaFuncSetAttribute(msq::keypoint_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    msq::TimedLaunch timed(msq::K_DETECTOR_GLUE, (cudaStream_t)stream);
    msq::keypoint_decode_kernel<<<n_rois * K, msq::kKpThreads, smem, (cudaStream_t)stream>>>(maps, rois, K, Hm, Wm, round_bf16, xyv, scores);
    MSQ_LAUNCH_OK("keypoint_decode");
    return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// detectron2 find_top_rpn_proposals up to (not including) the NMS, for a whole batch in ONE launch (model/ops.py ran it as
// ~40 PyTorch operators: per-level top-k = a full sort of every row, gathers, decode, clip, a second sort: 3 ms per 500
// images).  A CTA owns one image:
//   per pyramid level   objectness logits -> order-preserving 32-bit keys in shared memory; levels with more anchors than
//                       PRE_NMS_TOPK get an exact k-th largest by 4-pass radix select (warp-aggregated histogram votes), ties at
//                       the threshold are taken in anchor order; every selected anchor is decoded on the spot
//                       (Box2BoxTransform, weights 1) and clipped to the image;
//   across levels       candidates are sorted by descending logit with a bitonic network on 64-bit words
//                       (key | reversed (level, anchor) order | slot): invalid boxes (not finite, empty after clipping) sink
//                       to the end, equal logits keep (level, anchor) order;
//   output              boxes, logits, validity and the boxes shifted per level by (largest coordinate + 1) -- torchvision's
//                       batched_nms coordinate trick -- in sorted order, ready for msq_nms_sorted / msq_nms_sorted_long.
// ---------------------------------------------------------------------------------------------------------------
namespace msq {
namespace {

constexpr int kSelThreads = 512;
constexpr int kSelMaxLevels = 8;
constexpr int kSelMaxCand = 4096;

struct RpnLevelsArg {
    const void *pred[kSelMaxLevels];          // (n, H, W, 16) channels-last: 3 logits, 12 deltas, 1 padding channel
    int H[kSelMaxLevels], W[kSelMaxLevels], stride[kSelMaxLevels], take[kSelMaxLevels], offset[kSelMaxLevels];
    float cell[kSelMaxLevels][3][4];          // the three cell anchors of the level (x1, y1, x2, y2 around 0)
    int n_levels, total;                      // candidates per image = sum of take[]
};

__device__ __forceinline__ uint32_t float_key(float x) {                 // larger float <=> larger key; NaN lowest
    if (x != x) return 0u;
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(kSelThreads)
rpn_select_kernel(RpnLevelsArg L, float img_w, float img_h, float scale_clamp, int sort_len,
                  float4 *__restrict__ boxes_out, float4 *__restrict__ shifted_out, float *__restrict__ scores_out,
                  uint8_t *__restrict__ valid_out) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned long long *sort_keys = reinterpret_cast<unsigned long long *>(sel_smem);                   // [sort_len]
    float4 *cand_box = reinterpret_cast<float4 *>(sort_keys + sort_len);                                // [total]
    uint32_t *keys = reinterpret_cast<uint32_t *>(cand_box + L.total);                                   // [largest level]
    __shared__ int hist[256];
    __shared__ int counts[kSelThreads];
    __shared__ uint32_t s_prefix;
    __shared__ int s_remaining, s_greater;
    __shared__ float s_max[kSelThreads / 32];
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31;

    for (int lvl = 0; lvl < L.n_levels; ++lvl) {
        const int H = L.H[lvl], W = L.W[lvl], N = H * W * 3, take = L.take[lvl];
        const T *pred = static_cast<const T *>(L.pred[lvl]) + (size_t)img * H * W * 16;
        for (int i = tid; i < N; i += kSelThreads) keys[i] = float_key(to_f32(pred[(size_t)(i / 3) * 16 + (i % 3)]));
        __syncthreads();
        uint32_t thr = 0u;
        int take_equal = N;                                   // anchors equal to the threshold that are still taken
        if (take < N) {
            // ---- exact k-th largest key: most significant byte first ----
            uint32_t prefix = 0u, mask = 0u;
            int remaining = take;
            for (int pass = 0; pass < 4; ++pass) {
                const int shift = 24 - 8 * pass;
                for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
                __syncthreads();
                for (int i0 = 0; i0 < N; i0 += kSelThreads) {
                    const int i = i0 + tid;
                    int bin = 256;                            // lanes without a vote agree among themselves
                    if (i < N && (keys[i] & mask) == prefix) bin = (keys[i] >> shift) & 255;
                    const unsigned peers = __match_any_sync(0xffffffffu, bin);
                    if (bin < 256 && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
                }
                __syncthreads();
                if (tid == 0) {
                    int above = 0, b = 255;
                    for (; b > 0; --b) {
                        if (above + hist[b] >= remaining) break;
                        above += hist[b];
                    }
                    s_prefix = prefix | ((uint32_t)b << shift);
                    s_remaining = remaining - above;
                }
                __syncthreads();
                prefix = s_prefix; remaining = s_remaining; mask |= 0xffu << shift;
                __syncthreads();
            }
            thr = prefix;
            take_equal = remaining;
        }
        // ---- anchors above the threshold, then the first `take_equal` at the threshold in anchor order ----
        const int per = (N + kSelThreads - 1) / kSelThreads, i_lo = min(N, tid * per), i_hi = min(N, i_lo + per);
        int n_gt = 0, n_eq = 0;
        for (int i = i_lo; i < i_hi; ++i) { n_gt += keys[i] > thr; n_eq += keys[i] == thr; }
        counts[tid] = n_gt;
        __syncthreads();
        if (tid == 0) { int run = 0; for (int t = 0; t < kSelThreads; ++t) { const int c = counts[t]; counts[t] = run; run += c; } s_greater = run; }
        __syncthreads();
        int pos_gt = counts[tid];
        const int greater_total = s_greater;
        __syncthreads();
        counts[tid] = n_eq;
        __syncthreads();
        if (tid == 0) { int run = 0; for (int t = 0; t < kSelThreads; ++t) { const int c = counts[t]; counts[t] = run; run += c; } }
        __syncthreads();
        int pos_eq = counts[tid];
        __syncthreads();
        const int base = L.offset[lvl];
        const float stride = (float)L.stride[lvl];
        for (int i = i_lo; i < i_hi; ++i) {
            const uint32_t k = keys[i];
            int slot = -1;
            if (k > thr) slot = pos_gt++;
            else if (k == thr) { if (pos_eq < take_equal) slot = greater_total + pos_eq; ++pos_eq; }
            if (slot < 0 || slot >= take) continue;
            // decode anchor i = (y * W + x) * 3 + a with its deltas (detectron2 Box2BoxTransform.apply_deltas, weights 1)
            const int a = i % 3, pix = i / 3, y = pix / W, x = pix - y * W;
            const float sx = (float)x * stride, sy = (float)y * stride;
            const float ax1 = sx + L.cell[lvl][a][0], ay1 = sy + L.cell[lvl][a][1], ax2 = sx + L.cell[lvl][a][2], ay2 = sy + L.cell[lvl][a][3];
            const T *d = pred + (size_t)pix * 16 + 3 + 4 * a;
            const float dx = to_f32(d[0]), dy = to_f32(d[1]), dw = fminf(to_f32(d[2]), scale_clamp), dh = fminf(to_f32(d[3]), scale_clamp);
            const float wa = ax2 - ax1, ha = ay2 - ay1, cx = ax1 + 0.5f * wa, cy = ay1 + 0.5f * ha;
            const float pcx = dx * wa + cx, pcy = dy * ha + cy, pw = expf(dw) * wa, ph = expf(dh) * ha;
            float4 b = make_float4(pcx - 0.5f * pw, pcy - 0.5f * ph, pcx + 0.5f * pw, pcy + 0.5f * ph);
            const bool finite = isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w) && k != 0u && isfinite(key_float(k));
            b.x = fminf(fmaxf(b.x, 0.f), img_w); b.y = fminf(fmaxf(b.y, 0.f), img_h);
            b.z = fminf(fmaxf(b.z, 0.f), img_w); b.w = fminf(fmaxf(b.w, 0.f), img_h);
            const bool ok = finite && (b.z - b.x) > 0.f && (b.w - b.y) > 0.f;
            const int c = base + slot;
            cand_box[c] = b;
            // descending sort: valid boxes by logit, then (level, anchor) ascending; invalid ones last
            const unsigned long long order = 0xfffffull - (((unsigned long long)lvl << 14) | (unsigned long long)i);
            sort_keys[c] = ((unsigned long long)(ok ? k : 0u) << 32) | (order << 12) | (unsigned long long)c;
        }
        __syncthreads();
    }
    for (int i = L.total + tid; i < sort_len; i += kSelThreads) sort_keys[i] = 0ull;       // padding sinks below everything
    __syncthreads();
    // ---- bitonic sort, descending ----
    for (int k = 2; k <= sort_len; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < sort_len; i += kSelThreads) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = sort_keys[i], b = sort_keys[p];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { sort_keys[i] = b; sort_keys[p] = a; }
                }
            }
            __syncthreads();
        }
    }
    // ---- largest coordinate of the image's valid boxes (torchvision batched_nms offsets) ----
    float mx = -INFINITY;
    for (int i = tid; i < L.total; i += kSelThreads) {
        const unsigned long long w = sort_keys[i];
        if ((w >> 32) != 0ull) {
            const float4 b = cand_box[(int)(w & 0xfffull)];
            mx = fmaxf(fmaxf(mx, fmaxf(b.x, b.y)), fmaxf(b.z, b.w));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_max[tid >> 5] = mx;
    __syncthreads();
    mx = s_max[0];
    for (int wv = 1; wv < kSelThreads / 32; ++wv) mx = fmaxf(mx, s_max[wv]);
    if (!isfinite(mx)) mx = 0.f;
    const size_t o0 = (size_t)img * L.total;
    for (int i = tid; i < L.total; i += kSelThreads) {
        const unsigned long long w = sort_keys[i];
        const uint32_t k = (uint32_t)(w >> 32);
        const int c = (int)(w & 0xfffull);
        const int lvl = (int)((0xfffffull - ((w >> 12) & 0xfffffull)) >> 14);
        const float4 b = cand_box[c];
        const float off = (float)lvl * (mx + 1.f);
        boxes_out[o0 + i] = b;
        shifted_out[o0 + i] = make_float4(b.x + off, b.y + off, b.z + off, b.w + off);
        scores_out[o0 + i] = k ? key_float(k) : -INFINITY;
        valid_out[o0 + i] = k ? (uint8_t)(1 + lvl) : 0;               // non-zero = valid; the value carries the pyramid level
    }
}

}  // namespace
}  // namespace msq

extern "C" int msq_rpn_select(const void *const *pred_dev, const int *heights, const int *widths, const int *strides, const float *cell_anchors,
                              int n_levels, int is_bf16, int n, int pre_topk, int img_h, int img_w, float *boxes_dev, float *shifted_dev,
                              float *scores_dev, uint8_t *valid_dev, void *stream) {
    MSQ_REQUIRE(n_levels >= 1 && n_levels <= kSelMaxLevels && n >= 0 && pre_topk >= 1, MSQ_EINVAL, "msq_rpn_select: bad sizes");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(pred_dev && heights && widths && strides && cell_anchors && boxes_dev && shifted_dev && scores_dev && valid_dev, MSQ_EINVAL,
                "msq_rpn_select: null pointer");
    RpnLevelsArg L;
    L.n_levels = n_levels;
    int total = 0, largest = 0;
    for (int l = 0; l < n_levels; ++l) {
        MSQ_REQUIRE(pred_dev[l] && heights[l] > 0 && widths[l] > 0, MSQ_EINVAL, "msq_rpn_select: level %d", l);
        const int N = heights[l] * widths[l] * 3;
        MSQ_REQUIRE(N < (1 << 14), MSQ_EUNSUPPORTED, "msq_rpn_select: at most 16383 anchors per level (level %d has %d)", l, N);
        L.pred[l] = pred_dev[l]; L.H[l] = heights[l]; L.W[l] = widths[l]; L.stride[l] = strides[l];
        L.take[l] = std::min(pre_topk, N); L.offset[l] = total;
        for (int a = 0; a < 3; ++a) for (int k = 0; k < 4; ++k) L.cell[l][a][k] = cell_anchors[(l * 3 + a) * 4 + k];
        total += L.take[l];
        largest = std::max(largest, N);
    }
    MSQ_REQUIRE(total <= kSelMaxCand, MSQ_EUNSUPPORTED, "msq_rpn_select: at most %d candidates per image (got %d)", kSelMaxCand, total);
    L.total = total;
    int sort_len = 2;
    while (sort_len < total) sort_len <<= 1;
    const size_t smem = (size_t)sort_len * 8 + (size_t)total * 16 + (size_t)largest * 4;
    MSQ_REQUIRE(smem <= 200 * 1024, MSQ_EUNSUPPORTED, "msq_rpn_select: %zu bytes of shared memory needed", smem);
    MSQ_REQUIRE((uintptr_t)boxes_dev % 16 == 0 && (uintptr_t)shifted_dev % 16 == 0, MSQ_EINVAL, "msq_rpn_select: outputs must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    TimedLaunch timed(K_DETECTOR_GLUE, st);
    const float clamp = logf(1000.f / 16.f);
    if (is_bf16) {
        MSQ_CUDA_OK(cudaFuncSetAttribute(rpn_select_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rpn_select_kernel<__nv_bfloat16><<<n, kSelThreads, smem, st>>>(L, (float)img_w, (float)img_h, clamp, sort_len, reinterpret_cast<float4 *>(boxes_dev),
                                                                      reinterpret_cast<float4 *>(shifted_dev), scores_dev, valid_dev);
    } else {
        MSQ_CUDA_OK(cudaFuncSetAttribute(rpn_select_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rpn_select_kernel<float><<<n, kSelThreads, smem, st>>>(L, (float)img_w, (float)img_h, clamp, sort_len, reinterpret_cast<float4 *>(boxes_dev),
                                                              reinterpret_cast<float4 *>(shifted_dev), scores_dev, valid_dev);
    }
    MSQ_LAUNCH_OK("rpn_select");
    return MSQ_OK;
}
