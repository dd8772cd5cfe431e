// f4 (session setup): get_bground_im -- per-frame K x K median blur, then the per-pixel temporal median.
//   ref proc/roi.py:293-307: frames[i] = cv2.medianBlur(frames[i], med_scale); bground = np.median(frames, axis=0)
//   (io/session.py:217-218 feeds it every 500th frame of the session: ~216 frames of 512x424 for an hour of video).
// Both steps are exact order statistics, so the result is bit-identical to OpenCV + NumPy:
//   * cv2.medianBlur on 16-bit images supports K = 3 and 5, replicates the border and returns the middle of the K*K
//     sorted values; here one thread per pixel sorts its window with a fully unrolled bitonic network in registers
//     (padded with +inf to 16 / 32 entries, all indices compile-time constants);
//   * np.median over the frame axis is the middle value for an odd count and the float64 mean of the two middle values
//     for an even one; here a CTA parks the n values of 32..128 neighbouring pixels in shared memory as order-preserving
//     16-bit keys (coalesced loads, column = pixel) and each thread runs a 16-step radix select down its own column.
#include "common.cuh"
#include <algorithm>
#include <limits.h>

namespace msq {
namespace {

template <int N>
__device__ __forceinline__ void bitonic_sort(int (&a)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const int lo = min(a[i], a[l]), hi = max(a[i], a[l]);
                    if ((i & k) == 0) { a[i] = lo; a[l] = hi; } else { a[i] = hi; a[l] = lo; }
                }
            }
}

template <int K, int NPAD>
__global__ void __launch_bounds__(256)
median_blur16_kernel(const uint16_t *__restrict__ in, uint16_t *__restrict__ out, int n, int H, int W, int is_unsigned) {
    const size_t plane = (size_t)H * W;
    const size_t total = (size_t)n * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / plane;
        const int p = (int)(i - f * plane), y = p / W, x = p - y * W;
        const uint16_t *src = in + f * plane;
        int v[NPAD];
#pragma unroll
        for (int q = 0; q < NPAD; ++q) v[q] = INT_MAX;
#pragma unroll
        for (int dy = 0; dy < K; ++dy) {
            const int yy = min(max(y + dy - K / 2, 0), H - 1);          // BORDER_REPLICATE
#pragma unroll
            for (int dx = 0; dx < K; ++dx) {
                const int xx = min(max(x + dx - K / 2, 0), W - 1);
                const uint16_t raw = __ldg(src + (size_t)yy * W + xx);
                v[dy * K + dx] = is_unsigned ? (int)raw : (int)(int16_t)raw;
            }
        }
        bitonic_sort<NPAD>(v);
        out[i] = (uint16_t)v[(K * K) / 2];
    }
}

// keys[f][t]: order-preserving unsigned 16-bit key of frame f at pixel (pix0 + t)
template <int T>
__global__ void __launch_bounds__(T)
temporal_median_kernel(const uint16_t *__restrict__ frames, int n, size_t plane, int is_unsigned, double *__restrict__ out) {
    extern __shared__ __align__(16) uint16_t keys[];
    const size_t pix = (size_t)blockIdx.x * T + threadIdx.x;
    const uint16_t flip = is_unsigned ? 0u : 0x8000u;
    if (pix < plane)
        for (int f = 0; f < n; ++f) keys[(size_t)f * T + threadIdx.x] = __ldg(frames + (size_t)f * plane + pix) ^ flip;
    if (pix >= plane) return;                       // no barrier below: every thread only reads its own column
    const uint16_t *col = keys + threadIdx.x;
    const int k_lo = (n - 1) / 2, k_hi = n / 2;     // equal for odd n
    // radix select of the k_lo-th smallest key
    uint32_t prefix = 0u, mask = 0u;
    int k = k_lo;
    for (int bit = 15; bit >= 0; --bit) {
        const uint32_t m2 = mask | (1u << bit);
        int zeros = 0;
        for (int f = 0; f < n; ++f) zeros += ((uint32_t)col[(size_t)f * T] & m2) == prefix;
        if (k >= zeros) { k -= zeros; prefix |= 1u << bit; }
        mask = m2;
    }
    uint32_t second = prefix;
    if (k_hi != k_lo) {
        int le = 0;
        uint32_t next = 0xffffffffu;
        for (int f = 0; f < n; ++f) {
            const uint32_t key = col[(size_t)f * T];
            le += key <= prefix;
            if (key > prefix) next = min(next, key);
        }
        if (le <= k_hi) second = next;              // the (k_hi)-th value is the next distinct one
    }
    const int a = is_unsigned ? (int)prefix : (int)(int16_t)(uint16_t)(prefix ^ 0x8000u);
    const int b = is_unsigned ? (int)second : (int)(int16_t)(uint16_t)(second ^ 0x8000u);
    out[pix] = (k_hi == k_lo) ? (double)a : ((double)a + (double)b) / 2.0;
}

template <int T>
int launch_temporal(const uint16_t *frames, int n, size_t plane, int is_unsigned, double *out, cudaStream_t st) {
    const size_t smem = (size_t)n * T * sizeof(uint16_t);
    if (smem > 48 * 1024)
        MSQ_CUDA_OK(cudaFuncSetAttribute(temporal_median_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    temporal_median_kernel<T><<<(unsigned)((plane + T - 1) / T), T, smem, st>>>(frames, n, plane, is_unsigned, out);
    MSQ_LAUNCH_OK("temporal_median");
    return MSQ_OK;
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" size_t msq_bground_scratch_bytes(int n, int H, int W) {
    if (n <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)n * H * W * sizeof(uint16_t);
}

extern "C" int msq_get_bground_im(const void *frames, int n, int H, int W, int med_scale, int is_unsigned, double *out,
                                  void *scratch, size_t scratch_bytes, void *stream) {
    MSQ_REQUIRE(n >= 1 && H > 0 && W > 0, MSQ_EINVAL, "msq_get_bground_im: bad sizes n=%d H=%d W=%d", n, H, W);
    MSQ_REQUIRE(frames && out, MSQ_EINVAL, "msq_get_bground_im: null pointer");
    // cv2.medianBlur accepts 16-bit images only for ksize 3 and 5; 1 = no blur (not an OpenCV value, used by tests)
    MSQ_REQUIRE(med_scale == 1 || med_scale == 3 || med_scale == 5, MSQ_EUNSUPPORTED,
                "msq_get_bground_im: med_scale must be 3 or 5 for 16-bit frames (got %d)", med_scale);
    const size_t limit = 200 * 1024;
    MSQ_REQUIRE((size_t)n * 32 * sizeof(uint16_t) <= limit, MSQ_EUNSUPPORTED, "msq_get_bground_im: at most %zu frames (got %d)",
                limit / 64, n);
    cudaStream_t st = (cudaStream_t)stream;
    const uint16_t *src = static_cast<const uint16_t *>(frames);
    const size_t plane = (size_t)H * W;
    if (med_scale > 1) {
        MSQ_REQUIRE(scratch && scratch_bytes >= msq_bground_scratch_bytes(n, H, W) && (uintptr_t)scratch % 2 == 0, MSQ_ENOMEM,
                    "msq_get_bground_im: scratch must hold %zu bytes", msq_bground_scratch_bytes(n, H, W));
        uint16_t *blur = static_cast<uint16_t *>(scratch);
        const size_t total = (size_t)n * plane;
        const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
        TimedLaunch timed(K_BGROUND, st);
        if (med_scale == 5) median_blur16_kernel<5, 32><<<blocks, 256, 0, st>>>(src, blur, n, H, W, is_unsigned);
        else median_blur16_kernel<3, 16><<<blocks, 256, 0, st>>>(src, blur, n, H, W, is_unsigned);
        MSQ_LAUNCH_OK("median_blur16");
        src = blur;
    }
    TimedLaunch timed(K_BGROUND, st);
    if ((size_t)n * 128 * 2 <= limit) return launch_temporal<128>(src, n, plane, is_unsigned, out, st);
    if ((size_t)n * 64 * 2 <= limit) return launch_temporal<64>(src, n, plane, is_unsigned, out, st);
    return launch_temporal<32>(src, n, plane, is_unsigned, out, st);
}
