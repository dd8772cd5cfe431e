// a4 (R-CNN glue): segmented greedy NMS for the proposal filtering of a batch of images.
// torchvision's RegionProposalNetwork.filter_proposals runs one NMS per image (a kernel that builds an N x N suppression
// mask, a copy of it to the host and a sequential scan there): ~0.6 ms of launch / synchronisation latency per frame.
// Here ONE launch serves the whole batch: a warp per image walks that image's candidates in score order and keeps a box
// unless a box it kept earlier overlaps it by more than the threshold -- the same greedy rule -- and stops after
// `max_keep` boxes (the RPN only wants the first 100 survivors, which are found within the first few hundred candidates).
// Boxes arrive sorted by score and already shifted per pyramid level (torchvision's "coordinate trick", so that levels
// never suppress each other and the float32 arithmetic is the one torchvision does); IoU as in torchvision's devIoU.
#include "common.cuh"

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
    TimedLaunch timed(K_PASTE, (cudaStream_t)stream);          // counted with the other R-CNN glue kernel
    nms_sorted_kernel<<<(n + kNmsWarps - 1) / kNmsWarps, kNmsWarps * 32, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(boxes), valid, n, K, iou_threshold, max_keep, keep, count);
    MSQ_LAUNCH_OK("nms_sorted");
    return MSQ_OK;
}
