// a4: the R-CNN's stem on the 5th-generation tensor cores.
//   prepared uint8 frame -> a3 scaling -> (x - mean) / std -> zero padding -> 7x7 stride-2 convolution (1 input channel: the grey
//   image the reference replicates three times, weights summed over those channels) + bias -> ReLU -> 3x3 stride-2 max-pool
// as ONE kernel writing the (n, 64, 64, 64) channels-last bf16 input of res2 (see stem_conv_pool_kernel in prep.cu for the
// float32 CUDA-core form; this one serves the bf16 graph).
//
// The convolution is a GEMM with M = output pixels, N = 64 channels, K = 49 taps padded to 64: exactly one 128-byte K block.
// A CTA owns 16 x 7 pooled pixels = 33 x 15 convolution outputs = 495 GEMM rows = four M = 128 tiles:
//   1. the 71 x 35 input patch goes through a 256-entry table (scaling + normalisation, bf16) into shared memory;
//   2. the im2col operand A (4 x 128 rows x 64 k, bf16) is WRITTEN by the threads in the swizzle-128B K-major layout tcgen05.mma
//      reads (16-byte chunk index XOR row-in-atom) -- no TMA involved, the source is a gather;
//   3. one thread issues 4 x 4 tcgen05.mma (M 128, N 64, K 16) into 256 TMEM columns and commits to an mbarrier;
//   4. eight warps read the accumulators back (tcgen05.ld), add the bias, apply ReLU and park the bf16 convolution outputs in the
//      shared memory that held A (the same XOR pattern keeps the 16-byte stores conflict-free);
//   5. 3x3 / 2 max-pool out of shared memory, 128-byte channel rows to global memory.
// The CUDA-core kernel is bound by shared-memory operand traffic (every FFMA needs a broadcast weight): 2.5 ms per 500 frames;
// here the operands never pass through registers.
#include "common.cuh"
#include <cuda_bf16.h>

namespace msq {
namespace {

constexpr int kPoolW = 16, kPoolH = 7;                 // pooled pixels per CTA
constexpr int kConvW = 2 * kPoolW + 1, kConvH = 2 * kPoolH + 1;      // 33 x 15 convolution outputs
constexpr int kRows = kConvW * kConvH;                 // 495 GEMM rows
constexpr int kMTiles = 4;                             // 4 x 128 >= 495
constexpr int kInW = 2 * (kConvW - 1) + 7, kInH = 2 * (kConvH - 1) + 7;   // 71 x 35 input pixels
constexpr int kInPitch = 72;
constexpr int kStemThreads = 256;
constexpr int kATileBytes = 128 * 128;                 // 16 KB per M tile
constexpr int kTmemCols = 256;                         // 4 tiles x 64 columns

__device__ __forceinline__ uint32_t sm_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kStemThreads, 2)
stem_tc_kernel(const uint8_t *__restrict__ in, int h, int w, int conv_h, int conv_w, int pool_h, int pool_w, int tiles_x, float mean,
               float stdv, double vmin, double vmax, int vmin_is_int, const uint4 *__restrict__ b_tile /* 8 KB, swizzled bf16 [64 n][64 k] */,
               const float *__restrict__ bias64, __nv_bfloat16 *__restrict__ out) {
    extern __shared__ uint8_t stem_raw[];
    const uint32_t raw = sm_u32(stem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gen = stem_raw + (base - raw);
    uint8_t *smA = gen;                                                    // 64 KB: im2col operand, later the convolution outputs
    uint8_t *smB = gen + kMTiles * kATileBytes;                            // 8 KB
    __nv_bfloat16 *patch = reinterpret_cast<__nv_bfloat16 *>(smB + 8192);  // [35][72]
    __nv_bfloat16 *lut = patch + kInH * kInPitch;                          // [256]
    uint8_t *tail = reinterpret_cast<uint8_t *>(lut + 256);
    const uint32_t bar = (sm_u32(tail) + 7u) & ~7u;                        // one mbarrier
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(tail + (bar - sm_u32(tail)) + 8);
    __shared__ uint8_t lut8[256];
    __shared__ int tap_off[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x, ty = blockIdx.y / tiles_x, tx = blockIdx.y - ty * tiles_x;
    const int py0 = ty * kPoolH, px0 = tx * kPoolW;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1, iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm_u32((const void *)tmem_slot)), "r"((uint32_t)kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- 1. tables, weights, input patch ----
    build_scale_lut(lut8, vmin, vmax, vmin_is_int);                         // (ends with a block barrier)
    lut[threadIdx.x] = __float2bfloat16_rn(((float)lut8[threadIdx.x] - mean) / stdv);
    if (threadIdx.x < 64) tap_off[threadIdx.x] = threadIdx.x < 49 ? (threadIdx.x / 7) * kInPitch + (threadIdx.x % 7) : -1;
    for (int i = threadIdx.x; i < 512; i += kStemThreads) reinterpret_cast<uint4 *>(smB)[i] = b_tile[i];
    __syncthreads();
    const uint8_t *src = in + (size_t)img * h * w;
    for (int i = threadIdx.x; i < kInH * kInPitch; i += kStemThreads) {
        const int j = i / kInPitch, k = i - j * kInPitch;
        const int gy = iy0 + j, gx = ix0 + k;
        patch[i] = (k < kInW && gy >= 0 && gy < h && gx >= 0 && gx < w) ? lut[src[(size_t)gy * w + gx]] : __float2bfloat16_rn(0.f);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // ---- 2. im2col in the swizzle-128B K-major layout: row m (128 bytes), 16-byte chunk c holds k = 8c .. 8c+7 ----
    for (int i = threadIdx.x; i < kMTiles * 128 * 8; i += kStemThreads) {
        const int m = i >> 3, c = i & 7;
        uint32_t pk[4] = {0u, 0u, 0u, 0u};
        if (m < kRows) {
            const int oy = m / kConvW, ox = m - oy * kConvW;
            const __nv_bfloat16 *p0 = patch + 2 * oy * kInPitch + 2 * ox;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int off = tap_off[8 * c + e];
                const uint32_t v = off >= 0 ? (uint32_t)__bfloat16_as_ushort(p0[off]) : 0u;
                pk[e >> 1] |= v << (16 * (e & 1));
            }
        }
        const int t = m >> 7, r = m & 127;
        *reinterpret_cast<uint4 *>(smA + t * kATileBytes + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy writes -> visible to the tensor core's reads
    __syncthreads();
    // ---- 3. 4 x 4 MMAs ----
    if (threadIdx.x == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t bdesc = (uint64_t)((sm_u32(smB) >> 4) & 0x3fffu) | hi;
#pragma unroll
        for (int t = 0; t < kMTiles; ++t) {
            const uint64_t adesc = (uint64_t)((sm_u32(smA + t * kATileBytes) >> 4) & 0x3fffu) | hi;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t accumulate = k ? 1u : 0u;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                             ::"r"(tmem_base + (uint32_t)(t * 64)), "l"(adesc + (uint64_t)(2 * k)), "l"(bdesc + (uint64_t)(2 * k)), "r"(idesc), "r"(accumulate) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    {   // everyone waits for the accumulators (and thereby for the tensor core to be done reading A)
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();                                                       // A may now be overwritten by anyone
    // ---- 4. accumulators -> bias, ReLU, bf16 -> shared memory (row m, chunk c ^ (m & 7)) ----
    {
        const int q = warp & 3;
        for (int t = warp >> 2; t < kMTiles; t += 2) {
            const int m = t * 128 + q * 32 + lane;
            const int oy = m / kConvW, ox = m - oy * kConvW;
            const int gy = cy0 + oy, gx = cx0 + ox;
            const bool inside = m < kRows && gy >= 0 && gy < conv_h && gx >= 0 && gx < conv_w;    // outside the map counts as 0: exact for max after ReLU
#pragma unroll
            for (int cc = 0; cc < 64; cc += 32) {
                uint32_t v[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                               "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                               "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 64 + cc)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int ch = cc + 8 * g + 2 * k;
                        const float a = inside ? fmaxf(__uint_as_float(v[8 * g + 2 * k]) + bias64[ch], 0.f) : 0.f;
                        const float b = inside ? fmaxf(__uint_as_float(v[8 * g + 2 * k + 1]) + bias64[ch + 1], 0.f) : 0.f;
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                        o[k] = *reinterpret_cast<const uint32_t *>(&h2);
                    }
                    const int c = (cc >> 3) + g;
                    *reinterpret_cast<uint4 *>(smA + m * 128 + ((c ^ (m & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
    }
    // ---- 5. 3x3 stride-2 max-pool, 8 channels per thread ----
    for (int i = threadIdx.x; i < kPoolW * kPoolH * 8; i += kStemThreads) {
        const int c = i & 7, pp = i >> 3, py = pp / kPoolW, px = pp - py * kPoolW;
        const int gy = py0 + py, gx = px0 + px;
        if (gy >= pool_h || gx >= pool_w) continue;
        __nv_bfloat162 mx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) mx[k] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int m = (2 * py + dy) * kConvW + 2 * px + dx;
                const uint4 v = *reinterpret_cast<const uint4 *>(smA + m * 128 + ((c ^ (m & 7)) << 4));
                const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) mx[k] = __hmax2(mx[k], *reinterpret_cast<const __nv_bfloat162 *>(&vw[k]));
            }
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = *reinterpret_cast<const uint32_t *>(&mx[k]);
        *reinterpret_cast<uint4 *>(out + (((size_t)img * pool_h + gy) * pool_w + gx) * 64 + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" int msq_stem_conv_pool_tc(const uint8_t *in, int n, int h, int w, int ph, int pw, double vmin, double vmax, int vmin_is_int,
                                     float mean, float stdv, const void *b_tile, const float *bias64, void *out, void *stream) {
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0 && ph >= h && pw >= w, MSQ_EINVAL, "msq_stem_conv_pool_tc: bad sizes n=%d %dx%d in %dx%d", n, h, w, ph, pw);
    MSQ_REQUIRE(vmax != vmin && stdv != 0.f, MSQ_EINVAL, "msq_stem_conv_pool_tc: vmax == vmin or std == 0");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(in && b_tile && bias64 && out, MSQ_EINVAL, "msq_stem_conv_pool_tc: null pointer");
    MSQ_REQUIRE((uintptr_t)out % 16 == 0 && (uintptr_t)b_tile % 16 == 0, MSQ_EINVAL, "msq_stem_conv_pool_tc: pointers must be 16-byte aligned");
    const int conv_h = (ph + 6 - 7) / 2 + 1, conv_w = (pw + 6 - 7) / 2 + 1;
    const int pool_h = (conv_h + 2 - 3) / 2 + 1, pool_w = (conv_w + 2 - 3) / 2 + 1;
    const int tiles_x = (pool_w + kPoolW - 1) / kPoolW, tiles_y = (pool_h + kPoolH - 1) / kPoolH;
    const size_t smem = 1024 + (size_t)kMTiles * kATileBytes + 8192 + (size_t)kInH * kInPitch * 2 + 512 + 64;
    cudaStream_t st = (cudaStream_t)stream;
    MSQ_CUDA_OK(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TimedLaunch timed(K_DETECTOR_GLUE, st);
    stem_tc_kernel<<<dim3(n, tiles_x * tiles_y), kStemThreads, smem, st>>>(in, h, w, conv_h, conv_w, pool_h, pool_w, tiles_x, mean, stdv, vmin, vmax,
                                                                           vmin_is_int, static_cast<const uint4 *>(b_tile), bias64,
                                                                           static_cast<__nv_bfloat16 *>(out));
    MSQ_LAUNCH_OK("stem_conv_pool_tc");
    return MSQ_OK;
}
