// a4: the R-CNN's 1x1 / 3x3 convolutions (and Linear layers) as a Blackwell-native implicit GEMM.
//
//   out[pixel, co] = act( sum_{tap, ci} x[pixel + tap, ci] * w[co, tap, ci] + bias[co] (+ residual[pixel, co]) )
//
// Activations are channels-last bf16, so a tile of 128 output pixels x 64 input channels is a K-major 128 x 64 operand whose
// rows are 128 bytes: exactly what TMA writes into shared memory with the 128-byte swizzle and what tcgen05.mma reads.  A
// 3x3 convolution is nine such GEMMs accumulated into the same TMEM tile: for tap (r, s) the SAME 4-D tensor map
// (C, W, H, N) is read at the box origin shifted by (s - 1, r - 1) -- rows and columns that fall outside the image are
// zero-filled by the TMA unit, which IS the convolution's zero padding, so there is no im2col buffer and no halo logic.  A
// stride-2 1x1 convolution reads a tensor map whose W / H strides are doubled.  The pixel box (bw, bh, bn) of a tile is
// chosen on the host so that bw * bh * bn = 128 for every map size of the graph (64x64 ... 4x4, 14x14 and 7x7 RoI maps, or
// 128 rows of a plain matrix for Linear layers).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   TMA producer: A (pixels x 64 channels) and B (BN output channels x 64) tiles into a ring of shared-memory stages,
//            completion on mbarriers (cp.async.bulk.tensor + mbarrier::complete_tx)
//   warp 1   allocates TMEM (2 x BN columns: two accumulator tiles) and issues tcgen05.mma (M = 128, N = BN, K = 16 per
//            instruction, fp32 accumulation in TMEM); tcgen05.commit frees the stage / publishes the accumulator
//   warps 2-9  epilogue: tcgen05.ld their 32 TMEM lanes (= 32 pixels) x half of the columns, add bias and residual (requested one
//            column group ahead), ReLU, convert to bf16 and store the pixel's channels -- while warp 1 already accumulates the
//            next tile into the other half of TMEM.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace msq {
namespace {

constexpr int kTcM = 128;                      // pixels per tile (UMMA M, cta_group::1)
constexpr int kTcK = 64;                       // bf16 per K block: 128 bytes = one swizzle-128B row
constexpr int kTcEpiWarps = 8;                 // two warps per TMEM lane quarter, each takes half of the tile's columns
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;
constexpr int kTcABytes = kTcM * kTcK * 2;     // 16 KB

struct ConvTcParams {
    int taps, kw, pad, kblocks, cin;
    int n_img, Ho, Wo;                         // output pixels
    int bw, bh, bn;                            // pixel box of a tile, bw * bh * bn == 128
    int tiles_x, tiles_y, tiles_n, tiles_c;    // tile grid: pixels (x, y, image) and output-channel blocks
    int cout, relu, num_tiles, bias_bf16;
    const void *bias;                          // (cout) float32 or bf16, may be null
    const __nv_bfloat16 *residual;
    __nv_bfloat16 *out;
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor: K-major operand, rows of 128 bytes, 128-byte swizzle, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN> struct TcCfg {
    static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int kBBytes = BN * kTcK * 2;
    static constexpr int kTmemCols = 2 * BN;                                   // two accumulator tiles (power of two >= 32)
    static constexpr size_t kSmem = (size_t)kStages * (kTcABytes + kBBytes) + 1024 /* alignment slack */ + 256 /* barriers */;
};

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvTcParams p) {
    using Cfg = TcCfg<BN>;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                              // swizzle-128B tiles need 1024-byte alignment
    const uint32_t smA = base, smB = base + S * kTcABytes;
    const uint32_t bars = smB + S * Cfg::kBBytes;                              // full[S], empty[S], acc_full[2], acc_empty[2]
    const uint32_t full0 = bars, empty0 = bars + 8 * S, accf0 = bars + 16 * S, acce0 = accf0 + 16, slot = acce0 + 16;
    uint8_t *gen = tc_smem_raw + (base - raw);
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + (slot - base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(accf0 + 8 * i, 1); mbar_init(acce0 + 8 * i, 32 * kTcEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"((uint32_t)Cfg::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int iters = p.taps * p.kblocks;                                      // K blocks per tile
    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int ct = tile % p.tiles_c, mt = tile / p.tiles_c;
                const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tn = mt / (p.tiles_x * p.tiles_y);
                const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tn * p.bn, c0 = ct * BN;
                for (int tap = 0; tap < p.taps; ++tap) {
                    const int r = tap / p.kw, s = tap - r * p.kw;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(empty0 + 8 * stage, phase ^ 1u);
                        mbar_expect_tx(full0 + 8 * stage, (uint32_t)(kTcABytes + Cfg::kBBytes));
                        tma_load_4d(smA + stage * kTcABytes, &tmA, full0 + 8 * stage, kb * kTcK, x0 + s - p.pad, y0 + r - p.pad, n0);
                        tma_load_2d(smB + stage * Cfg::kBBytes, &tmB, full0 + 8 * stage, tap * p.cin + kb * kTcK, c0);
                        if (++stage == S) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            int acc = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(acce0 + 8 * acc, acc_phase ^ 1u);                   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * BN);
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(full0 + 8 * stage, phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smA + stage * kTcABytes), bdesc = umma_desc_sw128(smB + stage * Cfg::kBBytes);
#pragma unroll
                    for (int k = 0; k < kTcK / 16; ++k)                        // K = 16 per instruction: 32 bytes further along the row
                        umma_f16(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
                    umma_commit(empty0 + 8 * stage);                           // frees the stage when these MMAs have read it
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
                umma_commit(accf0 + 8 * acc);                                  // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ================= epilogue: 8 warps; warp w reads TMEM lane quarter w % 4 (32 pixels) and half of the columns ==========
        const int q = warp & 3;                                                // warps 2..9 -> lane quarters 2,3,0,1,2,3,0,1
        const int half = (warp - 2) >> 2;                                      // which half of the tile's BN columns
        const int m = q * 32 + lane;                                           // pixel of the tile this thread owns
        const int ix = m % p.bw, iy = (m / p.bw) % p.bh, in = m / (p.bw * p.bh);
        constexpr int kCols = BN / 2;                                          // columns per warp: 32, 64 or 128
        uint32_t acc_phase = 0;
        int acc = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int ct = tile % p.tiles_c, mt = tile / p.tiles_c;
            const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tn = mt / (p.tiles_x * p.tiles_y);
            const int x = tx * p.bw + ix, y = ty * p.bh + iy, n = tn * p.bn + in, c0 = ct * BN + half * kCols;
            const bool valid = x < p.Wo && y < p.Ho && n < p.n_img;
            const size_t pix = valid ? ((size_t)n * p.Ho + y) * p.Wo + x : 0;
            __nv_bfloat16 *dst = p.out + pix * p.cout + c0;
            const __nv_bfloat16 *res = p.residual ? p.residual + pix * p.cout + c0 : nullptr;
            // the residual of the first 32 columns is requested before the accumulator is even complete, the next group's
            // while the current one is converted: the loads' latency stays off the critical path
            uint4 rnext[4];
            if (res && valid) {
#pragma unroll
                for (int g = 0; g < 4; ++g) rnext[g] = ldg_stream_u4(res + 8 * g);
            }
            mbar_wait(accf0 + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * kCols);
#pragma unroll 1
            for (int c = 0; c < kCols; c += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c, v);
                uint4 rcur[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) rcur[g] = rnext[g];
                if (res && valid && c + 32 < kCols) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) rnext[g] = ldg_stream_u4(res + c + 32 + 8 * g);
                }
                if (valid) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {                              // 8 channels = one 16-byte store
                        float f[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(v[8 * g + k]);
                        if (p.bias) {
                            if (p.bias_bf16) {
                                const uint4 bv = *reinterpret_cast<const uint4 *>(static_cast<const __nv_bfloat16 *>(p.bias) + c0 + c + 8 * g);
                                const uint32_t bw4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) { f[2 * k] += __uint_as_float(bw4[k] << 16); f[2 * k + 1] += __uint_as_float(bw4[k] & 0xffff0000u); }
                            } else {
                                const float *bf = static_cast<const float *>(p.bias) + c0 + c + 8 * g;
                                const float4 b0 = *reinterpret_cast<const float4 *>(bf), b1 = *reinterpret_cast<const float4 *>(bf + 4);
                                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                            }
                        }
                        if (res) {
                            const uint32_t rw[4] = {rcur[g].x, rcur[g].y, rcur[g].z, rcur[g].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) { f[2 * k] += __uint_as_float(rw[k] << 16); f[2 * k + 1] += __uint_as_float(rw[k] & 0xffff0000u); }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
                        }
                        uint32_t o[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                            o[k] = *reinterpret_cast<const uint32_t *>(&h2);
                        }
                        stg_stream_u4(dst + c + 8 * g, make_uint4(o[0], o[1], o[2], o[3]));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acce0 + 8 * acc);                                      // all epilogue threads hand the accumulator back
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// pixel box (bw, bh, bn) with bw * bh * bn == 128 covering a (W, H, N) map with as little padding as possible
void choose_box(int W, int H, int N, int &bw, int &bh, int &bn) {
    long long best = -1;
    bw = 128; bh = 1; bn = 1;
    if (const char *e = getenv("MSQ_TC_BOX")) {                   // development aid: "bw,bh,bn"
        int a = 0, b = 0, c = 0;
        if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a * b * c == 128) { bw = a; bh = b; bn = c; return; }
    }
    for (int w = 1; w <= 128; w <<= 1)
        for (int h = 1; w * h <= 128; h <<= 1) {
            const int n = 128 / (w * h);
            if (w > 256 || h > 256 || n > 256) continue;
            const long long tiles = (long long)((W + w - 1) / w) * ((H + h - 1) / h) * ((N + n - 1) / n);
            // fewest tiles first; among equals prefer wide rows (longer contiguous runs for the TMA unit)
            const long long score = tiles * 1024 - w;
            if (best < 0 || score < best) { best = score; bw = w; bh = h; bn = n; }
        }
}

}  // namespace
}  // namespace msq

using namespace msq;

// x (n, H, W, cin) channels-last bf16; w (cout, ksize, ksize, cin) bf16 (= a channels-last (cout, cin, k, k) tensor's memory);
// out (n, Ho, Wo, cout) bf16 with Ho = (H - 1) / stride + 1.  ksize 1 (stride 1 or 2, pad 0) or 3 (stride 1, pad 1).
extern "C" int msq_conv_tc(const void *x, int n, int H, int W, int cin, const void *w, int cout, int ksize, int stride, const void *bias,
                           int bias_is_bf16, const void *residual, int relu, void *out, void *stream) {
    MSQ_REQUIRE(n >= 0 && H > 0 && W > 0 && cin > 0 && cout > 0, MSQ_EINVAL, "msq_conv_tc: bad sizes");
    MSQ_REQUIRE((ksize == 1 && (stride == 1 || stride == 2)) || (ksize == 3 && stride == 1), MSQ_EUNSUPPORTED,
                "msq_conv_tc: 1x1 (stride 1, 2) and 3x3 (stride 1, padding 1) convolutions only (got k=%d stride=%d)", ksize, stride);
    MSQ_REQUIRE(cin % kTcK == 0 && cout % 64 == 0, MSQ_EUNSUPPORTED, "msq_conv_tc: channel counts must be multiples of 64 (cin=%d cout=%d)", cin, cout);
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(x && w && out, MSQ_EINVAL, "msq_conv_tc: null pointer");
    MSQ_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)w % 16 == 0 && (uintptr_t)out % 16 == 0 && (uintptr_t)residual % 16 == 0 &&
                (uintptr_t)bias % 16 == 0, MSQ_EINVAL, "msq_conv_tc: pointers must be 16-byte aligned");
    EncodeTiledFn encode = encode_tiled_fn();
    MSQ_REQUIRE(encode, MSQ_ECUDA, "msq_conv_tc: cuTensorMapEncodeTiled is not available from the driver");
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    ConvTcParams p;
    p.taps = ksize * ksize; p.kw = ksize; p.pad = ksize / 2; p.kblocks = cin / kTcK; p.cin = cin;
    p.n_img = n; p.Ho = Ho; p.Wo = Wo;
    choose_box(Wo, Ho, n, p.bw, p.bh, p.bn);
    p.tiles_x = (Wo + p.bw - 1) / p.bw; p.tiles_y = (Ho + p.bh - 1) / p.bh; p.tiles_n = (n + p.bn - 1) / p.bn;
    const int BN = cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64);
    p.tiles_c = cout / BN; p.cout = cout; p.relu = relu; p.bias = bias; p.bias_bf16 = bias_is_bf16;
    p.residual = static_cast<const __nv_bfloat16 *>(residual); p.out = static_cast<__nv_bfloat16 *>(out);
    const long long tiles = (long long)p.tiles_x * p.tiles_y * p.tiles_n * p.tiles_c;
    MSQ_REQUIRE(tiles < (1ll << 31), MSQ_EUNSUPPORTED, "msq_conv_tc: too many tiles");
    p.num_tiles = (int)tiles;
    // A: the (strided) input as (C, Wo, Ho, N); out-of-range coordinates read as zero = the convolution's padding
    CUtensorMap tmA, tmB;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)n};
        const cuuint64_t strides[3] = {(cuuint64_t)stride * cin * 2, (cuuint64_t)stride * W * cin * 2, (cuuint64_t)H * W * cin * 2};
        const cuuint32_t box[4] = {(cuuint32_t)kTcK, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bn};
        const cuuint32_t ones[4] = {1, 1, 1, 1};
        const CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(x), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MSQ_REQUIRE(r == CUDA_SUCCESS, MSQ_ECUDA, "msq_conv_tc: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        const cuuint64_t dims[2] = {(cuuint64_t)p.taps * cin, (cuuint64_t)cout};
        const cuuint64_t strides[1] = {(cuuint64_t)p.taps * cin * 2};
        const cuuint32_t box[2] = {(cuuint32_t)kTcK, (cuuint32_t)BN};
        const cuuint32_t ones[2] = {1, 1};
        const CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        MSQ_REQUIRE(r == CUDA_SUCCESS, MSQ_ECUDA, "msq_conv_tc: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = std::min(p.num_tiles, sm_count());
    TimedLaunch timed(K_DETECTOR_GLUE, st);
#define MSQ_TC_LAUNCH(BN_)                                                                                                          \
    do {                                                                                                                            \
        MSQ_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<BN_>::kSmem)); \
        conv_tc_kernel<BN_><<<grid, kTcThreads, TcCfg<BN_>::kSmem, st>>>(tmA, tmB, p);                                              \
    } while (0)
    if (BN == 256) MSQ_TC_LAUNCH(256);
    else if (BN == 128) MSQ_TC_LAUNCH(128);
    else MSQ_TC_LAUNCH(64);
#undef MSQ_TC_LAUNCH
    MSQ_LAUNCH_OK("conv_tc");
    return MSQ_OK;
}
