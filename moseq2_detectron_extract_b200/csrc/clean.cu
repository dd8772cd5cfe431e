// a6 clean_frames(iters_tail=3): 3x3 median (replicate border) + ONE opening by the 9x9 ellipse.
// ref proc/proc.py:480-515 (medianBlur :506, morphologyEx :509 whose `iters_tail` lands in `dst`).
//
// One CTA cleans one horizontal strip (tile) of one frame entirely in shared memory:
//   load u8 tile (+9 px halo, replicate-clamped) -> u16 plane A
//   median3x3(A)            -> plane B (M)      pixels outside the image := 0xFFFF (erode identity)
//   row-min7 / row-min9 (M) -> planes C, D      the ellipse is rows of width 1,7,7,9,9,9,7,7,1
//   column combine          -> plane A (E)      pixels outside the image := 0 (dilate identity)
//   row-max7 / row-max9 (E) -> planes C, D
//   column combine          -> u8 global store
// Pixels are kept as 16-bit lanes so that every min/max is a single VIMNMX(3).U16x2 on two pixels
// (byte-wide SIMD min/max is emulated with 6 ALU ops on sm_100a, 16-bit is native).
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace msq {
namespace {

constexpr int kHX = 12;          // left/right halo columns kept in the planes (multiple of 4 >= 9)
constexpr int kHY = 9;           // top/bottom halo rows (1 median + 4 erode + 4 dilate)
constexpr int kCleanThreads = 512;

struct MinOp {
    static __device__ __forceinline__ uint32_t op3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
};
struct MaxOp {
    static __device__ __forceinline__ uint32_t op3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
};

// (p[k],p[k+1]),(p[k+2],p[k+3]) -> (p[k+1],p[k+2])
__device__ __forceinline__ uint32_t mid_pair(uint32_t a, uint32_t b) { return __funnelshift_r(a, b, 16); }

__device__ __forceinline__ uint32_t med3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
    return __vmaxu2(__vminu2(a, b), __vminu2(__vmaxu2(a, b), c));
}

// lanes of a pixel pair that lie inside the image (0xFFFF per valid lane)
__device__ __forceinline__ uint32_t inside_lanes(int y, int x, int h, int w) {
    if ((unsigned)y >= (unsigned)h) return 0u;
    return ((unsigned)x < (unsigned)w ? 0x0000ffffu : 0u) | ((unsigned)(x + 1) < (unsigned)w ? 0xffff0000u : 0u);
}

struct Tile {
    int tx0, ty0;     // image coords of the first output pixel of the tile
    int TW, TH;       // output tile size (TW multiple of 4)
    int PW, PH;       // plane size: TW + 2*kHX, TH + 2*kHY
};

// rows [r_lo, r_hi), column groups of 4 px starting at c_lo (multiple of 4) up to c_hi (exclusive)
template <class OP>
__device__ __forceinline__ void row_pass(const uint16_t *__restrict__ src, uint16_t *__restrict__ d7,
                                         uint16_t *__restrict__ d9, const Tile &t, int r_lo, int r_hi, int c_lo, int c_hi) {
    const int gpr = (c_hi - c_lo) >> 2;
    const int total = (r_hi - r_lo) * gpr;
    for (int g = threadIdx.x; g < total; g += kCleanThreads) {
        const int r = r_lo + g / gpr;
        const int c = c_lo + ((g % gpr) << 2);
        const uint16_t *p = src + r * t.PW + c;
        const uint2 a = *reinterpret_cast<const uint2 *>(p - 4);
        const uint2 b = *reinterpret_cast<const uint2 *>(p);
        const uint2 cc = *reinterpret_cast<const uint2 *>(p + 4);
        const uint32_t sa = mid_pair(a.x, a.y), sab = mid_pair(a.y, b.x), sb = mid_pair(b.x, b.y),
                       sbc = mid_pair(b.y, cc.x), sc = mid_pair(cc.x, cc.y);
        const uint32_t mid5 = OP::op3(OP::op3(sab, b.x, sb), b.y, sbc);
        const uint32_t h7_0 = OP::op3(mid5, sa, a.y);
        const uint32_t h7_1 = OP::op3(mid5, cc.x, sc);
        const uint32_t h9_0 = OP::op3(h7_0, a.x, cc.x);
        const uint32_t h9_1 = OP::op3(h7_1, a.y, cc.y);
        *reinterpret_cast<uint2 *>(d7 + r * t.PW + c) = make_uint2(h7_0, h7_1);
        *reinterpret_cast<uint2 *>(d9 + r * t.PW + c) = make_uint2(h9_0, h9_1);
    }
}

template <class OP>
__device__ __forceinline__ uint2 column_combine(const uint16_t *__restrict__ m, const uint16_t *__restrict__ s7,
                                                const uint16_t *__restrict__ s9, int PW, int r, int c) {
    const int o = r * PW + c;
    const uint2 m_up = *reinterpret_cast<const uint2 *>(m + o - 4 * PW);
    const uint2 m_dn = *reinterpret_cast<const uint2 *>(m + o + 4 * PW);
    const uint2 a3 = *reinterpret_cast<const uint2 *>(s7 + o - 3 * PW);
    const uint2 a2 = *reinterpret_cast<const uint2 *>(s7 + o - 2 * PW);
    const uint2 b2 = *reinterpret_cast<const uint2 *>(s7 + o + 2 * PW);
    const uint2 b3 = *reinterpret_cast<const uint2 *>(s7 + o + 3 * PW);
    const uint2 n1 = *reinterpret_cast<const uint2 *>(s9 + o - PW);
    const uint2 n0 = *reinterpret_cast<const uint2 *>(s9 + o);
    const uint2 p1 = *reinterpret_cast<const uint2 *>(s9 + o + PW);
    uint2 r2;
    r2.x = OP::op3(OP::op3(m_up.x, m_dn.x, a3.x), OP::op3(a2.x, b2.x, b3.x), OP::op3(n1.x, n0.x, p1.x));
    r2.y = OP::op3(OP::op3(m_up.y, m_dn.y, a3.y), OP::op3(a2.y, b2.y, b3.y), OP::op3(n1.y, n0.y, p1.y));
    return r2;
}

__global__ void __launch_bounds__(kCleanThreads)
clean_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int n, int h, int w, int TW, int TH,
             int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) uint16_t smem[];
    Tile t;
    t.TW = TW; t.TH = TH; t.PW = TW + 2 * kHX; t.PH = TH + 2 * kHY;
    const int plane = t.PW * t.PH;
    uint16_t *A = smem, *B = smem + plane, *C = smem + 2 * plane, *D = smem + 3 * plane;

    const int tiles_per_frame = tiles_x * tiles_y;
    const bool word_ok = (w % 4 == 0) && ((uintptr_t)in % 4 == 0) && ((uintptr_t)out % 4 == 0);

    for (int job = blockIdx.x; job < n * tiles_per_frame; job += gridDim.x) {
        const int f = job / tiles_per_frame;
        const int tt = job - f * tiles_per_frame;
        t.ty0 = (tt / tiles_x) * TH;
        t.tx0 = (tt % tiles_x) * TW;
        const uint8_t *src = in + (size_t)f * h * w;
        uint8_t *dst = out + (size_t)f * h * w;
        const int xbase = t.tx0 - kHX, ybase = t.ty0 - kHY;

        // ---- P0: load tile + halo into plane A (u16), clamped to the image (BORDER_REPLICATE)
        {
            const int gpr = t.PW >> 2;
            for (int g = threadIdx.x; g < t.PH * gpr; g += kCleanThreads) {
                const int r = g / gpr, c = (g % gpr) << 2;
                const int y = min(max(ybase + r, 0), h - 1);
                const int x = xbase + c;
                uint32_t v;
                if (word_ok && x >= 0 && x + 3 < w) {
                    v = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)y * w + x));
                } else {
                    const uint8_t *row = src + (size_t)y * w;
                    v = (uint32_t)row[min(max(x, 0), w - 1)] | ((uint32_t)row[min(max(x + 1, 0), w - 1)] << 8) |
                        ((uint32_t)row[min(max(x + 2, 0), w - 1)] << 16) | ((uint32_t)row[min(max(x + 3, 0), w - 1)] << 24);
                }
                *reinterpret_cast<uint2 *>(A + r * t.PW + c) = make_uint2(__byte_perm(v, 0, 0x4140), __byte_perm(v, 0, 0x4342));
            }
        }
        __syncthreads();

        // ---- P1: 3x3 median A -> B for rows [1,PH-1), cols [4,PW-4)
        {
            const int c_lo = 4, c_hi = t.PW - 4;
            const int gpr = (c_hi - c_lo) >> 2;
            for (int g = threadIdx.x; g < (t.PH - 2) * gpr; g += kCleanThreads) {
                const int r = 1 + g / gpr, c = c_lo + ((g % gpr) << 2);
                uint32_t lo[4], mi[4], hi[4];
                {
                    uint32_t v[3][4];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint16_t *p = A + (r - 1 + k) * t.PW + c;
                        v[k][0] = *reinterpret_cast<const uint32_t *>(p - 2);
                        const uint2 m = *reinterpret_cast<const uint2 *>(p);
                        v[k][1] = m.x; v[k][2] = m.y;
                        v[k][3] = *reinterpret_cast<const uint32_t *>(p + 4);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        lo[q] = __vimin3_u16x2(v[0][q], v[1][q], v[2][q]);
                        hi[q] = __vimax3_u16x2(v[0][q], v[1][q], v[2][q]);
                        mi[q] = med3_u16x2(v[0][q], v[1][q], v[2][q]);
                    }
                }
                uint32_t res[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {          // output pair q sits in slot q+1
                    const uint32_t lo_l = mid_pair(lo[q], lo[q + 1]), lo_r = mid_pair(lo[q + 1], lo[q + 2]);
                    const uint32_t mi_l = mid_pair(mi[q], mi[q + 1]), mi_r = mid_pair(mi[q + 1], mi[q + 2]);
                    const uint32_t hi_l = mid_pair(hi[q], hi[q + 1]), hi_r = mid_pair(hi[q + 1], hi[q + 2]);
                    const uint32_t max_lo = __vimax3_u16x2(lo_l, lo[q + 1], lo_r);
                    const uint32_t min_hi = __vimin3_u16x2(hi_l, hi[q + 1], hi_r);
                    const uint32_t med_mi = med3_u16x2(mi_l, mi[q + 1], mi_r);
                    res[q] = med3_u16x2(max_lo, med_mi, min_hi);
                    // outside the image the erosion must ignore the pixel
                    res[q] |= ~inside_lanes(ybase + r, xbase + c + 2 * q, h, w);
                }
                *reinterpret_cast<uint2 *>(B + r * t.PW + c) = make_uint2(res[0], res[1]);
            }
        }
        __syncthreads();

        // ---- P2: row minima of width 7 / 9 of M -> C, D   rows [1,PH-1), cols [8,PW-8)
        row_pass<MinOp>(B, C, D, t, 1, t.PH - 1, 8, t.PW - 8);
        __syncthreads();

        // ---- P3: erosion E -> A   rows [5,PH-5), cols [8,PW-8); outside the image := 0
        {
            const int c_lo = 8, c_hi = t.PW - 8, r_lo = 5, r_hi = t.PH - 5;
            const int gpr = (c_hi - c_lo) >> 2;
            for (int g = threadIdx.x; g < (r_hi - r_lo) * gpr; g += kCleanThreads) {
                const int r = r_lo + g / gpr, c = c_lo + ((g % gpr) << 2);
                uint2 e = column_combine<MinOp>(B, C, D, t.PW, r, c);
                e.x &= inside_lanes(ybase + r, xbase + c, h, w);
                e.y &= inside_lanes(ybase + r, xbase + c + 2, h, w);
                *reinterpret_cast<uint2 *>(A + r * t.PW + c) = e;
            }
        }
        __syncthreads();

        // ---- P4: row maxima of E -> C, D   rows [5,PH-5), cols [12,PW-12)
        row_pass<MaxOp>(A, C, D, t, 5, t.PH - 5, 12, t.PW - 12);
        __syncthreads();

        // ---- P5: dilation -> global u8   rows [9,PH-9), cols [12,PW-12)
        {
            const int gpr = t.TW >> 2;
            for (int g = threadIdx.x; g < t.TH * gpr; g += kCleanThreads) {
                const int r = kHY + g / gpr, c = kHX + ((g % gpr) << 2);
                const int y = ybase + r, x = xbase + c;
                if (y >= h || x >= w) continue;
                const uint2 d = column_combine<MaxOp>(A, C, D, t.PW, r, c);
                const uint32_t packed = __byte_perm(d.x, d.y, 0x6420);
                if (word_ok && x + 3 < w) {
                    *reinterpret_cast<uint32_t *>(dst + (size_t)y * w + x) = packed;
                } else {
                    for (int k = 0; k < 4 && x + k < w; ++k) dst[(size_t)y * w + x + k] = (uint8_t)(packed >> (8 * k));
                }
            }
        }
        __syncthreads();     // planes are reused by the next job
    }
}

}  // namespace

// tile geometry shared with the fused pipeline
struct CleanPlan { int TW, TH, tiles_x, tiles_y; size_t smem; };

CleanPlan make_clean_plan(int h, int w) {
    CleanPlan p;
    const int w4 = (w + 3) & ~3;
    p.TW = (w4 <= 256) ? w4 : 128;
    int th = 30;                                   // 2 CTAs/SM at 264-px planes (101 KB each)
    if (const char *e = getenv("MSQ_CLEAN_TH")) { int v = atoi(e); if (v >= 4 && v <= 256) th = v; }
    p.TH = std::min(th, h);
    p.tiles_x = (w + p.TW - 1) / p.TW;
    p.tiles_y = (h + p.TH - 1) / p.TH;
    p.smem = (size_t)4 * (p.TW + 2 * kHX) * (p.TH + 2 * kHY) * sizeof(uint16_t);
    return p;
}

int launch_clean(const uint8_t *in, uint8_t *out, int n, int h, int w, cudaStream_t st) {
    const CleanPlan p = make_clean_plan(h, w);
    MSQ_REQUIRE(p.smem <= 227 * 1024, MSQ_EUNSUPPORTED, "clean_frames: tile needs %zu B of shared memory", p.smem);
    static thread_local size_t configured = 0;
    if (p.smem > configured) {
        MSQ_CUDA_OK(cudaFuncSetAttribute(clean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        configured = p.smem;
    }
    const long long jobs = (long long)n * p.tiles_x * p.tiles_y;
    const int per_sm = std::max(1, (int)((227 * 1024) / (p.smem + 1024)));
    const int grid = (int)std::min<long long>(jobs, (long long)sm_count() * per_sm);
    TimedLaunch timed(K_CLEAN, st);
    clean_kernel<<<grid, kCleanThreads, p.smem, st>>>(in, out, n, h, w, p.TW, p.TH, p.tiles_x, p.tiles_y);
    MSQ_LAUNCH_OK("clean_frames");
    return MSQ_OK;
}

}  // namespace msq

extern "C" int msq_clean_frames(const uint8_t *in, uint8_t *out, int n, int h, int w, void *stream) {
    MSQ_REQUIRE(in && out && in != out, MSQ_EINVAL, "msq_clean_frames: null or aliased pointers");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0, MSQ_EINVAL, "msq_clean_frames: bad sizes n=%d h=%d w=%d", n, h, w);
    if (n == 0) return MSQ_OK;
    return msq::launch_clean(in, out, n, h, w, (cudaStream_t)stream);
}
