// a6 clean_frames(iters_tail=3): 3x3 median (replicate border) + ONE opening by the 9x9 ellipse.
// ref proc/proc.py:480-515 (medianBlur :506, morphologyEx :509 whose `iters_tail` lands in `dst`).
//
// One CTA cleans one horizontal strip (tile) of one frame entirely in shared memory:
//   load u8 tile (+halo, replicate-clamped) -> u16 plane A
//   median3x3(A)            -> plane B (M)      pixels outside the image := 255   (erode identity)
//   row-min7 / row-min9 (M) -> planes C, D      the ellipse is rows of width 1,7,7,9,9,9,7,7,1
//   column combine          -> plane A (E)      pixels outside the image := 0     (dilate identity)
//   row-max7 / row-max9 (E) -> planes C, D
//   column combine          -> u8 global store
// Pixels are 16-bit lanes, two per register, so every min/max is one VIMNMX(3).U16x2 (measured on
// B200: 2-input 4 warp-instr/clk/SM, 3-input 2; byte-wide __vminu4 is emulated with 6 ALU ops).
// A work item is 8 horizontally adjacent pixels (one 128-bit shared-memory access).  Lanes map to column
// groups and warps stride over rows, so the inner loops contain no index arithmetic at all
// (ncu on the first version: 85 % of the executed instructions were index arithmetic / division).
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace msq {
namespace {

constexpr int kHX = 16;          // left/right halo columns kept in the planes (multiple of 8 >= 9)
constexpr int kHY = 9;           // top/bottom halo rows (1 median + 4 erode + 4 dilate)

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
__device__ __forceinline__ uint32_t inside_lanes(bool row_ok, int x, int w) {
    if (!row_ok) return 0u;
    return ((unsigned)x < (unsigned)w ? 0x0000ffffu : 0u) | ((unsigned)(x + 1) < (unsigned)w ? 0xffff0000u : 0u);
}

struct Geometry {
    int h, w;
    int TW, TH, PW, PH;            // tile, plane (PW = TW + 2*kHX <= 272, PH = TH + 2*kHY)
    int tiles_x, tiles_y;
    int groups;                    // 8-px column groups of the stencil passes: (PW - 16) / 8  (<= 32)
};


// Thread mapping of every pass: lane <-> 8-pixel column group (one warp spans a full 256-pixel plane row with
// 128-bit shared-memory accesses), warps stride over rows.  No per-item index arithmetic is left.

// 7- and 9-wide row extrema of `src` at plane offset o (this lane's 8-pixel column group of one row)
template <class OP>
__device__ __forceinline__ void row_extrema(const uint16_t *__restrict__ src, uint16_t *__restrict__ d7,
                                            uint16_t *__restrict__ d9, int o) {
    const uint2 a = *reinterpret_cast<const uint2 *>(src + o - 4);      // px -4..-1
    const uint4 b = *reinterpret_cast<const uint4 *>(src + o);          // px  0..7
    const uint2 c = *reinterpret_cast<const uint2 *>(src + o + 8);      // px  8..11
    const uint32_t s_a = mid_pair(a.x, a.y), s_ab = mid_pair(a.y, b.x), s_b0 = mid_pair(b.x, b.y),
                   s_b1 = mid_pair(b.y, b.z), s_b2 = mid_pair(b.z, b.w), s_bc = mid_pair(b.w, c.x),
                   s_c = mid_pair(c.x, c.y);
    const uint32_t t1 = OP::op3(s_ab, b.x, s_b0), t2 = OP::op3(b.y, s_b1, b.z), t3 = OP::op3(s_b2, b.w, s_bc);
    uint4 h7, h9;
    h7.x = OP::op3(OP::op3(t1, s_a, a.y), b.y, s_b1);
    h7.y = OP::op3(t1, t2, s_b2);
    h7.z = OP::op3(s_b0, t2, t3);
    h7.w = OP::op3(OP::op3(s_b1, b.z, t3), c.x, s_c);
    h9.x = OP::op3(h7.x, a.x, b.z);
    h9.y = OP::op3(h7.y, a.y, b.w);
    h9.z = OP::op3(h7.z, b.x, c.x);
    h9.w = OP::op3(h7.w, b.y, c.y);
    *reinterpret_cast<uint4 *>(d7 + o) = h7;
    *reinterpret_cast<uint4 *>(d9 + o) = h9;
}

template <class OP>
__device__ __forceinline__ uint4 column_combine(const uint16_t *__restrict__ m, const uint16_t *__restrict__ s7,
                                                const uint16_t *__restrict__ s9, int PW, int o) {
    const uint4 m_up = *reinterpret_cast<const uint4 *>(m + o - 4 * PW);
    const uint4 m_dn = *reinterpret_cast<const uint4 *>(m + o + 4 * PW);
    const uint4 a3 = *reinterpret_cast<const uint4 *>(s7 + o - 3 * PW);
    const uint4 a2 = *reinterpret_cast<const uint4 *>(s7 + o - 2 * PW);
    const uint4 b2 = *reinterpret_cast<const uint4 *>(s7 + o + 2 * PW);
    const uint4 b3 = *reinterpret_cast<const uint4 *>(s7 + o + 3 * PW);
    const uint4 n1 = *reinterpret_cast<const uint4 *>(s9 + o - PW);
    const uint4 n0 = *reinterpret_cast<const uint4 *>(s9 + o);
    const uint4 p1 = *reinterpret_cast<const uint4 *>(s9 + o + PW);
    uint4 r;
    r.x = OP::op3(OP::op3(m_up.x, m_dn.x, a3.x), OP::op3(a2.x, b2.x, b3.x), OP::op3(n1.x, n0.x, p1.x));
    r.y = OP::op3(OP::op3(m_up.y, m_dn.y, a3.y), OP::op3(a2.y, b2.y, b3.y), OP::op3(n1.y, n0.y, p1.y));
    r.z = OP::op3(OP::op3(m_up.z, m_dn.z, a3.z), OP::op3(a2.z, b2.z, b3.z), OP::op3(n1.z, n0.z, p1.z));
    r.w = OP::op3(OP::op3(m_up.w, m_dn.w, a3.w), OP::op3(a2.w, b2.w, b3.w), OP::op3(n1.w, n0.w, p1.w));
    return r;
}

// 8 replicate-clamped input pixels starting at image column x of row `row`, widened to 16-bit lanes
__device__ __forceinline__ uint4 load_group(const uint8_t *__restrict__ row, int x, int w, bool vec_ok) {
    uint2 v;
    if (vec_ok) {
        // groups are 8-aligned: entirely inside, left of, or right of the image -> load the nearest in-image
        // group and broadcast its edge byte when outside (no divergence)
        v = __ldg(reinterpret_cast<const uint2 *>(row + min(max(x, 0), w - 8)));
        if (x < 0) { v.x = (v.x & 0xffu) * 0x01010101u; v.y = v.x; }
        else if (x >= w) { v.y = (v.y >> 24) * 0x01010101u; v.x = v.y; }
    } else {
        uint32_t b[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) b[k] = row[min(max(x + k, 0), w - 1)];
        v.x = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        v.y = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
    }
    return make_uint4(__byte_perm(v.x, 0, 0x4140), __byte_perm(v.x, 0, 0x4342), __byte_perm(v.y, 0, 0x4140),
                      __byte_perm(v.y, 0, 0x4342));
}

// PWT: plane pitch as a compile-time constant (0 = runtime).  With a constant pitch every shared-memory
// access of the stencil passes is base + immediate (ncu: with a runtime pitch the address arithmetic was as
// many instructions as the min/max work).  VEC: the 8-aligned fast path (w % 8 == 0, aligned pointers).
template <int PWT, bool VEC, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
clean_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int n, Geometry G) {
    extern __shared__ __align__(16) uint16_t smem[];
    const int PW = PWT ? PWT : G.PW;
    const int plane = PW * G.PH;
    uint16_t *A = smem, *B = smem + plane, *C = smem + 2 * plane, *D = smem + 3 * plane;
    const int h = G.h, w = G.w, PH = G.PH;
    const int tiles_per_frame = G.tiles_x * G.tiles_y;
    constexpr bool vec_ok = VEC;
    constexpr int kWarps = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool active = (PWT == 272) ? true : lane < G.groups;   // this lane owns a stencil column group
    const int col = 8 + (lane << 3);                // its first plane column

    for (int job = blockIdx.x; job < n * tiles_per_frame; job += gridDim.x) {
        const int f = job / tiles_per_frame;
        const int tt = job - f * tiles_per_frame;
        const int ty0 = (tt / G.tiles_x) * G.TH, tx0 = (tt % G.tiles_x) * G.TW;
        const uint8_t *src = in + (size_t)f * h * w;
        uint8_t *dst = out + (size_t)f * h * w;
        const int xbase = tx0 - kHX, ybase = ty0 - kHY;
        const int x_img = xbase + col;                                   // image column of this lane's group
        // aligned case: the whole group is inside or outside the image; general case: per-lane masks
        const bool grp_in = (unsigned)x_img < (unsigned)w;
        uint32_t in_lanes[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) in_lanes[q] = inside_lanes(true, x_img + 2 * q, w);

        // ---- P0: load tile + halo into plane A (u16), clamped to the image (BORDER_REPLICATE)
        for (int r = warp; r < PH; r += kWarps) {
            const uint8_t *row = src + (size_t)min(max(ybase + r, 0), h - 1) * w;
            if ((lane << 3) < PW)
                *reinterpret_cast<uint4 *>(A + r * PW + (lane << 3)) = load_group(row, xbase + (lane << 3), w, vec_ok);
            if ((lane << 3) + 256 < PW)                                  // plane columns 256.. (two more groups at PW = 272)
                *reinterpret_cast<uint4 *>(A + r * PW + 256 + (lane << 3)) = load_group(row, xbase + 256 + (lane << 3), w, vec_ok);
        }
        __syncthreads();

        // ---- P1: 3x3 median A -> B for rows [1,PH-1)
        if (active) {
            for (int r = 1 + warp; r < PH - 1; r += kWarps) {
                uint32_t lo[6], mi[6], hi[6];
                {
                    uint32_t v[3][6];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint16_t *p = A + (r - 1 + k) * PW + col;
                        v[k][0] = *reinterpret_cast<const uint32_t *>(p - 2);
                        const uint4 m = *reinterpret_cast<const uint4 *>(p);
                        v[k][1] = m.x; v[k][2] = m.y; v[k][3] = m.z; v[k][4] = m.w;
                        v[k][5] = *reinterpret_cast<const uint32_t *>(p + 8);
                    }
#pragma unroll
                    for (int q = 0; q < 6; ++q) {          // sort every pixel column of the 3 rows
                        lo[q] = __vimin3_u16x2(v[0][q], v[1][q], v[2][q]);
                        hi[q] = __vimax3_u16x2(v[0][q], v[1][q], v[2][q]);
                        mi[q] = med3_u16x2(v[0][q], v[1][q], v[2][q]);
                    }
                }
                uint32_t res[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {              // output pair q sits in slot q+1
                    const uint32_t max_lo = __vimax3_u16x2(mid_pair(lo[q], lo[q + 1]), lo[q + 1], mid_pair(lo[q + 1], lo[q + 2]));
                    const uint32_t min_hi = __vimin3_u16x2(mid_pair(hi[q], hi[q + 1]), hi[q + 1], mid_pair(hi[q + 1], hi[q + 2]));
                    const uint32_t med_mi = med3_u16x2(mid_pair(mi[q], mi[q + 1]), mi[q + 1], mid_pair(mi[q + 1], mi[q + 2]));
                    res[q] = med3_u16x2(max_lo, med_mi, min_hi);
                }
                // the erosion must ignore pixels outside the image: they become 255
                const bool row_ok = (unsigned)(ybase + r) < (unsigned)h;
                if (vec_ok) {
                    if (!(row_ok && grp_in)) res[0] = res[1] = res[2] = res[3] = 0x00ff00ffu;
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t keep = row_ok ? in_lanes[q] : 0u;
                        res[q] = (res[q] & keep) | (0x00ff00ffu & ~keep);
                    }
                }
                *reinterpret_cast<uint4 *>(B + r * PW + col) = make_uint4(res[0], res[1], res[2], res[3]);
            }
        }
        // ---- P2: row minima of width 7 / 9 of M -> C, D, same rows.  A warp owns whole plane rows, so the row
        //      pass only needs the median row this warp just wrote: __syncwarp() instead of a block barrier.
        //      (columns 4..7 and PW-8..PW-5 of B hold stale data; it only reaches outputs that are never used)
        __syncwarp();
        if (active)
            for (int r = 1 + warp; r < PH - 1; r += kWarps) row_extrema<MinOp>(B, C, D, r * PW + col);
        __syncthreads();

        // ---- P3: erosion E -> A (the input plane is dead)   rows [5,PH-5); outside the image := 0
        uint16_t *E = A;
        if (active) {
            for (int r = 5 + warp; r < PH - 5; r += kWarps) {
                uint4 e = column_combine<MinOp>(B, C, D, PW, r * PW + col);
                const bool row_ok = (unsigned)(ybase + r) < (unsigned)h;
                if (vec_ok) {
                    if (!(row_ok && grp_in)) e = make_uint4(0u, 0u, 0u, 0u);
                } else {
                    e.x &= row_ok ? in_lanes[0] : 0u; e.y &= row_ok ? in_lanes[1] : 0u;
                    e.z &= row_ok ? in_lanes[2] : 0u; e.w &= row_ok ? in_lanes[3] : 0u;
                }
                *reinterpret_cast<uint4 *>(E + r * PW + col) = e;
            }
        }
        __syncthreads();     // every warp is done reading C, D (the minima) before they are overwritten

        // ---- P4: row maxima of E -> C, D   rows [5,PH-5): again rows owned by this warp
        if (active)
            for (int r = 5 + warp; r < PH - 5; r += kWarps) row_extrema<MaxOp>(E, C, D, r * PW + col);
        __syncthreads();

        // ---- P5: dilation -> global u8   rows [9,PH-9), plane columns [kHX, PW-kHX)
        if (active && col >= kHX && col < PW - kHX && x_img < w) {
            for (int r = kHY + warp; r < PH - kHY; r += kWarps) {
                const int y = ybase + r;
                if (y >= h) break;
                const uint4 d = column_combine<MaxOp>(E, C, D, PW, r * PW + col);
                const uint2 packed = make_uint2(__byte_perm(d.x, d.y, 0x6420), __byte_perm(d.z, d.w, 0x6420));
                if (vec_ok) {
                    *reinterpret_cast<uint2 *>(dst + (y * w + x_img)) = packed;
                } else {
                    const unsigned long long bits = (unsigned long long)packed.x | ((unsigned long long)packed.y << 32);
                    for (int k = 0; k < 8 && x_img + k < w; ++k) dst[(size_t)y * w + x_img + k] = (uint8_t)(bits >> (8 * k));
                }
            }
        }
        __syncthreads();     // planes are reused by the next job
    }
}

Geometry make_geometry(int h, int w) {
    Geometry G;
    G.h = h; G.w = w;
    const int w8 = (w + 7) & ~7;
    G.TW = std::min(w8, 240);                      // 240 + 2*16 halo = 272 plane columns = 34 groups; 32 stencil groups = one warp
    int th = 80;                                   // 98-row planes: 213 KB, one 1024-thread CTA per SM (fastest of 24..80 on B200)
    if (const char *e = getenv("MSQ_CLEAN_TH")) { int v = atoi(e); if (v >= 4 && v <= 256) th = v; }
    G.TH = std::min(th, h);
    G.PW = G.TW + 2 * kHX;
    G.PH = G.TH + 2 * kHY;
    G.tiles_x = (w + G.TW - 1) / G.TW;
    G.tiles_y = (h + G.TH - 1) / G.TH;
    G.groups = (G.PW - 16) / 8;
    return G;
}

}  // namespace

// per (frame, 240-column tile): the band (int2) + one entry of the cost prefix (int, items + 1 of them)
size_t clean_scratch_bytes(int n, int w) {
    const size_t items = (size_t)(n > 0 ? n : 0) * ((w + 239) / 240);
    return items * sizeof(int2) + (items + 1) * sizeof(int);
}

int launch_clean(const uint8_t *in, uint8_t *out, int n, int h, int w, cudaStream_t st, int2 *bands, RowBands *written,
                 const uint32_t *positive_bits) {
    if (written) *written = {nullptr, 0};
    // fast path: the streaming warp-per-strip kernel (clean_stream.cu); the tiled kernel below serves widths that are
    // not a multiple of 8 / unaligned pointers, and MSQ_CLEAN_TILED=1 forces it for A/B comparisons
    static const bool force_tiled = getenv("MSQ_CLEAN_TILED") != nullptr;
    if (!force_tiled) {
        const int rc = launch_clean_stream(in, out, n, h, w, st, bands, written, positive_bits);
        if (rc != -100) return rc;
    }
    const Geometry G = make_geometry(h, w);
    const size_t smem = (size_t)4 * G.PW * G.PH * sizeof(uint16_t);
    MSQ_REQUIRE(smem <= 227 * 1024, MSQ_EUNSUPPORTED, "clean_frames: tile needs %zu B of shared memory", smem);
    const long long jobs = (long long)n * G.tiles_x * G.tiles_y;
    const int per_sm = std::max(1, std::min(2, (int)((227 * 1024) / (smem + 1024))));
    const int grid = (int)std::min<long long>(jobs, (long long)sm_count() * per_sm);
    const bool vec = (w % 8 == 0) && ((uintptr_t)in % 8 == 0) && ((uintptr_t)out % 8 == 0);
    // strips that leave room for two CTAs per SM run 512 threads each, taller strips one CTA of 1024
    const bool big = per_sm < 2;
    void (*kernel)(const uint8_t *, uint8_t *, int, Geometry);
    if (big) kernel = (G.PW == 272 && vec) ? clean_kernel<272, true, 1024> : (vec ? clean_kernel<0, true, 1024> : clean_kernel<0, false, 1024>);
    else kernel = (G.PW == 272 && vec) ? clean_kernel<272, true, 512> : (vec ? clean_kernel<0, true, 512> : clean_kernel<0, false, 512>);
    MSQ_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TimedLaunch timed(K_CLEAN, st);
    kernel<<<grid, big ? 1024 : 512, smem, st>>>(in, out, n, G);
    MSQ_LAUNCH_OK("clean_frames");
    return MSQ_OK;
}

}  // namespace msq

extern "C" int msq_clean_frames(const uint8_t *in, uint8_t *out, int n, int h, int w, void *stream) {
    MSQ_REQUIRE(n == 0 || (in && out && in != out), MSQ_EINVAL, "msq_clean_frames: null or aliased pointers");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0, MSQ_EINVAL, "msq_clean_frames: bad sizes n=%d h=%d w=%d", n, h, w);
    if (n == 0) return MSQ_OK;
    return msq::launch_clean(in, out, n, h, w, (cudaStream_t)stream);
}

extern "C" size_t msq_clean_scratch_bytes(int n, int h, int w) { (void)h; return msq::clean_scratch_bytes(n, w); }

extern "C" int msq_clean_frames_ws(const uint8_t *in, const uint32_t *positive_bits, uint8_t *out, int n, int h, int w, void *scratch,
                                   size_t scratch_bytes, void *stream) {
    MSQ_REQUIRE((uintptr_t)positive_bits % 4 == 0, MSQ_EINVAL, "msq_clean_frames_ws: positive_bits must be 4-byte aligned");
    MSQ_REQUIRE(n == 0 || (in && out && in != out), MSQ_EINVAL, "msq_clean_frames_ws: null or aliased pointers");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0, MSQ_EINVAL, "msq_clean_frames_ws: bad sizes n=%d h=%d w=%d", n, h, w);
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(scratch && (uintptr_t)scratch % 8 == 0 && scratch_bytes >= msq::clean_scratch_bytes(n, w), MSQ_ENOMEM,
                "msq_clean_frames_ws: scratch must be 8-byte aligned and >= %zu bytes", msq::clean_scratch_bytes(n, w));
    return msq::launch_clean(in, out, n, h, w, (cudaStream_t)stream, static_cast<int2 *>(scratch), nullptr, positive_bits);
}
