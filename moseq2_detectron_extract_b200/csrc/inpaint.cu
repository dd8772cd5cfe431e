// Invalid-pixel in-painting: cv2.inpaint(frame, mask, 3, cv2.INPAINT_NS) of ref proc/proc.py:189-210,
// bit-exact (OpenCV photo/src/inpaint.cpp: icvInpaint + icvNSInpaintFMM, 8-bit single channel).
//
// The algorithm is a fast-marching sweep: pixels leave a priority queue in order of their distance T to the
// known region; each newly reached unknown pixel gets T from its 4-neighbours (FastMarching_solve) and a value
// from a weighted sum over the known pixels within `radius`, and is pushed with its T.  It is inherently
// sequential per frame, so parallelism is ACROSS frames (one warp per flagged frame) and, within a pixel,
// across the (2*radius+1)^2 window taps:
//   * INSIDE flags are a bit image in shared memory (KNOWN vs BAND is never tested by the NS variant)
//   * OpenCV's queue is a sorted linked list with FIFO order among equal T == a min-heap on (T, sequence no.)
//   * T lives in a per-frame float image that is initialised sparsely: only INSIDE pixels (1e6) and band
//     pixels (0) are ever read, plus the 1-pixel ring around the image which is handled by a bounds check
//   * the window taps are computed by the lanes in parallel, but Ia and s are float32 sums in OpenCV's
//     raster order, so one lane adds the 49 terms in that order (zeros for skipped taps do not change them)
// float32/float64 operations use the round-to-nearest intrinsics in OpenCV's operation order (no contraction).
#include "common.cuh"
#include <algorithm>
#include <math.h>

namespace msq {
namespace {

constexpr int kInpaintWarps = 2;           // frames per CTA
constexpr int kHeapSmem = 1024;            // heap entries kept in shared memory per frame (larger heaps spill to scratch)
constexpr int kMaxTaps = 81;               // radius <= 4

struct FrameCtx {
    uint32_t *F;            // INSIDE bits, (h, WB) words in shared memory
    int h, w, WB;
    uint8_t *out;           // the frame, updated in place (global)
    float *t;               // (h, w) distances (global scratch, sparsely initialised)
};

__device__ __forceinline__ bool is_inside(const FrameCtx &c, int ke, int le) {      // extended coordinates
    const int y = ke - 1, x = le - 1;
    if ((unsigned)y >= (unsigned)c.h || (unsigned)x >= (unsigned)c.w) return false;
    return (c.F[y * c.WB + (x >> 5)] >> (x & 31)) & 1u;
}
__device__ __forceinline__ float read_t(const FrameCtx &c, int ke, int le) {
    const int y = ke - 1, x = le - 1;
    if ((unsigned)y >= (unsigned)c.h || (unsigned)x >= (unsigned)c.w) return 1.0e6f;   // the ring keeps its initial value
    return __ldcg(c.t + y * c.w + x);
}
__device__ __forceinline__ int read_out(const FrameCtx &c, int y, int x) { return (int)__ldcg(c.out + y * c.w + x); }

// FastMarching_solve of inpaint.cpp
__device__ float fmm_solve(const FrameCtx &c, int i1, int j1, int i2, int j2) {
    const double a11 = (double)read_t(c, i1, j1), a22 = (double)read_t(c, i2, j2);
    const double m12 = a11 < a22 ? a11 : a22;
    double sol;
    if (!is_inside(c, i1, j1)) {
        if (!is_inside(c, i2, j2)) {
            if (fabs(a11 - a22) >= 1.0) sol = 1 + m12;
            else sol = __dmul_rn(__dadd_rn(__dadd_rn(a11, a22), __dsqrt_rn(__dsub_rn(2.0, __dmul_rn(a11 - a22, a11 - a22)))), 0.5);
        } else {
            sol = 1 + a11;
        }
    } else if (!is_inside(c, i2, j2)) {
        sol = 1 + a22;
    } else {
        sol = 1 + m12;
    }
    return (float)sol;
}

// one window tap of icvNSInpaintFMM for the pixel (i, j) (extended coords): returns weight w and w * value
__device__ void ns_tap(const FrameCtx &c, int i, int j, int k, int l, int radius, float &w_out, float &wv_out) {
    w_out = 0.0f; wv_out = 0.0f;
    const int er = c.h + 2, ec = c.w + 2;
    if (!(k > 0 && l > 0 && k < er - 1 && l < ec - 1)) return;
    if (is_inside(c, k, l)) return;
    if ((l - j) * (l - j) + (k - i) * (k - i) > radius * radius) return;
    const int km = k - 1 + (k == 1), kp = k - 1 - (k == er - 2);
    const int lm = l - 1 + (l == 1), lp = l - 1 - (l == ec - 2);
    const float ry = (float)(i - k), rx = (float)(j - l);
    const float len_r = __fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry));
    const float dst = __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(len_r, len_r), 1.0f));
    float gx, gy;
    if (!is_inside(c, k + 1, l)) {
        if (!is_inside(c, k - 1, l))
            gx = (float)(abs(read_out(c, kp + 1, lm) - read_out(c, kp, lm)) + abs(read_out(c, kp, lm) - read_out(c, km - 1, lm)));
        else
            gx = __fmul_rn((float)abs(read_out(c, kp + 1, lm) - read_out(c, kp, lm)), 2.0f);
    } else {
        if (!is_inside(c, k - 1, l)) gx = __fmul_rn((float)abs(read_out(c, kp, lm) - read_out(c, km - 1, lm)), 2.0f);
        else gx = 0.0f;
    }
    if (!is_inside(c, k, l + 1)) {
        if (!is_inside(c, k, l - 1))
            gy = (float)(abs(read_out(c, km, lp + 1) - read_out(c, km, lm)) + abs(read_out(c, km, lm) - read_out(c, km, lm - 1)));
        else
            gy = __fmul_rn((float)abs(read_out(c, km, lp + 1) - read_out(c, km, lm)), 2.0f);
    } else {
        if (!is_inside(c, k, l - 1)) gy = __fmul_rn((float)abs(read_out(c, km, lm) - read_out(c, km, lm - 1)), 2.0f);
        else gy = 0.0f;
    }
    gx = -gx;
    const float dot = __fadd_rn(__fmul_rn(rx, gx), __fmul_rn(ry, gy));
    float dir;
    if (fabsf(dot) <= 0.01f) {           // the C code compares fabs(float) <= 0.01 (double): same outcome for these magnitudes
        dir = 0.000001f;
    } else {
        const float len_g = __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy));
        dir = fabsf(__fdiv_rn(dot, __fsqrt_rn(__fmul_rn(len_r, len_g))));
    }
    const float w = __fmul_rn(dst, dir);
    w_out = w;
    wv_out = __fmul_rn(w, (float)read_out(c, k - 1, l - 1));
}

// ---- binary min-heap on 64-bit keys (float bits of T << 32 | sequence number), payload = packed (i, j) -------
struct Heap {
    unsigned long long *key;
    uint32_t *pos;
    int size;
};
__device__ __forceinline__ void heap_push(Heap &hp, unsigned long long k, uint32_t p) {
    int i = hp.size++;
    while (i > 0) {
        const int parent = (i - 1) >> 1;
        const unsigned long long pk = hp.key[parent];
        if (pk <= k) break;
        hp.key[i] = pk; hp.pos[i] = hp.pos[parent];
        i = parent;
    }
    hp.key[i] = k; hp.pos[i] = p;
}
__device__ __forceinline__ uint32_t heap_pop(Heap &hp) {
    const uint32_t top = hp.pos[0];
    const int n = --hp.size;
    if (n > 0) {
        const unsigned long long k = hp.key[n];
        const uint32_t p = hp.pos[n];
        int i = 0;
        for (;;) {
            int child = 2 * i + 1;
            if (child >= n) break;
            if (child + 1 < n && hp.key[child + 1] < hp.key[child]) ++child;
            if (hp.key[child] >= k) break;
            hp.key[i] = hp.key[child]; hp.pos[i] = hp.pos[child];
            i = child;
        }
        hp.key[i] = k; hp.pos[i] = p;
    }
    return top;
}

__global__ void __launch_bounds__(kInpaintWarps * 32)
inpaint_kernel(uint8_t *__restrict__ frames, const uint8_t *__restrict__ invalid_bits, const int *__restrict__ frame_idx,
               int m, int h, int w, int radius, char *__restrict__ scratch, size_t scratch_per_frame) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int WB = (w + 31) >> 5, bytes_per_row = (w + 7) >> 3;
    const size_t per_warp = (size_t)h * WB * 4 + (size_t)kHeapSmem * 12 + kMaxTaps * 8;
    unsigned char *base = smem_raw + warp * ((per_warp + 15) & ~(size_t)15);
    uint32_t *F = reinterpret_cast<uint32_t *>(base);
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(base + (((size_t)h * WB * 4 + 7) & ~(size_t)7));
    uint32_t *s_pos = reinterpret_cast<uint32_t *>(s_key + kHeapSmem);
    float *term_w = reinterpret_cast<float *>(s_pos + kHeapSmem);
    float *term_v = term_w + kMaxTaps;
    const int side = 2 * radius + 1, taps = side * side;

    for (int slot = blockIdx.x * kInpaintWarps + warp; slot < m; slot += gridDim.x * kInpaintWarps) {
        const int f = frame_idx ? frame_idx[slot] : slot;
        char *my = scratch + (size_t)slot * scratch_per_frame;
        FrameCtx c;
        c.F = F; c.h = h; c.w = w; c.WB = WB;
        c.out = frames + (size_t)f * h * w;
        c.t = reinterpret_cast<float *>(my);
        unsigned long long *g_key = reinterpret_cast<unsigned long long *>(my + (((size_t)h * w * 4 + 7) & ~(size_t)7));
        uint32_t *g_pos = reinterpret_cast<uint32_t *>(g_key + (size_t)h * w);

        // ---- INSIDE bit rows from the byte-packed invalid mask (bit b of byte B = pixel 8B+b) ----
        const uint8_t *bits = invalid_bits + (size_t)f * h * bytes_per_row;
        for (int idx = lane; idx < h * WB; idx += 32) {
            const int y = idx / WB, k = idx - y * WB;
            uint32_t v = 0;
            for (int b = 0; b < 4; ++b) {
                const int byte = 4 * k + b;
                if (byte < bytes_per_row) v |= (uint32_t)bits[(size_t)y * bytes_per_row + byte] << (8 * b);
            }
            if (k == WB - 1 && (w & 31)) v &= (1u << (w & 31)) - 1u;
            F[idx] = v;
        }
        __syncwarp();

        // ---- band = cross-dilation of INSIDE minus INSIDE; count, init T, fill the heap in raster order ----
        // pass 1: per-row band popcounts -> raster ranks
        int total_band = 0, total_inside = 0;
        {
            int my_cnt = 0;                       // lanes own whole rows round-robin for the counting
            for (int y = lane; y < h; y += 32) {
                for (int k = 0; k < WB; ++k) {
                    const uint32_t x = F[y * WB + k];
                    const uint32_t up = y > 0 ? F[(y - 1) * WB + k] : 0u, dn = y + 1 < h ? F[(y + 1) * WB + k] : 0u;
                    const uint32_t lf = (x << 1) | (k > 0 ? F[y * WB + k - 1] >> 31 : 0u);
                    const uint32_t rt = (x >> 1) | (k + 1 < WB ? F[y * WB + k + 1] << 31 : 0u);
                    uint32_t band = (up | dn | lf | rt) & ~x;
                    if (k == WB - 1 && (w & 31)) band &= (1u << (w & 31)) - 1u;
                    my_cnt += __popc(band);
                    total_inside += __popc(x);
                }
            }
            total_band = warp_sum(my_cnt);
            total_inside = warp_sum(total_inside);
        }
        if (total_inside == 0) { __syncwarp(); continue; }
        const int need = total_band + total_inside;      // every band / INSIDE pixel enters the queue exactly once
        Heap hp;
        hp.key = need <= kHeapSmem ? s_key : g_key;
        hp.pos = need <= kHeapSmem ? s_pos : g_pos;
        hp.size = 0;
        // pass 2 (lane 0, sequential raster order): T init and heap fill; keys with T = 0 and increasing sequence
        // numbers appended in order already form a valid heap
        uint32_t seq = 0;
        if (lane == 0) {
            for (int y = 0; y < h; ++y) {
                for (int k = 0; k < WB; ++k) {
                    const uint32_t x = F[y * WB + k];
                    const uint32_t up = y > 0 ? F[(y - 1) * WB + k] : 0u, dn = y + 1 < h ? F[(y + 1) * WB + k] : 0u;
                    const uint32_t lf = (x << 1) | (k > 0 ? F[y * WB + k - 1] >> 31 : 0u);
                    const uint32_t rt = (x >> 1) | (k + 1 < WB ? F[y * WB + k + 1] << 31 : 0u);
                    uint32_t band = (up | dn | lf | rt) & ~x;
                    if (k == WB - 1 && (w & 31)) band &= (1u << (w & 31)) - 1u;
                    uint32_t ins = x;
                    while (ins) { const int b = __ffs(ins) - 1; ins &= ins - 1; __stcg(c.t + y * w + (k << 5) + b, 1.0e6f); }
                    while (band) {
                        const int b = __ffs(band) - 1; band &= band - 1;
                        const int xx = (k << 5) + b;
                        __stcg(c.t + y * w + xx, 0.0f);
                        hp.key[hp.size] = (unsigned long long)seq++;                 // T = 0
                        hp.pos[hp.size] = ((uint32_t)(y + 1) << 16) | (uint32_t)(xx + 1);   // extended coordinates
                        hp.size++;
                    }
                }
            }
        }
        __syncwarp();

        // ---- fast marching ----
        for (;;) {
            uint32_t packed = 0xffffffffu;
            if (lane == 0 && hp.size > 0) packed = heap_pop(hp);
            packed = __shfl_sync(0xffffffffu, packed, 0);
            if (packed == 0xffffffffu) break;
            const int ii = (int)(packed >> 16), jj = (int)(packed & 0xffffu);
            for (int q = 0; q < 4; ++q) {
                const int i = ii + (q == 0 ? -1 : (q == 2 ? 1 : 0));
                const int j = jj + (q == 1 ? -1 : (q == 3 ? 1 : 0));
                if (i <= 0 || j <= 0 || i > h + 1 || j > w + 1) continue;
                if (!is_inside(c, i, j)) continue;                       // warp-uniform
                float dist = 0.0f;
                if (lane == 0) {
                    dist = fminf(fminf(fmm_solve(c, i - 1, j, i, j - 1), fmm_solve(c, i + 1, j, i, j - 1)),
                                 fminf(fmm_solve(c, i - 1, j, i, j + 1), fmm_solve(c, i + 1, j, i, j + 1)));
                    __stcg(c.t + (i - 1) * w + (j - 1), dist);
                }
                for (int a = lane; a < taps; a += 32) {
                    const int dk = a / side - radius, dl = a % side - radius;
                    float tw, tv;
                    ns_tap(c, i, j, i + dk, j + dl, radius, tw, tv);
                    term_w[a] = tw; term_v[a] = tv;
                }
                __syncwarp();
                if (lane == 0) {
                    float Ia = 0.0f, s = 1.0e-20f;
                    for (int a = 0; a < taps; ++a) { Ia = __fadd_rn(Ia, term_v[a]); s = __fadd_rn(s, term_w[a]); }
                    const double val = __ddiv_rn((double)Ia, (double)s);
                    int iv = __double2int_rn(val);                      // cv::saturate_cast<uchar>(double): cvRound + clamp
                    iv = iv < 0 ? 0 : (iv > 255 ? 255 : iv);
                    __stcg(c.out + (i - 1) * w + (j - 1), (uint8_t)iv);
                    F[(i - 1) * WB + ((j - 1) >> 5)] &= ~(1u << ((j - 1) & 31));        // f = BAND
                    heap_push(hp, ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)seq++,
                              ((uint32_t)i << 16) | (uint32_t)j);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" size_t msq_inpaint_scratch_bytes(int m, int h, int w) {
    if (m <= 0 || h <= 0 || w <= 0) return 0;
    const size_t px = (size_t)h * w;
    const size_t per = align_up(align_up(px * 4, 8) + px * 8 + px * 4, 256);
    return per * (size_t)m;
}

extern "C" int msq_inpaint_frames(uint8_t *frames, const uint8_t *invalid_bits, const int32_t *frame_idx, int m, int h,
                                  int w, int radius, void *scratch, size_t scratch_bytes, void *stream) {
    MSQ_REQUIRE(m >= 0 && h > 0 && w > 0, MSQ_EINVAL, "msq_inpaint_frames: bad sizes m=%d h=%d w=%d", m, h, w);
    if (m == 0) return MSQ_OK;
    MSQ_REQUIRE(frames && invalid_bits, MSQ_EINVAL, "msq_inpaint_frames: null pointer");
    MSQ_REQUIRE(radius >= 1 && radius <= 4, MSQ_EUNSUPPORTED, "msq_inpaint_frames: radius %d not in 1..4", radius);
    MSQ_REQUIRE(h < 65535 && w < 65535, MSQ_EUNSUPPORTED, "msq_inpaint_frames: frame too large");
    MSQ_REQUIRE(scratch && (uintptr_t)scratch % 8 == 0 && scratch_bytes >= msq_inpaint_scratch_bytes(m, h, w), MSQ_ENOMEM,
                "msq_inpaint_frames: scratch must be 8-byte aligned and >= %zu bytes", msq_inpaint_scratch_bytes(m, h, w));
    const int WB = (w + 31) / 32;
    const size_t per_warp = align_up((size_t)h * WB * 4 + (size_t)kHeapSmem * 12 + kMaxTaps * 8, 16);
    const size_t smem = per_warp * kInpaintWarps + 16;
    MSQ_REQUIRE(smem <= 227 * 1024, MSQ_EUNSUPPORTED, "msq_inpaint_frames: %dx%d frames need %zu B of shared memory", h, w, smem);
    MSQ_CUDA_OK(cudaFuncSetAttribute(inpaint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min((m + kInpaintWarps - 1) / kInpaintWarps, sm_count() * 8);
    const size_t per_frame = msq_inpaint_scratch_bytes(1, h, w);
    TimedLaunch timed(K_INPAINT, (cudaStream_t)stream);
    inpaint_kernel<<<grid, kInpaintWarps * 32, smem, (cudaStream_t)stream>>>(frames, invalid_bits, frame_idx, m, h, w, radius,
                                                                          reinterpret_cast<char *>(scratch), per_frame);
    MSQ_LAUNCH_OK("inpaint_frames");
    return MSQ_OK;
}
