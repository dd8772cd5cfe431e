// a7 get_frame_features + im_moment_features: centroid / orientation / ellipse axes of the largest
// outer contour of (cleaned > thr) & mask.   ref proc/proc.py:237-302 and :518-549.
//
// OpenCV's findContours + contourArea + moments(contour) are replaced by an exact, contour-free
// formulation (SURVEY.md section 7 trap 1, pinned against OpenCV in tests/test_oracle_vs_golden.py):
//   * the polygon traced around a hole-free 8-connected blob equals the union of unit squares
//     (2x2 pixel-centre blocks with 4 corners set) and corner triangles (exactly 3 set);
//   * its moments are sums of integer cell integrals (x24), accumulated in int64.
// One WARP owns one frame, kept as bit rows in shared memory (lane k <-> 32-pixel word k of a row):
//   phase 0  threshold+mask -> bit rows (coalesced 128-bit loads, the only HBM traffic)
//   phase 1  4-connected flood of the background from outside, row sweeps with a warp-wide
//            carry-lookahead run fill (ballot + integer add), gives the hole-filled set F
//   phase 2  peel 8-connected blobs of F in raster order; per blob popcount-based cell sums
//   phase 3  float64 epilogue in OpenCV's operation order (lane 0)
#include "common.cuh"
#include "bitrows.cuh"
#include <algorithm>
#include <limits.h>
#include <math.h>

namespace msq {
namespace {

constexpr int kFeatWarps = 4;            // frames per CTA

// x | x<<1 | x>>1 across the whole row (8-connected vertical neighbourhood)
__device__ __forceinline__ uint32_t spread3(uint32_t x, int lane) {
    uint32_t below = __shfl_up_sync(0xffffffffu, x, 1);
    uint32_t above = __shfl_down_sync(0xffffffffu, x, 1);
    if (lane == 0) below = 0u;
    if (lane == 31) above = 0u;
    return x | (x << 1) | (x >> 1) | (below >> 31) | (above << 31);
}

// sum of set-bit positions and of their squares in a 32-bit word
__device__ __forceinline__ void bit_moments(uint32_t x, int &n, int &s1, int &s2) {
    const int p0 = __popc(x & 0xAAAAAAAAu), p1 = __popc(x & 0xCCCCCCCCu), p2 = __popc(x & 0xF0F0F0F0u),
              p3 = __popc(x & 0xFF00FF00u), p4 = __popc(x & 0xFFFF0000u);
    n = __popc(x);
    s1 = p0 + 2 * p1 + 4 * p2 + 8 * p3 + 16 * p4;
    int cross = 0;
    const uint32_t m[5] = {0xAAAAAAAAu, 0xCCCCCCCCu, 0xF0F0F0F0u, 0xFF00FF00u, 0xFFFF0000u};
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = a + 1; b < 5; ++b) cross += __popc(x & m[a] & m[b]) << (a + b + 1);
    s2 = p0 + 4 * p1 + 16 * p2 + 64 * p3 + 256 * p4 + cross;
}

struct CellAcc {            // per-lane accumulators of one cell class
    long long n, i, j, ii, ij, jj;
};

// 24 x (A, U, V, UU, UV, VV): area integrals of the covered part of a unit cell, cell-local coords
__device__ __constant__ int kCell24[5][6] = {
    {24, 12, 12, 8, 6, 8},    // all four corners
    {12, 8, 8, 6, 5, 6},      // top-left missing
    {12, 4, 8, 2, 3, 6},      // top-right missing
    {12, 8, 4, 6, 3, 2},      // bottom-left missing
    {12, 4, 4, 2, 1, 2},      // bottom-right missing
};

// OpenCV contourMoments/completeMomentState + ref proc/proc.py:529-547, float64, no contraction
__device__ void moment_epilogue(const long long s[6], double *centroid, double *orientation, double *axis) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double a00 = (double)s[0] / 12.0;
    if (!(fabs(a00) > 1.1920928955078125e-07)) {
        centroid[0] = centroid[1] = nan; *orientation = nan; axis[0] = axis[1] = nan;
        return;
    }
    const double a10 = (double)s[1] / 4.0, a01 = (double)s[2] / 4.0;
    const double a20 = (double)s[3] / 2.0, a11 = (double)s[4], a02 = (double)s[5] / 2.0;
    const double m00 = __dmul_rn(a00, 0.5);
    const double m10 = __dmul_rn(a10, 0.16666666666666666666666666666667);
    const double m01 = __dmul_rn(a01, 0.16666666666666666666666666666667);
    const double m20 = __dmul_rn(a20, 0.083333333333333333333333333333333);
    const double m11 = __dmul_rn(a11, 0.041666666666666666666666666666667);
    const double m02 = __dmul_rn(a02, 0.083333333333333333333333333333333);
    const double inv = __ddiv_rn(1.0, m00);
    const double cx = __dmul_rn(m10, inv), cy = __dmul_rn(m01, inv);
    const double mu20 = __dsub_rn(m20, __dmul_rn(m10, cx));
    const double mu11 = __dsub_rn(m11, __dmul_rn(m10, cy));
    const double mu02 = __dsub_rn(m02, __dmul_rn(m01, cy));
    const double num = __dmul_rn(2.0, mu11);
    const double den = __dsub_rn(mu20, mu02);
    const double common = __dsqrt_rn(__dadd_rn(__dmul_rn(4.0, __dmul_rn(mu11, mu11)), __dmul_rn(den, den)));
    *orientation = __dmul_rn(-0.5, atan2(num, den));
    centroid[0] = __ddiv_rn(m10, m00);
    centroid[1] = __ddiv_rn(m01, m00);
    const double two_root2 = 2.8284271247461903;        // 2*np.sqrt(2)
    const double tr = __dadd_rn(mu20, mu02);
    axis[0] = __dmul_rn(two_root2, __dsqrt_rn(__ddiv_rn(__dadd_rn(tr, common), m00)));
    axis[1] = __dmul_rn(two_root2, __dsqrt_rn(__ddiv_rn(__dsub_rn(tr, common), m00)));
}

// bits 7,15,23,31 -> bits 0..3
__device__ __forceinline__ uint32_t gather_nibble(uint32_t on) {
    return (((on >> 7) * 0x01020408u) >> 24) & 0xfu;
}

struct RowWordRaw { uint4 c0, m0, c1, m1; };     // the 32 pixels of one (row, word): two 16-byte halves of each image

__device__ __forceinline__ void load_row_word(RowWordRaw &raw, const uint8_t *__restrict__ crow,
                                              const uint8_t *__restrict__ mrow, int k, int w) {
    const int x = k << 5;
    raw.c0 = ldg_stream_u4(crow + x);
    raw.m0 = ldg_stream_u4(mrow + x);
    if (x + 16 < w) {
        raw.c1 = ldg_stream_u4(crow + x + 16);
        raw.m1 = ldg_stream_u4(mrow + x + 16);
    } else {
        raw.c1 = raw.m1 = make_uint4(0u, 0u, 0u, 0u);
    }
}

__device__ __forceinline__ uint32_t threshold_raw(const RowWordRaw &raw, const ByteTest &t) {
    const uint32_t cw[8] = {raw.c0.x, raw.c0.y, raw.c0.z, raw.c0.w, raw.c1.x, raw.c1.y, raw.c1.z, raw.c1.w};
    const uint32_t mw[8] = {raw.m0.x, raw.m0.y, raw.m0.z, raw.m0.w, raw.m1.x, raw.m1.y, raw.m1.z, raw.m1.w};
    // the instance mask covers a few per cent of the frame: 32 pixels without a single mask byte need no thresholding,
    // and a row step in which that holds for all 32 lanes skips the ~100 SWAR instructions altogether
    if (((mw[0] | mw[1]) | (mw[2] | mw[3]) | (mw[4] | mw[5]) | (mw[6] | mw[7])) == 0u) return 0u;
    uint32_t bits = 0u;
#pragma unroll
    for (int q = 0; q < 8; ++q) bits |= gather_nibble(bytes_ge(cw[q], t) & bytes_nonzero(mw[q])) << (4 * q);
    return bits;
}

// generic (unaligned) form: (cleaned >= ge) & (mask != 0) for pixels [32k, 32k+32) of one row
__device__ __forceinline__ uint32_t threshold_word_scalar(const uint8_t *__restrict__ crow, const uint8_t *__restrict__ mrow,
                                                          int k, int w, int ge) {
    uint32_t bits = 0u;
    const int x0 = k << 5;
    for (int b = 0; b < 32 && x0 + b < w; ++b)
        bits |= (uint32_t)(((int)crow[x0 + b] >= ge) && (mrow[x0 + b] != 0)) << b;
    return bits;
}

template <int LPR>      // lanes per row: power of two >= words per row
__global__ void __launch_bounds__(kFeatWarps * 32)
features_kernel(const uint8_t *__restrict__ cleaned, const uint8_t *__restrict__ mask, int n, int h, int w,
                int ge, double *__restrict__ centroid, double *__restrict__ orientation,
                double *__restrict__ axis_length, long long *__restrict__ sums24, const int *__restrict__ list) {
    extern __shared__ __align__(16) uint32_t smem_bits[];
    constexpr int RPW = 32 / LPR;                       // rows handled per warp step in the parallel phases
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpr = (w + 31) >> 5;                      // words per row actually used
    const int sub = lane / LPR, k = lane % LPR;         // (row within step, word) mapping of the parallel phases
    uint32_t *P = smem_bits + (size_t)warp * 2 * h * LPR;   // fm, later the current blob
    uint32_t *Q = P + (size_t)h * LPR;                       // reach, later the remaining set
    const bool vec_ok = (w % 16 == 0) && ((uintptr_t)cleaned % 16 == 0) && ((uintptr_t)mask % 16 == 0);
    const uint32_t lane_mask_row =                      // valid pixel bits of word `lane` (sequential phases)
        lane < wpr - 1 ? 0xffffffffu : (lane == wpr - 1 ? ((w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu) : 0u);

    // with a list (written by features_stream_kernel: [0] = count, [1..] = frame indices) only those frames are processed
    const int todo = list ? list[0] : n;
    for (int item = blockIdx.x * kFeatWarps + warp; item < todo; item += gridDim.x * kFeatWarps) {
        const int f = list ? list[1 + item] : item;
        const uint8_t *cf = cleaned + (size_t)f * h * w;
        const uint8_t *mf = mask + (size_t)f * h * w;

        // ---------------- phase 0: bit rows ----------------
        // A row step is 2 KB per warp (every lane: 32 B of the frame + 32 B of the mask).  kStages steps are kept in flight
        // with cp.async into the (still unused) Q plane -- no registers are tied up, so three steps ride the HBM latency
        // instead of one (ncu on the register-prefetch version: 53 % of the stall samples sat on this loop, DRAM 36 % busy).
        int r_lo = h, r_hi = -1;
        constexpr int kStages = 3;
        const bool staged = vec_ok && (size_t)h * LPR * sizeof(uint32_t) >= (size_t)kStages * 2048;
        if (staged) {
            const ByteTest bt = make_byte_test(ge);
            const bool lane_on = k < wpr;
            const int x = k << 5;
            const bool second = x + 16 < w;
            unsigned char *ring = reinterpret_cast<unsigned char *>(Q);      // [stage][part 0..3][lane] x 16 B
            const uint32_t ring_sa = (uint32_t)__cvta_generic_to_shared(ring) + lane * 16;
            auto issue = [&](int r, int stage) {
                if (lane_on && r < h) {
                    const uint8_t *c = cf + (size_t)r * w + x, *m = mf + (size_t)r * w + x;
                    const uint32_t dst = ring_sa + stage * 2048;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(c) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 1024), "l"(m) : "memory");
                    if (second) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 512), "l"(c + 16) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + 1536), "l"(m + 16) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
#pragma unroll
            for (int sgl = 0; sgl < kStages; ++sgl) issue(sub + sgl * RPW, sgl);
            int stage = 0;
#pragma unroll 3
            for (int r = sub; r < h; r += RPW) {
                asm volatile("cp.async.wait_group %0;" :: "n"(kStages - 1) : "memory");
                uint32_t bits = 0u;
                if (lane_on) {
                    const uint4 *slot = reinterpret_cast<const uint4 *>(ring + stage * 2048) + lane;
                    RowWordRaw cur;
                    cur.c0 = slot[0];
                    cur.m0 = slot[64];
                    if (second) { cur.c1 = slot[32]; cur.m1 = slot[96]; }
                    else cur.c1 = cur.m1 = make_uint4(0u, 0u, 0u, 0u);
                    bits = threshold_raw(cur, bt);
                    if (x + 32 > w) bits &= (w - x >= 32) ? 0xffffffffu : ((1u << (w - x)) - 1u);   // drop pixels beyond the row end
                }
                P[r * LPR + k] = bits;
                if (bits) { r_lo = min(r_lo, r); r_hi = max(r_hi, r); }
                issue(r + kStages * RPW, stage);                            // refill the slot that was just consumed
                stage = (stage + 1 == kStages) ? 0 : stage + 1;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else if (vec_ok) {
            // small frames (the Q plane cannot hold the ring): one step of register prefetch
            const ByteTest bt = make_byte_test(ge);
            const bool lane_on = k < wpr;
            RowWordRaw cur, nxt;
            if (lane_on && sub < h) load_row_word(cur, cf + (size_t)sub * w, mf + (size_t)sub * w, k, w);
            for (int r = sub; r < h; r += RPW) {
                const int rn = r + RPW;
                if (lane_on && rn < h) load_row_word(nxt, cf + (size_t)rn * w, mf + (size_t)rn * w, k, w);
                uint32_t bits = 0u;
                if (lane_on) {
                    bits = threshold_raw(cur, bt);
                    const int x0 = k << 5;                                  // drop pixels beyond the row end
                    if (x0 + 32 > w) bits &= (w - x0 >= 32) ? 0xffffffffu : ((1u << (w - x0)) - 1u);
                }
                P[r * LPR + k] = bits;
                if (bits) { r_lo = min(r_lo, r); r_hi = max(r_hi, r); }
                cur = nxt;
            }
        } else {
            for (int r = sub; r < h; r += RPW) {
                uint32_t bits = 0u;
                if (k < wpr) bits = threshold_word_scalar(cf + (size_t)r * w, mf + (size_t)r * w, k, w, ge);
                P[r * LPR + k] = bits;
                if (bits) { r_lo = min(r_lo, r); r_hi = max(r_hi, r); }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            r_lo = min(r_lo, __shfl_xor_sync(0xffffffffu, r_lo, o));
            r_hi = max(r_hi, __shfl_xor_sync(0xffffffffu, r_hi, o));
        }
        __syncwarp();

        long long best[6] = {0, 0, 0, 0, 0, 0};
        bool have = false;

        // exact cell sums of the blob held in P (rows r0..rmax), lanes = (row-in-step, word)
        auto blob_sums = [&](int r0, int rmax, long long (&s)[6]) {
        CellAcc acc[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[c] = {0, 0, 0, 0, 0, 0};
        for (int r = r0 + sub; r < rmax; r += RPW) {
            if (k >= wpr) continue;
            const uint32_t a = P[r * LPR + k], b = P[(r + 1) * LPR + k];
            const uint32_t an = (k + 1 < wpr) ? P[r * LPR + k + 1] : 0u;
            const uint32_t bn = (k + 1 < wpr) ? P[(r + 1) * LPR + k + 1] : 0u;
            if ((a | b) == 0u) continue;
            const uint32_t a1 = (a >> 1) | (an << 31), b1 = (b >> 1) | (bn << 31);
            const uint32_t cls[5] = {a & a1 & b & b1, ~a & a1 & b & b1, a & ~a1 & b & b1,
                                     a & a1 & ~b & b1, a & a1 & b & ~b1};
            const long long x0 = (long long)k << 5, y = r;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                if (cls[c] == 0u) continue;
                int cnt, s1, s2;
                bit_moments(cls[c], cnt, s1, s2);
                const long long si = x0 * cnt + s1;
                const long long sii = x0 * x0 * cnt + 2 * x0 * s1 + s2;
                acc[c].n += cnt;
                acc[c].i += si;
                acc[c].j += y * cnt;
                acc[c].ii += sii;
                acc[c].ij += y * si;
                acc[c].jj += y * y * cnt;
            }
        }
        for (int q = 0; q < 6; ++q) s[q] = 0;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const long long A = kCell24[c][0], U = kCell24[c][1], V = kCell24[c][2], UU = kCell24[c][3],
                            UV = kCell24[c][4], VV = kCell24[c][5];
            s[0] += A * acc[c].n;
            s[1] += A * acc[c].i + U * acc[c].n;
            s[2] += A * acc[c].j + V * acc[c].n;
            s[3] += A * acc[c].ii + 2 * U * acc[c].i + UU * acc[c].n;
            s[4] += A * acc[c].ij + V * acc[c].i + U * acc[c].j + UV * acc[c].n;
            s[5] += A * acc[c].jj + 2 * V * acc[c].j + VV * acc[c].n;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) s[q] = warp_sum_ll(s[q]);
        };

        // ---------------- fast path: a row-convex blob ----------------
        // If every row of [r_lo, r_hi] holds exactly one run and consecutive runs touch (8-connected), the foreground is one
        // component, every background pixel reaches the left or right image edge along its own row (no holes), and the
        // general machinery below -- background flood, peeling, 3 + 3 sequential sweeps over the rows -- would return the
        // foreground itself.  The test is one parallel pass; mouse-shaped masks pass it almost always.
        bool simple = r_hi >= 0;
        if (simple) {
            for (int base = r_lo; base <= r_hi; base += RPW) {
                const int r = base + sub;
                const bool in = r <= r_hi && k < wpr;
                const uint32_t v = in ? P[r * LPR + k] : 0u;
                const uint32_t nxt = (in && r < r_hi) ? P[(r + 1) * LPR + k] : 0u;
                uint32_t lo_nb = __shfl_up_sync(0xffffffffu, v, 1), hi_nb = __shfl_down_sync(0xffffffffu, v, 1);
                if (k == 0) lo_nb = 0u;
                if (k == LPR - 1) hi_nb = 0u;
                int starts = __popc(v & ~((v << 1) | (lo_nb >> 31)));                       // run starts in this word
                int touch = ((v | (v << 1) | (v >> 1) | (lo_nb >> 31) | (hi_nb << 31)) & nxt) != 0u;
#pragma unroll
                for (int o = LPR >> 1; o > 0; o >>= 1) {
                    starts += __shfl_xor_sync(0xffffffffu, starts, o);
                    touch |= __shfl_xor_sync(0xffffffffu, touch, o);
                }
                const bool row_ok = r > r_hi || (starts == 1 && (r == r_hi || touch));
                simple = __all_sync(0xffffffffu, row_ok) && simple;
            }
        }


        if (simple) {
            blob_sums(r_lo, r_hi, best);
            have = true;
        } else if (r_hi >= 0) {
            // the general path needs the Q plane (phase 0 may have used it as its load ring): nothing reached yet
            for (int r = sub; r < h; r += RPW) Q[r * LPR + k] = 0u;
            __syncwarp();
            // ---------------- phase 1: flood the background from outside (4-connected) ----------------
            const uint32_t edge = (lane == 0 ? 1u : 0u) | (lane == ((w - 1) >> 5) ? (1u << ((w - 1) & 31)) : 0u);
            const bool act = lane < LPR;
            bool down = true;
            for (int sweep = 0;; ++sweep) {
                bool changed = false;
                uint32_t prev = lane_mask_row;                      // the row outside [r_lo,r_hi] is all reached
                const int r_begin = down ? r_lo : r_hi, r_end = down ? r_hi + 1 : r_lo - 1, dr = down ? 1 : -1;
                for (int r = r_begin; r != r_end; r += dr) {
                    const uint32_t open = act ? (~P[r * LPR + lane] & lane_mask_row) : 0u;
                    const uint32_t old = act ? Q[r * LPR + lane] : 0u;
                    const uint32_t now = fill_row((prev | old | edge) & open, open, lane);
                    if (act) Q[r * LPR + lane] = now;
                    changed |= (now != old);
                    prev = now;
                }
                // a sweep that changes nothing after a sweep in the other direction means closure
                if (!__any_sync(0xffffffffu, changed) && sweep > 0) break;
                down = !down;
            }
            // F = everything not reached (foreground + enclosed holes) is the "remaining" set of phase 2.  It is never
            // materialised: Q keeps meaning "reached by the background flood OR already peeled", and phase 2 reads
            // ~Q & lane_mask_row (a separate complement pass over the rows was 14 % of the kernel's instructions).
            __syncwarp();

            // ---------------- phase 2: peel blobs in raster order ----------------
            int scan = r_lo;
            while (true) {
                // raster-first remaining pixel
                int r0 = -1; uint32_t seed = 0u;
                for (int r = scan; r <= r_hi; ++r) {
                    const uint32_t v = act ? (~Q[r * LPR + lane] & lane_mask_row) : 0u;
                    const uint32_t b = __ballot_sync(0xffffffffu, v != 0u);
                    if (b) {
                        const int kw = __ffs(b) - 1;
                        const uint32_t word = __shfl_sync(0xffffffffu, v, kw);
                        seed = (lane == kw) ? (1u << (__ffs(word) - 1)) : 0u;
                        r0 = r;
                        break;
                    }
                }
                if (r0 < 0) break;
                scan = r0;

                // 8-connected flood of the blob inside the remaining set; blob rows are [r0, rmax]
                int rmax = r0;
                {
                    const uint32_t rem0 = act ? (~Q[r0 * LPR + lane] & lane_mask_row) : 0u;
                    uint32_t prev = fill_row(seed, rem0, lane);
                    if (act) P[r0 * LPR + lane] = prev;
                    bool go_down = true;
                    for (int sweep = 0;; ++sweep) {
                        bool changed = false;
                        if (go_down) {
                            prev = act ? P[r0 * LPR + lane] : 0u;
                            for (int r = r0 + 1; r <= r_hi; ++r) {
                                const uint32_t rem = act ? (~Q[r * LPR + lane] & lane_mask_row) : 0u;
                                const uint32_t old = (act && r <= rmax) ? P[r * LPR + lane] : 0u;
                                const uint32_t now = fill_row((spread3(prev, lane) | old) & rem, rem, lane);
                                const bool any_now = __any_sync(0xffffffffu, now != 0u);
                                if (!any_now && r > rmax) break;
                                if (act) P[r * LPR + lane] = now;
                                if (any_now) rmax = max(rmax, r);
                                changed |= (now != old);
                                prev = now;
                            }
                        } else {
                            prev = act ? P[rmax * LPR + lane] : 0u;
                            for (int r = rmax - 1; r >= r0; --r) {
                                const uint32_t rem = act ? (~Q[r * LPR + lane] & lane_mask_row) : 0u;
                                const uint32_t old = act ? P[r * LPR + lane] : 0u;
                                const uint32_t now = fill_row((spread3(prev, lane) | old) & rem, rem, lane);
                                if (act) P[r * LPR + lane] = now;
                                changed |= (now != old);
                                prev = now;
                            }
                        }
                        const bool any_change = __any_sync(0xffffffffu, changed);
                        if (sweep > 0 && !any_change) break;
                        if (sweep == 0 && rmax == r0) break;          // single-row blob
                        go_down = !go_down;
                    }
                }
                __syncwarp();

                long long s[6];
                blob_sums(r0, rmax, s);
                // OpenCV lists sibling contours in reverse raster order and np.argmax keeps the first
                // maximum, so a later blob with an equal area replaces the current best
                if (!have || s[0] >= best[0]) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) best[q] = s[q];
                    have = true;
                }
                // remove the blob from the remaining set
                for (int r = r0 + sub; r <= rmax; r += RPW)
                    if (k < wpr) Q[r * LPR + k] |= P[r * LPR + k];
                __syncwarp();
            }
        }

        // ---------------- phase 3: float64 epilogue ----------------
        if (lane == 0) {
            double c[2], o, a[2];
            moment_epilogue(best, c, &o, a);
            centroid[2 * (size_t)f] = c[0];
            centroid[2 * (size_t)f + 1] = c[1];
            orientation[f] = o;
            axis_length[2 * (size_t)f] = a[0];
            axis_length[2 * (size_t)f + 1] = a[1];
            if (sums24)
                for (int q = 0; q < 6; ++q) sums24[6 * (size_t)f + q] = best[q];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming fast path.  For a row-convex foreground (every row one run, consecutive runs touching) the answer is the cell
// sums of the foreground itself, and both the test and the sums only ever look at a row and the row above it.  So the frame
// is consumed in ONE pass straight from global memory: lane (sub, k) thresholds word k of row base + sub, the row above
// comes from the neighbouring lane group (or, for the first row of a step, from the previous step's registers), run counts
// and contacts are popcounts + shuffles, and nothing touches shared memory.  ~100 registers and no shared memory give
// 16+ warps per SM instead of the 12 the general kernel gets.  Frames that fail the test are appended to `fallback`
// ([0] = count, [1..] = frame indices) and redone by features_kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStreamWarps = 4;
constexpr int kChunkSteps = 2;         // row steps per bulk copy: 8 rows x 240 B = 1.9 KB contiguous per array.  (12 rows were as good while whole frames were read; with the row bands of the cleaning pass shorter chunks waste less at the band ends: +2 %; 4 rows: no better)
constexpr int kStreamStages = 2;       // bulk copies in flight per warp

// ---- mbarrier / bulk-copy (TMA, 1-D) primitives ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
                    "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

template <int LPR>
__global__ void __launch_bounds__(kStreamWarps * 32, 4)
features_stream_kernel(const uint8_t *__restrict__ cleaned, const uint8_t *__restrict__ mask, int n, int h, int w, int ge,
                       double *__restrict__ centroid, double *__restrict__ orientation, double *__restrict__ axis_length,
                       long long *__restrict__ sums24, int *__restrict__ fallback, RowBands rows) {
    constexpr int RPW = 32 / LPR;
    constexpr int kChunkRows = kChunkSteps * RPW;
    constexpr unsigned kAll = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char stream_smem[];
    __shared__ uint64_t bars[kStreamWarps][kStreamStages];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpr = (w + 31) >> 5;
    const int sub = lane / LPR, k = lane % LPR;
    const bool lane_on = k < wpr;
    const int x = k << 5;
    const bool second = x + 16 < w;
    const uint32_t tail = (x + 32 > w) ? ((w - x >= 32) ? 0xffffffffu : ((w - x) > 0 ? ((1u << (w - x)) - 1u) : 0u)) : 0xffffffffu;
    const ByteTest bt = make_byte_test(ge);
    // per warp: kStreamStages slots of [kChunkRows rows of the frame | kChunkRows rows of the mask], filled by 1-D bulk copies
    const size_t slot_bytes = (size_t)2 * kChunkRows * w;
    unsigned char *slots = stream_smem + (size_t)warp * kStreamStages * slot_bytes;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kStreamStages; ++st) mbar_init(&bars[warp][st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int first = blockIdx.x * kStreamWarps + warp, stride = gridDim.x * kStreamWarps;
    const int n_chunks = (h + kChunkRows - 1) / kChunkRows;
    const int my_frames = first < n ? (n - first + stride - 1) / stride : 0;
    // Chunks [c0, c1) of frame number fi of this warp that have to be read.  With the row bands of the cleaning pass (the
    // opened frame is zero outside them, so is the thresholded foreground) that is ~1/3 of the frame; rows of a chunk that lie
    // outside the band are zero in memory as well, so whole chunks are read as they are.
    auto chunk_span = [&](int fi, int &c0, int &c1) {
        c0 = 0; c1 = fi < my_frames ? n_chunks : 0;
        if (rows.bands == nullptr || fi >= my_frames) return;
        const int2 *b = rows.bands + (size_t)(first + fi * stride) * rows.tiles_x;
        int lo = INT_MAX, hi = -1;
        for (int t = 0; t < rows.tiles_x; ++t) {
            const int2 v = __ldg(b + t);
            if (v.y >= 0) { lo = min(lo, v.x); hi = max(hi, v.y); }
        }
        if (hi < 0) { c1 = 0; return; }
        c0 = max(0, lo - 4) / kChunkRows;
        c1 = (min(h, hi + 5) + kChunkRows - 1) / kChunkRows;
    };
    // the bulk copies run kStreamStages items ahead of the consumer, across frame boundaries (warp-uniform cursor, lane 0 issues)
    int iss_fi = 0, iss_c, iss_c1;
    long long iss_item = 0;
    chunk_span(0, iss_c, iss_c1);
    auto issue_next = [&]() {
        while (iss_fi < my_frames && iss_c >= iss_c1) { ++iss_fi; chunk_span(iss_fi, iss_c, iss_c1); }
        if (iss_fi >= my_frames) return;
        if (lane == 0) {
            const int f = first + iss_fi * stride, st = (int)(iss_item % kStreamStages);
            const int nrows = min(kChunkRows, h - iss_c * kChunkRows);
            const uint32_t bytes = (uint32_t)nrows * w;
            const size_t off = (size_t)f * h * w + (size_t)iss_c * kChunkRows * w;
            unsigned char *slot = slots + (size_t)st * slot_bytes;
            mbar_expect_tx(&bars[warp][st], 2 * bytes);
            bulk_g2s(slot, cleaned + off, bytes, &bars[warp][st]);
            bulk_g2s(slot + (size_t)kChunkRows * w, mask + off, bytes, &bars[warp][st]);
        }
        ++iss_c; ++iss_item;
    };
    for (int i = 0; i < kStreamStages; ++i) issue_next();
    uint32_t parity = 0u;                                                // bit st: phase of stage st

    long long item = 0;
    for (int fi = 0; fi < my_frames; ++fi) {
        const int f = first + fi * stride;
        CellAcc acc[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[c] = {0, 0, 0, 0, 0, 0};
        bool simple = true, started = false, ended = false, prev_ne = false;      // warp-uniform
        int my_starts = 0, ne_rows = 0;                                            // run starts seen by this lane; non-empty rows
        uint32_t prev_bits = 0u;
        int c_first, c_end;
        chunk_span(fi, c_first, c_end);
        for (int c = c_first; c < c_end; ++c, ++item) {
            const int st = (int)(item % kStreamStages);
            mbar_wait(&bars[warp][st], (parity >> st) & 1u);
            parity ^= 1u << st;
            const unsigned char *slot = slots + (size_t)st * slot_bytes;
            if (simple) {
#pragma unroll
                for (int j = 0; j < kChunkSteps; ++j) {
                    const int rr = j * RPW + sub;                        // row inside the chunk
                    const int r = c * kChunkRows + rr;
                    uint32_t b = 0u;
                    if (lane_on && r < h) {
                        RowWordRaw cur;
                        const unsigned char *crow = slot + (size_t)rr * w + x, *mrow = crow + (size_t)kChunkRows * w;
                        cur.c0 = *reinterpret_cast<const uint4 *>(crow);
                        cur.m0 = *reinterpret_cast<const uint4 *>(mrow);
                        if (second) {
                            cur.c1 = *reinterpret_cast<const uint4 *>(crow + 16);
                            cur.m1 = *reinterpret_cast<const uint4 *>(mrow + 16);
                        } else {
                            cur.c1 = cur.m1 = make_uint4(0u, 0u, 0u, 0u);
                        }
                        b = threshold_raw(cur, bt) & tail;
                    }
                    // word k of the row above: the lane group below this one, or the last row of the previous step
                    const uint32_t up = __shfl_up_sync(kAll, b, LPR & 31);
                    const uint32_t carry = __shfl_sync(kAll, prev_bits, (RPW - 1) * LPR + k);
                    const uint32_t a = (sub == 0) ? carry : up;
                    uint32_t a_lo = __shfl_up_sync(kAll, a, 1), a_hi = __shfl_down_sync(kAll, a, 1);
                    uint32_t b_lo = __shfl_up_sync(kAll, b, 1), b_hi = __shfl_down_sync(kAll, b, 1);
                    if (k == 0) { a_lo = 0u; b_lo = 0u; }
                    if (k == LPR - 1) { a_hi = 0u; b_hi = 0u; }
                    // ---- row-convexity: runs in this row, contact with the row above.  Two ballots give, for all RPW rows
                    // of the step at once, which rows are non-empty and which touch the row above; run starts are only
                    // counted per lane and compared with the number of non-empty rows once per frame (every non-empty row
                    // has >= 1 start, so the totals agree iff every such row is a single run). ----
                    my_starts += __popc(b & ~((b << 1) | (b_lo >> 31)));
                    const unsigned nz = __ballot_sync(kAll, b != 0u);
                    const unsigned tz = __ballot_sync(kAll, ((a | (a << 1) | (a >> 1) | (a_lo >> 31) | (a_hi << 31)) & b) != 0u);
                    constexpr unsigned kGroup = (LPR == 32) ? 0xffffffffu : ((1u << LPR) - 1u);
#pragma unroll
                    for (int q = 0; q < RPW; ++q) {
                        const bool ne = ((nz >> (q * LPR)) & kGroup) != 0u;     // rows >= h hold no bits
                        if (ne) {
                            if (ended || (prev_ne && ((tz >> (q * LPR)) & kGroup) == 0u)) simple = false;
                            started = true;
                            ++ne_rows;
                        } else if (started) {
                            ended = true;
                        }
                        prev_ne = ne;
                    }
                    // ---- exact cell sums of the row pair (r - 1, r), as in features_kernel ----
                    if ((a | b) != 0u) {
                        const uint32_t a1 = (a >> 1) | (a_hi << 31), b1 = (b >> 1) | (b_hi << 31);
                        const uint32_t cls[5] = {a & a1 & b & b1, ~a & a1 & b & b1, a & ~a1 & b & b1, a & a1 & ~b & b1,
                                                 a & a1 & b & ~b1};
                        const long long x0 = (long long)x, y = r - 1;
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            if (cls[q] == 0u) continue;
                            int cnt, s1, s2;
                            bit_moments(cls[q], cnt, s1, s2);
                            const long long si = x0 * cnt + s1;
                            const long long sii = x0 * x0 * cnt + 2 * x0 * s1 + s2;
                            acc[q].n += cnt;
                            acc[q].i += si;
                            acc[q].j += y * cnt;
                            acc[q].ii += sii;
                            acc[q].ij += y * si;
                            acc[q].jj += y * y * cnt;
                        }
                    }
                    prev_bits = b;
                }
            }
            __syncwarp();                                                // every lane is done with the slot
            issue_next();
        }
        if (simple) simple = warp_sum(my_starts) == ne_rows;
        if (!simple) {
            if (lane == 0) fallback[1 + atomicAdd(fallback, 1)] = f;
            continue;
        }
        long long s[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const long long A = kCell24[c][0], U = kCell24[c][1], V = kCell24[c][2], UU = kCell24[c][3],
                            UV = kCell24[c][4], VV = kCell24[c][5];
            s[0] += A * acc[c].n;
            s[1] += A * acc[c].i + U * acc[c].n;
            s[2] += A * acc[c].j + V * acc[c].n;
            s[3] += A * acc[c].ii + 2 * U * acc[c].i + UU * acc[c].n;
            s[4] += A * acc[c].ij + V * acc[c].i + U * acc[c].j + UV * acc[c].n;
            s[5] += A * acc[c].jj + 2 * V * acc[c].j + VV * acc[c].n;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) s[q] = warp_sum_ll(s[q]);
        if (lane == 0) {
            double c[2], o, ax[2];
            moment_epilogue(s, c, &o, ax);
            centroid[2 * (size_t)f] = c[0];
            centroid[2 * (size_t)f + 1] = c[1];
            orientation[f] = o;
            axis_length[2 * (size_t)f] = ax[0];
            axis_length[2 * (size_t)f + 1] = ax[1];
            if (sums24)
                for (int q = 0; q < 6; ++q) sums24[6 * (size_t)f + q] = s[q];
        }
    }
}

template <int LPR>
int launch_features(const uint8_t *cleaned, const uint8_t *mask, int n, int h, int w, int ge,
                    double *centroid, double *orientation, double *axis, long long *sums24, int *fallback, cudaStream_t st,
                    cudaEvent_t after_stream, RowBands rows) {
    const size_t smem = (size_t)kFeatWarps * 2 * h * LPR * sizeof(uint32_t);
    MSQ_REQUIRE(smem <= 227 * 1024, MSQ_EUNSUPPORTED,
                "frame_features: %dx%d frames need %zu B of shared memory per CTA (max 232448)", h, w, smem);
    MSQ_CUDA_OK(cudaFuncSetAttribute(features_kernel<LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = std::max(1, std::min(16, (int)((227 * 1024) / (smem + 1024))));
    const int ctas = (n + kFeatWarps - 1) / kFeatWarps;
    const int grid = std::min(ctas, sm_count() * per_sm);
    // streaming fast path first (needs 16-byte aligned rows and a place for the list of frames it could not settle)
    const bool stream_ok = fallback && (w % 16 == 0) && ((uintptr_t)cleaned % 16 == 0) && ((uintptr_t)mask % 16 == 0) &&
                           (size_t)kStreamWarps * kStreamStages * 2 * (kChunkSteps * (32 / LPR)) * w <= 160 * 1024;
    if (stream_ok) {
        MSQ_CUDA_OK(cudaMemsetAsync(fallback, 0, sizeof(int), st));
        TimedLaunch timed(K_FEATURES, st);
        const size_t ssmem = (size_t)kStreamWarps * kStreamStages * 2 * (kChunkSteps * (32 / LPR)) * w;
        MSQ_CUDA_OK(cudaFuncSetAttribute(features_stream_kernel<LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
        const int sgrid = std::min((n + kStreamWarps - 1) / kStreamWarps, sm_count() * 4);
        features_stream_kernel<LPR><<<sgrid, kStreamWarps * 32, ssmem, st>>>(cleaned, mask, n, h, w, ge, centroid, orientation,
                                                                            axis, sums24, fallback, rows);
        MSQ_LAUNCH_OK("frame_features (streaming)");
    }
    // The general kernel only has the few frames of the list to do, each a long sequential chain in one warp: a short,
    // latency-bound launch.  A caller with independent bandwidth-bound work (the whole-chunk pipeline: the masked sums) passes
    // an event, recorded here between the two launches, and starts that work on another stream behind it.
    if (after_stream) MSQ_CUDA_OK(cudaEventRecord(after_stream, st));
    {
        TimedLaunch timed(K_FEATURES, st);          // a launch of its own in the counters
        features_kernel<LPR><<<grid, kFeatWarps * 32, smem, st>>>(cleaned, mask, n, h, w, ge, centroid, orientation,
                                                                axis, sums24, stream_ok ? fallback : nullptr);
        MSQ_LAUNCH_OK("frame_features");
    }
    return MSQ_OK;
}

}  // namespace

int launch_frame_features(const uint8_t *cleaned, const uint8_t *mask, int n, int h, int w, double frame_threshold,
                          double *centroid, double *orientation, double *axis, int64_t *sums24, int *fallback,
                          cudaStream_t st, cudaEvent_t after_stream, RowBands rows) {
    MSQ_REQUIRE(w <= 1024, MSQ_EUNSUPPORTED, "frame_features: width %d > 1024 is not supported", w);
    // pixel > thr on integers  <=>  pixel >= floor(thr) + 1; clamp to [0, 256] (0: all pass, 256: none)
    int ge;
    if (!(frame_threshold >= -1.0)) ge = 0;
    else if (frame_threshold >= 255.0) ge = 256;
    else ge = (int)floor(frame_threshold) + 1;
    const int wpr = (w + 31) / 32;
    long long *s24 = reinterpret_cast<long long *>(sums24);
    if (wpr <= 1) return launch_features<1>(cleaned, mask, n, h, w, ge, centroid, orientation, axis, s24, fallback, st, after_stream, rows);
    if (wpr <= 2) return launch_features<2>(cleaned, mask, n, h, w, ge, centroid, orientation, axis, s24, fallback, st, after_stream, rows);
    if (wpr <= 4) return launch_features<4>(cleaned, mask, n, h, w, ge, centroid, orientation, axis, s24, fallback, st, after_stream, rows);
    if (wpr <= 8) return launch_features<8>(cleaned, mask, n, h, w, ge, centroid, orientation, axis, s24, fallback, st, after_stream, rows);
    if (wpr <= 16) return launch_features<16>(cleaned, mask, n, h, w, ge, centroid, orientation, axis, s24, fallback, st, after_stream, rows);
    return launch_features<32>(cleaned, mask, n, h, w, ge, centroid, orientation, axis, s24, fallback, st, after_stream, rows);
}

}  // namespace msq

// scratch = the list of frames the streaming fast path hands to the general kernel: a count + up to n frame indices
extern "C" size_t msq_frame_features_scratch_bytes(int n, int, int) { return n > 0 ? ((size_t)n + 1) * sizeof(int) : 0; }

extern "C" int msq_frame_features(const uint8_t *cleaned, const uint8_t *mask, int n, int h, int w,
                                  double frame_threshold, double *centroid, double *orientation, double *axis,
                                  int64_t *sums24, void *scratch, size_t scratch_bytes, void *stream) {
    MSQ_REQUIRE(cleaned && mask && centroid && orientation && axis, MSQ_EINVAL, "msq_frame_features: null pointer");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0, MSQ_EINVAL, "msq_frame_features: bad sizes n=%d h=%d w=%d", n, h, w);
    if (n == 0) return MSQ_OK;
    // without (enough, 4-byte aligned) scratch every frame goes through the general kernel
    int *fallback = (scratch && scratch_bytes >= msq_frame_features_scratch_bytes(n, h, w) && (uintptr_t)scratch % 4 == 0)
                        ? static_cast<int *>(scratch) : nullptr;
    return msq::launch_frame_features(cleaned, mask, n, h, w, frame_threshold, centroid, orientation, axis, sums24, fallback,
                                      (cudaStream_t)stream);
}
