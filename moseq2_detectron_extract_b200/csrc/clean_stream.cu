// a6 clean_frames, streaming formulation (the fast path; clean.cu keeps the tiled kernel for odd widths).
// ref proc/proc.py:480-515: 3x3 median (replicate border) + one opening by the 9x9 ellipse.
//
// One WARP owns a full-width strip of a frame (32 lanes x 8 pixels = 256 plane columns: 240 output columns plus
// an 8-pixel halo lane on either side) and walks down its rows.  Per input row, in registers:
//   raw row  --3x3 median-->  M[m]  --row min 7/9-->  H7[m], H9[m]
//   V7[m-1] = min(H7[m-1], H7[m])            T3[m-1] = min(H9[m-2], H9[m-1], H9[m])
//   A[a]  = min(M[a-4], V7[a-3], T3[a])                          a = m-1     (the upper rows of the ellipse, as soon as T3 exists)
//   E[e]  = min(A[e], V7[e+2], M[e+4])                           e = m-4     (the 9x9 ellipse is rows of width 1,7,7,9,9,9,7,7,1)
//   ... and the same chain with max on E for the dilation, d = e-4.
// Horizontal neighbours come from warp shuffles (lane +-1); vertical history is a set of per-lane delay lines in
// shared memory (each lane only ever touches its own 16-byte slots, so there is no synchronisation at all), and the
// pair/triple pre-combination (V7, T3, A) means every delay line is written once and read once per row; 12 rows of delay
// per pass (6 + 3 + 3; an earlier version delayed M by 8, V7 by 6 and T3 by 3 = 17 rows) keep 12 KB per warp, so that
// shared memory allows 16 warps per SM, and with 6-row trips every slot index is a literal.
// ncu on the tiled kernel showed 52 M shared-memory wavefronts / 1000 frames (LSU pipe 68 % busy) next to a 53 %
// busy ALU pipe; this formulation needs ~3x fewer wavefronts, no block barriers and no second pass over the raw tile.
//
// That pipeline only ever runs where the result can be non-zero.  With scratch memory (msq_clean_frames_ws, msq_extract_chunk)
// a clean step is three launches:
//   clean_band_kernel   a 1-bit row scan per (frame, 240-column tile): the exact rows [lo, hi] on which the erosion is non-zero
//                       (from prep's positive-pixel bit rows when the caller has them, else thresholded from the frame); the
//                       same CTA stores the zero rows of the output, everything outside [lo - 4, hi + 4];
//   clean_plan_kernel   prefix sums of (band rows + pipeline lead-in) over the (frame, tile) columns;
//   clean_stream_kernel the pipeline above on equal-cost shares of those rows, one share per resident CTA.
// Without scratch (msq_clean_frames) one launch does all of it, strip by strip, at the price of a load-imbalanced tail.
#include "common.cuh"
#include <algorithm>
#include <limits.h>
#include <stdlib.h>
#include <type_traits>

namespace msq {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kOutCols = 240;              // output columns per warp-row (lanes 1..30)
constexpr int kDelayM = 6;                 // rows in the M delay line   (M[m-5] is read from the slot after the one M[m] is written to)
constexpr int kDelayV = 3;                 // rows in the V7 delay line  (V7[m-4] is read, then V7[m-1] overwrites it)
constexpr int kDelayA = 3;                 // rows in the A delay line   (A[m-4] is read, then A[m-1] overwrites it)
constexpr int kPrefetch = 3;               // raw rows requested ahead of use (divides kUnroll: the queue rotates by renaming)
constexpr int kUnroll = 6;                 // lcm of the delay-line / history periods (6, 3, 2): no register moves, static slots
constexpr int kWarpsPerCta = 1;

struct MinOp {
    static __device__ __forceinline__ uint32_t op3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
    static __device__ __forceinline__ uint32_t op2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
};
struct MaxOp {
    static __device__ __forceinline__ uint32_t op3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
    static __device__ __forceinline__ uint32_t op2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
};

// (p[k],p[k+1]),(p[k+2],p[k+3]) -> (p[k+1],p[k+2])   (moving this to the FMA pipe as IMAD.HI + IMAD was measured: no gain)
__device__ __forceinline__ uint32_t mid_pair(uint32_t a, uint32_t b) { return __funnelshift_r(a, b, 16); }
// med3(a,b,c) = max(min(a,b), min(max(a,b),c)) is 4 ALU ops; the same median as a sum, a+b+c-min-max:  Lanes hold values <= 255, so neither the 16-bit sums nor the differences carry
// across lanes.  `one`/`neg1` are kernel parameters (1, -1) so that the adds are IMADs on the otherwise idle FMA pipe:
// measured (tools/microbench/pipe_mix): in mixed code every ALU-pipe op (2- or 3-input VIMNMX, SHF, IADD3) costs
// 0.5 cycle/SM, IMAD runs beside them.
__device__ __forceinline__ uint32_t med3_of_sorted(uint32_t a, uint32_t b, uint32_t c, uint32_t lo, uint32_t hi, uint32_t one,
                                                   uint32_t neg1) {
    uint32_t s = a * one + b;
    s = c * one + s;
    s = lo * neg1 + s;
    return hi * neg1 + s;
}
__device__ __forceinline__ uint32_t med3_sum(uint32_t a, uint32_t b, uint32_t c, uint32_t one, uint32_t neg1) {
    return med3_of_sorted(a, b, c, __vimin3_u16x2(a, b, c), __vimax3_u16x2(a, b, c), one, neg1);
}
__device__ __forceinline__ uint4 splat(uint32_t v) { return make_uint4(v, v, v, v); }

struct RawRow {            // 8 pixels of one row as 16-bit lanes + the pixel just left / right of them
    uint4 c;
    uint32_t left;         // high lane = pixel -1
    uint32_t right;        // low lane  = pixel  8
};

// 7- and 9-wide row extrema of one row held in registers; neighbours by shuffle.  Lane 0 / 31 receive their own
// values as "neighbours": that only reaches the outer 4 columns of the halo lanes, which nothing consumes.
template <class OP>
__device__ __forceinline__ void row_extrema(const uint4 &b, uint4 &h7, uint4 &h9) {
    const uint32_t ax = __shfl_up_sync(kFull, b.z, 1), ay = __shfl_up_sync(kFull, b.w, 1);      // px -4..-1
    const uint32_t cx = __shfl_down_sync(kFull, b.x, 1), cy = __shfl_down_sync(kFull, b.y, 1);  // px  8..11
    const uint32_t s_a = mid_pair(ax, ay), s_ab = mid_pair(ay, b.x), s_b0 = mid_pair(b.x, b.y), s_b1 = mid_pair(b.y, b.z),
                   s_b2 = mid_pair(b.z, b.w), s_bc = mid_pair(b.w, cx), s_c = mid_pair(cx, cy);
    const uint32_t t1 = OP::op3(s_ab, b.x, s_b0), t2 = OP::op3(b.y, s_b1, b.z), t3 = OP::op3(s_b2, b.w, s_bc);
    h7.x = OP::op3(OP::op3(t1, s_a, ay), b.y, s_b1);
    h7.y = OP::op3(t1, t2, s_b2);
    h7.z = OP::op3(s_b0, t2, t3);
    h7.w = OP::op3(OP::op3(s_b1, b.z, t3), cx, s_c);
    h9.x = OP::op3(h7.x, ax, b.z);
    h9.y = OP::op3(h7.y, ay, b.w);
    h9.z = OP::op3(h7.z, b.x, cx);
    h9.w = OP::op3(h7.w, b.y, cy);
}

template <class OP> __device__ __forceinline__ uint4 op2_4(const uint4 &a, const uint4 &b) {
    return make_uint4(OP::op2(a.x, b.x), OP::op2(a.y, b.y), OP::op2(a.z, b.z), OP::op2(a.w, b.w));
}
template <class OP> __device__ __forceinline__ uint4 op3_4(const uint4 &a, const uint4 &b, const uint4 &c) {
    return make_uint4(OP::op3(a.x, b.x, c.x), OP::op3(a.y, b.y, c.y), OP::op3(a.z, b.z, c.z), OP::op3(a.w, b.w, c.w));
}

// ---- pass 1: which rows of a strip can hold a non-zero output at all? -----------------------------------------------------
// The opening's erosion E(x, e) is the minimum of the median image M over the 9x9 ellipse around (x, e) -- rows of width
// 1,7,7,9,9,9,7,7,1; pixels outside the image take no part in it.  So E(x, e) > 0 exactly when M > 0 on every in-image pixel
// of that ellipse, and M(p) > 0 exactly when at least 5 of the 9 raw pixels around p (replicate border) are positive.  Both
// statements are about ONE BIT per pixel: a lane turns 32 pixels of a row into a 32-bit mask, the 3x3 counts are bit-sliced
// adders (a dozen LOP3 for 32 pixels), horizontal runs are ANDs of funnel-shifted masks, the ellipse an AND over nine rows.
// The warp's four 8-lane groups each walk a quarter of the strip's rows with their vertical history in registers:
// ~30 instructions per 256-pixel row against ~230 for the full median + erosion + dilation.  On a prepared depth frame the
// floor is noise (about half of the pixels positive) and only the rows under the animal come out positive; every other row
// of the result is zero and is never sent through the full pipeline.  Exact, not a heuristic: the one uncertainty (pixels
// beyond the ends of the warp-row on images wider than one tile) is resolved as "positive".
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t x, int shift) {       // 4 bytes -> 4 bits placed at bit `shift`
    const uint32_t m = (bytes_nonzero(x) >> 7) * 0x00204081u;                      // bits 21..24
    return shift >= 21 ? (m << (shift - 21)) & (0xfu << shift) : (m >> (21 - shift)) & (0xfu << shift);
}

// The scan's view of one raw row: 32 "pixel is positive" bits per lane, taken either from the u8 frame itself (BytesRow: 32
// bytes -> 32 bits, ~60 ALU instructions per lane and row) or from the bit rows msq_prep_frames can write beside the frame
// (BitsRow: two aligned words and a funnel shift).  Both replicate the image border like the median filter does.
struct BytesRow {
    const uint8_t *src;
    int h, w, xb;
    struct Raw { uint2 v[4]; };
    __device__ __forceinline__ void init(const void *frame, int h_, int w_, int xb_) { src = static_cast<const uint8_t *>(frame); h = h_; w = w_; xb = xb_; }
    __device__ __forceinline__ Raw load_row(int y) const {                         // raw row y (clamped), this lane's 32 pixels
        const uint8_t *row = src + (size_t)min(max(y, 0), h - 1) * w;
        Raw r;
#pragma unroll
        for (int k = 0; k < 4; ++k) r.v[k] = __ldg(reinterpret_cast<const uint2 *>(row + min(max(xb + 8 * k, 0), w - 8)));
        return r;
    }
    __device__ __forceinline__ uint32_t positive(const Raw &r) const {             // bit k = pixel xb + k > 0
        uint32_t bits = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xg = xb + 8 * k;
            uint32_t b8 = (nonzero_nibble(r.v[k].x, 0) | nonzero_nibble(r.v[k].y, 4));
            if (xg < 0) b8 = (b8 & 1u) ? 0xffu : 0u;                               // replicate border
            else if (xg >= w) b8 = (b8 & 0x80u) ? 0xffu : 0u;
            bits |= b8 << (8 * k);
        }
        return bits;
    }
};

struct BitsRow {
    const uint32_t *src;       // (h, wpr) words of this frame: bit b of word i = pixel 32 i + b is positive
    int h, wpr, i_lo, i_hi, sh, last_bit;
    uint32_t lo_all, hi_all;   // words that lie wholly outside the image: 1 = left of it, 2 = right of it
    uint32_t tail;             // valid bits of the last word of a row
    struct Raw { uint32_t lo, hi; };
    __device__ __forceinline__ void init(const void *frame_bits, int h_, int w, int xb) {
        src = static_cast<const uint32_t *>(frame_bits); h = h_; wpr = (w + 31) >> 5;
        const int wi = xb >> 5;                                                    // arithmetic: floor for xb < 0
        sh = xb & 31;
        lo_all = wi < 0 ? 1u : (wi >= wpr ? 2u : 0u);
        hi_all = wi + 1 < 0 ? 1u : (wi + 1 >= wpr ? 2u : 0u);
        i_lo = min(max(wi, 0), wpr - 1); i_hi = min(max(wi + 1, 0), wpr - 1);
        last_bit = (w - 1) & 31;
        tail = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;
    }
    __device__ __forceinline__ Raw load_row(int y) const {
        const uint32_t *row = src + (size_t)min(max(y, 0), h - 1) * wpr;
        Raw r;
        r.lo = __ldg(row + i_lo); r.hi = __ldg(row + i_hi);
        return r;
    }
    __device__ __forceinline__ uint32_t fix(uint32_t v, uint32_t all, int idx) const {
        if (all == 1u) return (v & 1u) ? 0xffffffffu : 0u;                         // left of the image: pixel 0 replicated
        const uint32_t last = ((v >> last_bit) & 1u) ? 0xffffffffu : 0u;          // (idx == wpr - 1 whenever this is used)
        if (all == 2u) return last;
        return idx == wpr - 1 ? (v & tail) | (last & ~tail) : v;
    }
    __device__ __forceinline__ uint32_t positive(const Raw &r) const {
        return __funnelshift_r(fix(r.lo, lo_all, i_lo), fix(r.hi, hi_all, i_hi), sh);
    }
};

template <class Row>
struct BandScan {
    Row row;
    int w, xb, j;                    // image width, first column of this lane's 32-pixel block, block index 0..7 in the row
    uint32_t out_mask;               // bits of the block that lie outside the image

    // per pixel the number of positive pixels among (left, self, right) as two bit planes
    __device__ __forceinline__ void row_counts(const typename Row::Raw &r, uint32_t &s0, uint32_t &s1) const {
        const uint32_t P = row.positive(r);
        const uint32_t pl = __shfl_up_sync(kFull, P, 1, 8), pr = __shfl_down_sync(kFull, P, 1, 8);
        // beyond either end of the warp-row: the image border replicates the lane's own edge pixel, anything else is unknown (1)
        const uint32_t left = j == 0 ? ((xb - 1 < 0) ? ((P & 1u) << 31) : 0x80000000u) : pl;
        const uint32_t right = j == 7 ? ((xb + 32 >= w) ? (P >> 31) : 1u) : pr;
        const uint32_t L = __funnelshift_l(left, P, 1), R = __funnelshift_r(P, right, 1);
        s0 = L ^ P ^ R;
        s1 = (L & P) | (L & R) | (P & R);
    }
};

// rows e in [e_first, e_last] (in-image) on which the erosion is non-zero somewhere in this warp-row: [lo, hi] or hi < lo
template <class Row>
__device__ __forceinline__ void scan_band(const void *src, int h, int w, int x_span0, int e_first, int e_last, int lane,
                                          int &band_lo, int &band_hi) {
    BandScan<Row> sc;
    sc.w = w; sc.j = lane & 7; sc.xb = x_span0 + 32 * sc.j;
    sc.row.init(src, h, w, sc.xb);
    sc.out_mask = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((unsigned)(sc.xb + 8 * k) >= (unsigned)w) sc.out_mask |= 0xffu << (8 * k);
    const int rows = e_last - e_first + 1, per = (rows + 3) >> 2, grp = lane >> 3;
    const int q0 = e_first + grp * per, q1 = min(e_last, q0 + per - 1);            // centre rows of this 8-lane group
    int lo = INT_MAX, hi = -1;
    // median row m needs raw rows m-1..m+1; centre e needs median rows e-4..e+4: raw rows q0-5 .. q1+5
    uint32_t a0, a1, b0, b1;
    sc.row_counts(sc.row.load_row(q0 - 5), a0, a1);
    sc.row_counts(sc.row.load_row(q0 - 4), b0, b1);
    typename Row::Raw p0 = sc.row.load_row(q0 - 3), p1 = sc.row.load_row(q0 - 2);              // raw rows are requested two steps ahead
    // acc[i]: the ellipse AND of the centre row that completes i rows from now (acc0 completes with the current median row)
    uint32_t acc0 = ~0u, acc1 = ~0u, acc2 = ~0u, acc3 = ~0u, acc4 = ~0u, acc5 = ~0u, acc6 = ~0u, acc7 = ~0u, acc8 = ~0u;
    const int steps = per + 8;                                                      // the same for every group (warp-uniform loop)
    for (int i = 0; i < steps; ++i) {
        const int m = q0 - 4 + i;                                                   // median row of this step
        uint32_t c0, c1;
        const typename Row::Raw p2 = sc.row.load_row(m + 3);
        sc.row_counts(p0, c0, c1);                                                  // raw row m + 1
        p0 = p1; p1 = p2;
        // total = t0 + 2 u0 + 4 v0 + 8 v1 of the three rows' 2-bit counts; median positive <=> total >= 5
        const uint32_t t0 = a0 ^ b0 ^ c0, k0 = (a0 & b0) | (a0 & c0) | (b0 & c0);
        const uint32_t t1 = a1 ^ b1 ^ c1, k1 = (a1 & b1) | (a1 & c1) | (b1 & c1);
        const uint32_t u0 = k0 ^ t1, u1 = k0 & t1, v0 = u1 ^ k1, v1 = u1 & k1;
        uint32_t M = v1 | (v0 & (u0 | t0));
        M |= sc.out_mask;                                                           // outside the image: not part of the minimum
        if ((unsigned)m >= (unsigned)h) M = ~0u;
        const uint32_t ml = __shfl_up_sync(kFull, M, 1, 8), mr = __shfl_down_sync(kFull, M, 1, 8);
        const uint32_t ML = sc.j == 0 ? ~0u : ml, MR = sc.j == 7 ? ~0u : mr;
        const uint32_t c3 = M & __funnelshift_r(M, MR, 1) & __funnelshift_l(ML, M, 1);
        const uint32_t c5 = c3 & __funnelshift_r(M, MR, 2) & __funnelshift_l(ML, M, 2);
        const uint32_t c7 = c5 & __funnelshift_r(M, MR, 3) & __funnelshift_l(ML, M, 3);
        const uint32_t c9 = c7 & __funnelshift_r(M, MR, 4) & __funnelshift_l(ML, M, 4);
        // row m is row e+4 of centre e = m-4 (acc0), e+3 of m-3 (acc1), ... e-4 of m+4 (acc8): widths 1,7,7,9,9,9,7,7,1
        acc0 &= M; acc1 &= c7; acc2 &= c7; acc3 &= c9; acc4 &= c9; acc5 &= c9; acc6 &= c7; acc7 &= c7; acc8 &= M;
        const int e = m - 4;
        if ((acc0 & ~sc.out_mask) != 0u && e >= q0 && e <= q1) { lo = min(lo, e); hi = max(hi, e); }
        acc0 = acc1; acc1 = acc2; acc2 = acc3; acc3 = acc4; acc4 = acc5; acc5 = acc6; acc6 = acc7; acc7 = acc8; acc8 = ~0u;
        a0 = b0; a1 = b1; b0 = c0; b1 = c1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(kFull, lo, o));
        hi = max(hi, __shfl_xor_sync(kFull, hi, o));
    }
    band_lo = lo; band_hi = hi;
}

struct StreamGeom {
    int h, w, tiles_x, rows_per_cta;
    uint32_t one, neg1;        // 1 and -1, opaque to the compiler (see med3_of_sorted)
};

// Pass 1 as its own launch (msq_clean_frames_ws / msq_extract_chunk, which have scratch memory for its result): a CTA of two
// warps per (frame, 240-column tile), one warp per half of the height, finds the band [lo, hi] of rows with a non-zero erosion
// and writes it; the output rows outside [lo - 4, hi + 4] are zero, and the same CTA stores those zeros -- so the pipeline
// kernel below only ever sees rows that need the full treatment.  ~70 registers, 64 threads: 14 CTAs per SM hide the load
// latency that the same scan suffers inside the 128-register pipeline kernel (ncu: long-scoreboard stalls 2.3 per issue
// there), and 6000 frames are ~3 full waves of CTAs (one warp per frame was 1.45 waves: the second wave half empty).
constexpr int kBandWarps = 2;
constexpr int kLeadRows = 20;              // rows a strip spends filling the pipeline before its first output row
template <class Row>      // BytesRow: src = the u8 frames; BitsRow: src = their positive-pixel bit rows (n, h, ceil(w / 32)) u32
__global__ void __launch_bounds__(32 * kBandWarps)
clean_band_kernel(const void *__restrict__ src, uint8_t *__restrict__ out, int n, int h, int w, int tiles_x, int2 *__restrict__ bands) {
    __shared__ int2 part[kBandWarps];
    const int item = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = item / tiles_x, tx = item - f * tiles_x;
    const int half = (h + kBandWarps - 1) / kBandWarps;
    const int e0 = warp * half, e1 = min(h, e0 + half) - 1;
    int lo = INT_MAX, hi = -1;
    const void *frame = std::is_same<Row, BitsRow>::value
                            ? static_cast<const void *>(static_cast<const uint32_t *>(src) + (size_t)f * h * ((w + 31) >> 5))
                            : static_cast<const void *>(static_cast<const uint8_t *>(src) + (size_t)f * h * w);
    if (e0 <= e1) scan_band<Row>(frame, h, w, tx * kOutCols - 8, e0, e1, lane, lo, hi);
    if (lane == 0) part[warp] = make_int2(lo, hi);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBandWarps; ++k) { lo = min(lo, part[k].x); hi = max(hi, part[k].y); }
    if (threadIdx.x == 0) bands[item] = make_int2(lo, hi);
    // zero rows: everything outside [lo - 4, hi + 4]
    const int act0 = hi < 0 ? h : max(0, lo - 4), act1 = hi < 0 ? h : min(h, hi + 5);
    const int x0 = tx * kOutCols, groups = (min(w, x0 + kOutCols) - x0) >> 3;          // 8-pixel groups of this tile's columns
    uint8_t *dst = out + (size_t)f * h * w + x0;
    if (tiles_x == 1 && (w & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        // one tile per row: the zero rows are two contiguous byte ranges (frame bases are 16-byte aligned when h * w is)
        uint4 *top = reinterpret_cast<uint4 *>(dst), *bot = reinterpret_cast<uint4 *>(dst + (size_t)act1 * w);
        const int n_top = (act0 * w) >> 4, n_bot = ((h - act1) * w) >> 4;
        for (int i = threadIdx.x; i < n_top; i += 32 * kBandWarps) top[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < n_bot; i += 32 * kBandWarps) bot[i] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    const int zero_rows = act0 + (h - act1);
    for (int i = threadIdx.x; i < zero_rows * 32; i += 32 * kBandWarps) {
        const int zr = i >> 5, g = i & 31;
        const int y = zr < act0 ? zr : act1 + (zr - act0);
        if (g < groups) *reinterpret_cast<uint2 *>(dst + (size_t)y * w + 8 * g) = make_uint2(0u, 0u);
    }
}

// prefix[i] = cost of the (frame, tile) columns before column i, cost = rows that go through the pipeline + its lead-in.
// One CTA; a few thousand items.
__global__ void __launch_bounds__(1024)
clean_plan_kernel(const int2 *__restrict__ bands, int items, int h, int *__restrict__ prefix) {
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // thread t owns the consecutive items [t * per, (t + 1) * per): one block-wide scan of the per-thread sums
    const int per = (items + 1023) >> 10, i0 = threadIdx.x * per, i1 = min(items, i0 + per);
    auto cost = [&](int i) {
        const int2 b = __ldg(bands + i);
        return b.y >= 0 ? min(h, b.y + 5) - max(0, b.x - 4) + kLeadRows : 0;
    };
    int mine = 0;
    for (int i = i0; i < i1; ++i) mine += cost(i);
    int v = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, v, o); if (lane >= o) v += t; }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int ws = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, ws, o); if (lane >= o) ws += t; }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    int run = v - mine + (warp > 0 ? warp_sums[warp - 1] : 0);         // cost of everything before item i0
    if (threadIdx.x == 0) prefix[0] = 0;
    for (int i = i0; i < i1; ++i) { run += cost(i); prefix[i + 1] = run; }
}

struct Rings {
    // per-lane delay lines: [row slot][lane] uint4
    uint4 m[kDelayM][32], v[kDelayV][32], a[kDelayA][32];
    uint4 e[kDelayM][32], w[kDelayV][32], b[kDelayA][32];
};

// rows [ys0, ys1) of tile tx of frame f through median -> erosion -> dilation (one warp)
__device__ __forceinline__ void run_strip(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, const StreamGeom &G, Rings &R,
                                          int f, int tx, int ys0, int ys1, int lane) {
    const int h = G.h, w = G.w;
    uint4 (*ring_m)[32] = R.m, (*ring_v)[32] = R.v, (*ring_a)[32] = R.a, (*ring_e)[32] = R.e, (*ring_w)[32] = R.w, (*ring_b)[32] = R.b;
    const int x_lane = tx * kOutCols - 8 + (lane << 3);            // image column of this lane's first pixel
    const bool col_in = (unsigned)x_lane < (unsigned)w;            // groups are 8-aligned: all in or all out
    const bool writes = lane >= 1 && lane <= 30 && col_in;
    const uint8_t *src = in + (size_t)f * h * w;
    uint8_t *dst = out + (size_t)f * h * w;
    // column addressing of the raw loads (replicate border): nearest in-image group + which byte to splat
    const int xg = min(max(x_lane, 0), w - 8);
    const int x_left = min(max(x_lane - 1, 0), w - 1), x_right = min(max(x_lane + 8, 0), w - 1);
    const bool need_edge = G.tiles_x > 1;

    auto issue = [&](int y, uint2 &v, uint32_t &edge) {           // raw bytes of input row y (clamped)
        const uint8_t *row = src + (size_t)min(max(y, 0), h - 1) * w;
        v = __ldg(reinterpret_cast<const uint2 *>(row + xg));
        // pixel -1 / pixel 8 come from the neighbouring lane except at the two ends of the warp-row
        // (only when the image is wider than one warp-row: otherwise lanes 0 / 31 lie outside the image and the
        //  pixel beyond them reaches nothing that is consumed -- a warp-uniform test, no divergence in the common case)
        edge = 0u;
        if (need_edge) edge = (lane == 0) ? (uint32_t)__ldg(row + x_left) : ((lane == 31) ? (uint32_t)__ldg(row + x_right) : 0u);
    };
    auto decode = [&](uint2 v, uint32_t edge) {
        if (x_lane < 0) { v.x = (v.x & 0xffu) * 0x01010101u; v.y = v.x; }
        else if (x_lane >= w) { v.y = (v.y >> 24) * 0x01010101u; v.x = v.y; }
        RawRow r;
        r.c = make_uint4(__byte_perm(v.x, 0, 0x4140), __byte_perm(v.x, 0, 0x4342), __byte_perm(v.y, 0, 0x4140),
                         __byte_perm(v.y, 0, 0x4342));
        const uint32_t from_left = __shfl_up_sync(kFull, r.c.w, 1), from_right = __shfl_down_sync(kFull, r.c.x, 1);
        r.left = (lane == 0) ? (edge << 16) : from_left;
        r.right = (lane == 31) ? edge : from_right;
        return r;
    };

    const int y_first = ys0 - 9;                                   // first input row of the strip
    // The three stages are software-pipelined across iterations: iteration s computes the median row s-1, the
    // erosion from the median row of iteration s-1 and the dilation from the erosion row of iteration s-2, so the
    // stages inside one iteration are mutually independent instruction streams (ILP for the ~10 resident warps).
    const int steps = (ys1 - ys0) + 20;
    const uint32_t one = G.one, neg1 = G.neg1;
    uint2 pf_v[kPrefetch];
    uint32_t pf_e[kPrefetch];
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) issue(y_first + k, pf_v[k], pf_e[k]);

    RawRow r0, r1, r2;                                             // the three newest raw rows (r2 newest)
    r0.c = r1.c = r2.c = splat(0u); r0.left = r1.left = r2.left = 0u; r0.right = r1.right = r2.right = 0u;
    uint4 h7_prev = splat(0u), h9_prev = splat(0u), h9_prev2 = splat(0u), v7_prev = splat(0u);
    uint4 g7_prev = splat(0u), g9_prev = splat(0u), g9_prev2 = splat(0u), w7_prev = splat(0u);
    uint4 M_cur = splat(0u), E_cur = splat(0u);                    // M[s-2] and E[s-7] entering iteration s

    // kUnroll iterations per trip: every rotation below (raw rows 3, prefetch queue 3, H9 history 3, V7 history 2, the
    // 6- and 3-slot delay lines) has a period dividing 6, so after unrolling the compiler renames instead of moving and
    // the V/T slot indices are literals.  Trips past `steps` only touch clamped rows and store nothing.
    for (int s0 = 0; s0 < steps; s0 += kUnroll) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int s = s0 + u;
            const int slot_w = u % kDelayM, slot_r = (u + 1) % kDelayM, slot3 = u % kDelayV;
            // ================= stage 3: dilation row d = e-4 from E_cur = E[e], e = s-7 =================
            {
                const int e = s - 7;
                uint4 g7, g9;
                row_extrema<MaxOp>(E_cur, g7, g9);
                const uint4 w7 = op2_4<MaxOp>(g7_prev, g7);                 // W7[e-1]
                const uint4 u3 = op3_4<MaxOp>(g9_prev2, g9_prev, g9);       // U3[e-1]
                const uint4 e_old = ring_e[slot_r][lane];                   // E[e-5]
                ring_e[slot_w][lane] = E_cur;
                const uint4 w_old = ring_w[slot3][lane];                    // W7[e-4], then W7[e-1] takes its slot
                ring_w[slot3][lane] = w7;
                const uint4 b_new = op3_4<MaxOp>(e_old, w_old, u3);         // B[e-1] = max(E[e-5], W7[e-4], U3[e-1])
                const uint4 b_old = ring_b[slot3][lane];                    // B[e-4], then B[e-1] takes its slot
                ring_b[slot3][lane] = b_new;
                const uint4 D = op3_4<MaxOp>(b_old, w7_prev, E_cur);        // D[e-4] = max(B[e-4], W7[e-2], E[e])
                g7_prev = g7; g9_prev2 = g9_prev; g9_prev = g9; w7_prev = w7;
                const int yd = y_first + e - 4;
                if (writes && yd >= ys0 && yd < ys1) {
                    const uint2 packed = make_uint2(__byte_perm(D.x, D.y, 0x6420), __byte_perm(D.z, D.w, 0x6420));
                    *reinterpret_cast<uint2 *>(dst + ((size_t)yd * w + x_lane)) = packed;
                }
            }
            // ================= stage 2: erosion row e = m-4 from M_cur = M[m], m = s-2 =================
            uint4 E_next;
            {
                const int m = s - 2;
                uint4 h7, h9;
                row_extrema<MinOp>(M_cur, h7, h9);
                const uint4 v7 = op2_4<MinOp>(h7_prev, h7);                 // V7[m-1]
                const uint4 t3 = op3_4<MinOp>(h9_prev2, h9_prev, h9);       // T3[m-1]
                const uint4 m_old = ring_m[slot_r][lane];                   // M[m-5]
                ring_m[slot_w][lane] = M_cur;
                const uint4 v_old = ring_v[slot3][lane];                    // V7[m-4], then V7[m-1] takes its slot
                ring_v[slot3][lane] = v7;
                const uint4 a_new = op3_4<MinOp>(m_old, v_old, t3);         // A[m-1] = min(M[m-5], V7[m-4], T3[m-1])
                const uint4 a_old = ring_a[slot3][lane];                    // A[m-4], then A[m-1] takes its slot
                ring_a[slot3][lane] = a_new;
                E_next = op3_4<MinOp>(a_old, v7_prev, M_cur);               // E[m-4] = min(A[m-4], V7[m-2], M[m])
                h7_prev = h7; h9_prev2 = h9_prev; h9_prev = h9; v7_prev = v7;
                const int ye = y_first + m - 4;
                if (!(col_in && (unsigned)ye < (unsigned)h)) E_next = splat(0u);               // dilation identity outside the image
            }
            // ================= stage 1: median row m = s-1 (image row y_first + s - 1), centre row r1 =================
            uint4 M_next;
            {
                r0 = r1; r1 = r2;
                r2 = decode(pf_v[0], pf_e[0]);                              // raw row y_first + s
#pragma unroll
                for (int k = 0; k + 1 < kPrefetch; ++k) { pf_v[k] = pf_v[k + 1]; pf_e[k] = pf_e[k + 1]; }
                issue(y_first + s + kPrefetch, pf_v[kPrefetch - 1], pf_e[kPrefetch - 1]);     // rows are clamped: always in bounds
                uint32_t lo[6], mi[6], hi[6];
                const uint32_t a0[6] = {r0.left, r0.c.x, r0.c.y, r0.c.z, r0.c.w, r0.right};
                const uint32_t a1[6] = {r1.left, r1.c.x, r1.c.y, r1.c.z, r1.c.w, r1.right};
                const uint32_t a2[6] = {r2.left, r2.c.x, r2.c.y, r2.c.z, r2.c.w, r2.right};
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    lo[q] = __vimin3_u16x2(a0[q], a1[q], a2[q]);
                    hi[q] = __vimax3_u16x2(a0[q], a1[q], a2[q]);
                    mi[q] = med3_of_sorted(a0[q], a1[q], a2[q], lo[q], hi[q], one, neg1);
                }
                uint32_t res[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t max_lo = __vimax3_u16x2(mid_pair(lo[q], lo[q + 1]), lo[q + 1], mid_pair(lo[q + 1], lo[q + 2]));
                    const uint32_t min_hi = __vimin3_u16x2(mid_pair(hi[q], hi[q + 1]), hi[q + 1], mid_pair(hi[q + 1], hi[q + 2]));
                    const uint32_t med_mi = med3_sum(mid_pair(mi[q], mi[q + 1]), mi[q + 1], mid_pair(mi[q + 1], mi[q + 2]), one, neg1);
                    res[q] = med3_sum(max_lo, med_mi, min_hi, one, neg1);
                }
                const int ym = y_first + s - 1;
                const bool in_img = col_in && (unsigned)ym < (unsigned)h;   // the erosion must ignore pixels outside the image
                M_next = in_img ? make_uint4(res[0], res[1], res[2], res[3]) : splat(0x00ff00ffu);
            }
            M_cur = M_next;
            E_cur = E_next;
        }
    }
    __syncwarp();
}

template <bool kBandsGiven>
__global__ void __launch_bounds__(32 * kWarpsPerCta, 16)
clean_stream_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int n, StreamGeom G, const int2 *__restrict__ bands,
                    const int *__restrict__ prefix) {
    __shared__ Rings R;
    const int lane = threadIdx.x & 31;
    const int h = G.h, w = G.w;
    if (kBandsGiven) {
        // Work is cut by COST: the launch is one long line of (lead-in + band rows) segments, one per (frame, tile) with a
        // non-empty band, and every CTA owns the same length of it -- the zero rows are already written (clean_band_kernel).
        const int items = n * G.tiles_x;
        const int total = prefix[items];
        const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
        long long p = (long long)blockIdx.x * per;
        const long long p_end = min((long long)total, p + per);
        if (p >= p_end) return;
        int j = 0, j_hi = items;                                       // last j with prefix[j] <= p: the segment that holds p
        while (j_hi - j > 1) {
            const int mid = (j + j_hi) >> 1;
            if (prefix[mid] <= p) j = mid; else j_hi = mid;
        }
        while (p < p_end) {
            const int base = prefix[j], next = prefix[j + 1];
            if (next == base) { ++j; continue; }
            const int2 b = bands[j];
            const int act0 = max(0, b.x - 4);
            const int a = (int)(p - base), e = (int)(min(p_end, (long long)next) - base);
            const int ys0 = act0 + max(0, a - kLeadRows), ys1 = act0 + max(0, e - kLeadRows);
            const int item = j;
            p = (long long)base + e;
            if (e == next - base) ++j;
            if (ys0 < ys1) run_strip(in, out, G, R, item / G.tiles_x, item % G.tiles_x, ys0, ys1, lane);
        }
        return;
    }
    // No scratch: work is cut by ROWS.  The launch is one tall stack of n * tiles_x columns of h rows each, and CTA c owns
    // rows [c * rows_per_cta, (c + 1) * rows_per_cta) of it, processed as one strip per column it touches (each strip scans
    // its own rows for the band, writes the zero rows and pays the lead-in for the rest).
    const long long total_rows = (long long)n * G.tiles_x * h;
    long long g0 = (long long)blockIdx.x * G.rows_per_cta;
    const long long g1 = min(total_rows, g0 + G.rows_per_cta);
    while (g0 < g1) {
        const long long col = g0 / h;                                  // (frame, tile) column of the stack
        const int f = (int)(col / G.tiles_x), tx = (int)(col - (long long)f * G.tiles_x);
        const int y_out0 = (int)(g0 - col * h), y_out1 = (int)min((long long)h, y_out0 + (g1 - g0));
        g0 += y_out1 - y_out0;
        const int x_lane = tx * kOutCols - 8 + (lane << 3);
        const bool writes = lane >= 1 && lane <= 30 && (unsigned)x_lane < (unsigned)w;
        uint8_t *dst = out + (size_t)f * h * w;
        int band_lo = INT_MAX, band_hi = -1;
        scan_band<BytesRow>(in + (size_t)f * h * w, h, w, tx * kOutCols - 8, max(0, y_out0 - 4), min(h, y_out1 + 4) - 1, lane, band_lo, band_hi);
        // rows outside [band_lo - 4, band_hi + 4] are zero
        const int act0 = band_hi < 0 ? y_out1 : min(y_out1, max(y_out0, band_lo - 4)), act1 = band_hi < 0 ? y_out1 : max(act0, min(y_out1, band_hi + 5));
        if (writes) {
            for (int y = y_out0; y < act0; ++y) *reinterpret_cast<uint2 *>(dst + ((size_t)y * w + x_lane)) = make_uint2(0u, 0u);
            for (int y = act1; y < y_out1; ++y) *reinterpret_cast<uint2 *>(dst + ((size_t)y * w + x_lane)) = make_uint2(0u, 0u);
        }
        if (act0 < act1) run_strip(in, out, G, R, f, tx, act0, act1, lane);
    }
}

}  // namespace

// returns MSQ_EUNSUPPORTED-like negative hint (-100) when the streaming kernel cannot serve the shape.
// bands: n * tiles_x int2 of scratch for the separate pre-pass launch, or nullptr (then every strip scans its own rows).
int launch_clean_stream(const uint8_t *in, uint8_t *out, int n, int h, int w, cudaStream_t st, int2 *bands, RowBands *written,
                        const uint32_t *positive_bits) {
    const bool vec = (w % 8 == 0) && w >= 8 && ((uintptr_t)in % 8 == 0) && ((uintptr_t)out % 8 == 0);
    if (!vec) return -100;
    StreamGeom G;
    G.h = h; G.w = w; G.one = 1u; G.neg1 = 0xffffffffu;
    G.tiles_x = (w + kOutCols - 1) / kOutCols;
    static thread_local int resident = 0;          // co-resident CTAs per SM (shared-memory limited, ~10)
    if (resident == 0) {
        int r = 0;
        MSQ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, clean_stream_kernel<false>, 32 * kWarpsPerCta, 0));
        resident = std::max(1, r);
    }
    // persistent grid: every CTA is resident and gets the same number of rows; at least ~120 rows each so that the
    // 20-row lead-in of a strip stays a small part of it
    const long long total_rows = (long long)n * G.tiles_x * h;
    long long ctas = std::min<long long>((long long)sm_count() * resident, std::max<long long>(1, total_rows / 120));
    G.rows_per_cta = (int)((total_rows + ctas - 1) / ctas);
    const int grid = (int)((total_rows + G.rows_per_cta - 1) / G.rows_per_cta);
    TimedLaunch timed(K_CLEAN, st, bands ? 3 : 1);
    if (bands) {
        const int items = n * G.tiles_x;
        int *prefix = reinterpret_cast<int *>(bands + items);
        if (positive_bits) clean_band_kernel<BitsRow><<<items, 32 * kBandWarps, 0, st>>>(positive_bits, out, n, h, w, G.tiles_x, bands);
        else clean_band_kernel<BytesRow><<<items, 32 * kBandWarps, 0, st>>>(in, out, n, h, w, G.tiles_x, bands);
        clean_plan_kernel<<<1, 1024, 0, st>>>(bands, items, h, prefix);
        const int sgrid = (int)std::min<long long>((long long)sm_count() * resident, std::max<long long>(1, (long long)items));
        clean_stream_kernel<true><<<sgrid, 32 * kWarpsPerCta, 0, st>>>(in, out, n, G, bands, prefix);
        if (written) *written = {bands, G.tiles_x};
    } else {
        clean_stream_kernel<false><<<grid, 32 * kWarpsPerCta, 0, st>>>(in, out, n, G, nullptr, nullptr);
    }
    MSQ_LAUNCH_OK("clean_frames (streaming)");
    return MSQ_OK;
}

}  // namespace msq
