// a2 prep_raw_frames and a3 scale_raw_frames as streaming HBM-bound kernels.
//
// prep: ref proc/proc.py:129-172.  out[n,r,c] = u8( clamp( (bg[y0+r,x0+c] - raw[n,y0+r,x0+c]) * roi ) ).
// Only the ROI bounding box of the raw frame is ever read (apply_roi crops everything else away,
// ref proc/roi.py:229-235), so the algorithmic traffic is 2 B in + 1 B out per box pixel.
// Layout: a thread owns 8 horizontally adjacent box pixels at a FIXED (row, col) and walks over
// frames, so the background and ROI values live in registers for the whole walk and every raw
// access is one 128-bit streaming load (8 x int16) / one 64-bit store.
#include "common.cuh"
#include <cuda_bf16.h>
#include <type_traits>
#include <algorithm>

namespace msq {
namespace {

constexpr int kPrepThreads = 256;
constexpr int kPrepUnroll = 4;

template <typename T> struct PrepMath;

// float32 background: NumPy computes float32 - int16 -> float32 and compares against the
// (weak) Python scalars in float32.
template <> struct PrepMath<float> {
    using acc_t = float;
    float lo, hi;
    __host__ PrepMath(double vmin, double vmax) : lo((float)vmin), hi((float)vmax) {}
    __device__ __forceinline__ uint32_t run(float bg, int raw, float roi, int flags) const {
        float d = __fmul_rn(__fsub_rn(bg, (float)raw), roi);
        if ((flags & MSQ_PREP_HAS_VMIN) && d < lo) d = 0.0f;
        if ((flags & MSQ_PREP_HAS_VMAX) && d > hi) d = hi;
        return (uint32_t)(int)d & 0xffu;
    }
};
template <> struct PrepMath<double> {
    using acc_t = double;
    double lo, hi;
    __host__ PrepMath(double vmin, double vmax) : lo(vmin), hi(vmax) {}
    __device__ __forceinline__ uint32_t run(double bg, int raw, double roi, int flags) const {
        double d = __dmul_rn(__dsub_rn(bg, (double)raw), roi);
        if ((flags & MSQ_PREP_HAS_VMIN) && d < lo) d = 0.0;
        if ((flags & MSQ_PREP_HAS_VMAX) && d > hi) d = hi;
        return (uint32_t)(int)d & 0xffu;
    }
};
// uint16 background (TIFF cache) or no background: integer arithmetic in int32.
template <> struct PrepMath<int> {
    using acc_t = int;
    double lo, hi;
    int hi_i;
    __host__ PrepMath(double vmin, double vmax) : lo(vmin), hi(vmax), hi_i((int)vmax) {}
    __device__ __forceinline__ uint32_t run(int bg_minus_raw, int /*raw*/, int roi, int flags) const {
        int d = bg_minus_raw * roi;
        if ((flags & MSQ_PREP_HAS_VMIN) && (double)d < lo) d = 0;
        if ((flags & MSQ_PREP_HAS_VMAX) && (double)d > hi) d = hi_i;
        return (uint32_t)d & 0xffu;
    }
};

__device__ __forceinline__ int s16_lo(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int s16_hi(uint32_t v) { return (int)(short)(v >> 16); }

// Positive-pixel bit rows (optional second output, read by the cleaning pass's row scan instead of the frame itself):
// per frame and row ceil(w / 32) little-endian words, bit b of word i = prepared pixel 32 i + b is > 0.
__device__ __host__ __forceinline__ int positive_row_bytes(int w) { return 4 * ((w + 31) >> 5); }
__device__ __forceinline__ uint32_t nonzero_bytes_to_nibble(uint32_t x) {     // bit k = byte k of x is non-zero
    return (((bytes_nonzero(x) >> 7) * 0x00204081u) >> 21) & 0xfu;
}

// bit k = byte k of lo is non-zero, bit 4 + k = byte k of hi is non-zero.  The flags (bit 7 of every byte) of hi and, moved down
// by 4, of lo share one word; one multiplication by 2^21 + 2^14 + 2^7 + 1 then lines all eight up in the top byte (no two
// partial products meet below bit 24, so nothing carries into it).
__device__ __forceinline__ uint8_t nonzero_bytes_to_byte(uint32_t lo, uint32_t hi) {
    const uint32_t flags = (bytes_nonzero(lo) >> 4) | bytes_nonzero(hi);
    return (uint8_t)((flags * 0x00204081u) >> 24);
}

// BG: element type of the background in global memory; ACC: arithmetic type (float/double/int)
template <typename BG, typename ACC, bool HAS_BG>
__global__ void __launch_bounds__(kPrepThreads)
prep_vec8_kernel(const int16_t *__restrict__ frames, int n, int H, int W,
                 const BG *__restrict__ bground, const uint8_t *__restrict__ roi,
                 int y0, int x0, int h, int w, PrepMath<ACC> math, int flags, int frames_per_group,
                 uint8_t *__restrict__ out, int32_t *__restrict__ invalid_count, uint8_t *__restrict__ invalid_bits,
                 uint8_t *__restrict__ positive_bits) {
    const int w8 = w >> 3;
    const int pos = blockIdx.x * kPrepThreads + threadIdx.x;
    if (pos >= h * w8) return;
    const int r = pos / w8;
    const int c = (pos - r * w8) << 3;
    const size_t in_off = (size_t)(y0 + r) * W + (x0 + c);
    const size_t out_off = (size_t)r * w + c;
    const int pbr = positive_row_bytes(w);

    ACC bg[8], rm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        bg[k] = HAS_BG ? (ACC)bground[in_off + k] : (ACC)0;
        rm[k] = roi ? (ACC)(roi[in_off + k] != 0) : (ACC)1;
    }
    const size_t in_stride = (size_t)H * W;
    const size_t out_stride = (size_t)h * w;
    const int f_begin = blockIdx.y * frames_per_group;
    const int f_end = min(n, f_begin + frames_per_group);

    for (int f = f_begin; f < f_end; f += kPrepUnroll) {
        uint4 raw[kPrepUnroll];
#pragma unroll
        for (int u = 0; u < kPrepUnroll; ++u)
            if (f + u < f_end) raw[u] = ldg_stream_u4(frames + (size_t)(f + u) * in_stride + in_off);
#pragma unroll
        for (int u = 0; u < kPrepUnroll; ++u) {
            if (f + u >= f_end) break;
            const uint32_t words[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
            uint32_t packed[2] = {0u, 0u};
            int bad = 0;
            uint32_t bad_bits = 0u, pos_bits = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int v = (k & 1) ? s16_hi(words[k >> 1]) : s16_lo(words[k >> 1]);
                const bool is_bad = (v == 0) && (rm[k] != (ACC)0);
                bad += is_bad;
                bad_bits |= (uint32_t)is_bad << k;
                uint32_t o;
                if constexpr (std::is_same<ACC, int>::value)
                    o = math.run(HAS_BG ? (int)bg[k] - v : v, v, (int)rm[k], flags);
                else
                    o = math.run(bg[k], v, rm[k], flags);
                packed[k >> 2] |= o << ((k & 3) * 8);
                pos_bits |= (uint32_t)(o != 0u) << k;
            }
            stg_stream_u2(out + (size_t)(f + u) * out_stride + out_off, make_uint2(packed[0], packed[1]));
            if (invalid_count && bad) atomicAdd(invalid_count + f + u, bad);
            if (invalid_bits) invalid_bits[((size_t)(f + u) * h + r) * w8 + (c >> 3)] = (uint8_t)bad_bits;
            if (positive_bits) positive_bits[((size_t)(f + u) * h + r) * pbr + (c >> 3)] = (uint8_t)pos_bits;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// float32 background, vmin == 0, 0 <= vmax <= 255 (the reference's configuration: min_height 0, max_height 100).
// ncu on the generic kernel above: 21 instructions per pixel, ALU pipe 69 % and the conversion unit (I2F/F2I) 37 % busy
// next to 39 % DRAM -- it is instruction-bound, not bandwidth-bound.  Same arithmetic, fewer instructions:
//   * int16 -> float without the conversion unit: the bytes of (raw ^ 0x8000) dropped into the mantissa of 2^23 give
//     2^23 + raw + 32768 exactly (one PRMT), one exact FADD removes the bias;
//   * d = bg - raw in float32 as before;
//   * truncation without F2I: adding 2^23 with round-toward-zero leaves floor(d) in the low mantissa bits;
//   * "d < 0 -> 0, d > vmax -> vmax, * roi" on that integer is relu(min(., roi ? trunc(vmax) : 0)), one instruction
//     (roi == 0 gives +-0 in the reference, which its clamps and the cast turn into 0 as well); a PRMT per two pixels
//     packs the bytes;
//   * invalid pixels (raw == 0 inside the ROI) are found with a has-zero-halfword test on the packed words; the per-pixel
//     bookkeeping only runs for the (rare) 8-pixel groups that contain one.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t zero_halfwords(uint32_t w) {        // 0x8000 flag per 16-bit half that is zero
    return (w - 0x00010001u) & ~w & 0x80008000u;
}

__global__ void __launch_bounds__(kPrepThreads)
prep_vec8_f32_fast_kernel(const int16_t *__restrict__ frames, int n, int H, int W, const float *__restrict__ bground,
                          const uint8_t *__restrict__ roi, int y0, int x0, int h, int w, float hi, int frames_per_group,
                          uint8_t *__restrict__ out, int32_t *__restrict__ invalid_count, uint8_t *__restrict__ invalid_bits,
                          uint8_t *__restrict__ positive_bits) {
    const int w8 = w >> 3;
    const int pos = blockIdx.x * kPrepThreads + threadIdx.x;
    if (pos >= h * w8) return;
    const int r = pos / w8;
    const int c = (pos - r * w8) << 3;
    const size_t in_off = (size_t)(y0 + r) * W + (x0 + c);
    const size_t out_off = (size_t)r * w + c;
    const int pbr = positive_row_bytes(w);
    uint8_t *pos_base = positive_bits ? positive_bits + (size_t)r * pbr + (c >> 3) : nullptr;
    const size_t pos_stride = (size_t)h * pbr;
    const bool last_group = (c >> 3) == w8 - 1;
    float bg[8];
    int top[8];                                               // trunc(vmax) inside the ROI, 0 outside
    uint32_t roi_bits = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        bg[k] = bground[in_off + k];
        const bool in_roi = roi ? roi[in_off + k] != 0 : true;
        top[k] = in_roi ? (int)hi : 0;
        roi_bits |= (uint32_t)in_roi << k;
    }
    const size_t in_stride = (size_t)H * W, out_stride = (size_t)h * w;
    const int f_begin = blockIdx.y * frames_per_group;
    const int f_end = min(n, f_begin + frames_per_group);
    const uint32_t two23 = 0x4B000000u;                      // 2^23 as float bits
    const float bias = 8421376.0f;                            // 2^23 + 32768

    for (int f = f_begin; f < f_end; f += kPrepUnroll) {
        uint4 raw[kPrepUnroll];
#pragma unroll
        for (int u = 0; u < kPrepUnroll; ++u)
            if (f + u < f_end) raw[u] = ldg_stream_u4(frames + (size_t)(f + u) * in_stride + in_off);
#pragma unroll
        for (int u = 0; u < kPrepUnroll; ++u) {
            if (f + u >= f_end) break;
            const uint32_t words[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
            uint32_t t[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t wf = words[q] ^ 0x80008000u;
                const float r0 = __fadd_rn(__uint_as_float(__byte_perm(wf, two23, 0x7610)), -bias);
                const float r1 = __fadd_rn(__uint_as_float(__byte_perm(wf, two23, 0x7632)), -bias);
                // bits(2^23 + d, rounded toward zero) - bits(2^23) = floor(d) for |d| < 2^23, and positive floats order like
                // integers beyond that; one VIADDMNMX.RELU then does "d < 0 -> 0, d > vmax -> vmax, outside the ROI -> 0"
                const int i0 = __float_as_int(__fadd_rz(__fsub_rn(bg[2 * q], r0), 8388608.0f));
                const int i1 = __float_as_int(__fadd_rz(__fsub_rn(bg[2 * q + 1], r1), 8388608.0f));
                t[2 * q] = (uint32_t)__viaddmin_s32_relu(i0, -0x4B000000, top[2 * q]);
                t[2 * q + 1] = (uint32_t)__viaddmin_s32_relu(i1, -0x4B000000, top[2 * q + 1]);
            }
            const uint32_t p01 = __byte_perm(t[0], t[1], 0x0040), p23 = __byte_perm(t[2], t[3], 0x0040);
            const uint32_t p45 = __byte_perm(t[4], t[5], 0x0040), p67 = __byte_perm(t[6], t[7], 0x0040);
            const uint2 o8 = make_uint2(__byte_perm(p01, p23, 0x5410), __byte_perm(p45, p67, 0x5410));
            stg_stream_u2(out + (size_t)(f + u) * out_stride + out_off, o8);
            if (positive_bits) {
                // (measured: the byte store, not the bit gathering, is what this output costs -- ~40 us per 6000 frames; wider
                // stores through shuffles do not change it.  The last group of a row also writes the row's pad bytes, so that
                // every 32-byte sector ends up fully written.)
                uint8_t *dstb = pos_base + (size_t)(f + u) * pos_stride;
                stg_stream_u8(dstb, nonzero_bytes_to_byte(o8.x, o8.y));
                if (last_group)
                    for (int q = 1; q <= pbr - w8; ++q) dstb[q] = 0;
            }
            const uint32_t z0 = zero_halfwords(words[0]), z1 = zero_halfwords(words[1]), z2 = zero_halfwords(words[2]),
                           z3 = zero_halfwords(words[3]);
            uint32_t bad_bits = 0u;
            if ((z0 | z1 | z2 | z3) != 0u) {
                const uint32_t zs[4] = {z0, z1, z2, z3};
#pragma unroll
                for (int q = 0; q < 4; ++q) bad_bits |= (((zs[q] >> 15) & 1u) | ((zs[q] >> 30) & 2u)) << (2 * q);
                bad_bits &= roi_bits;
                if (invalid_count && bad_bits) atomicAdd(invalid_count + f + u, __popc(bad_bits));
            }
            if (invalid_bits) invalid_bits[((size_t)(f + u) * h + r) * w8 + (c >> 3)] = (uint8_t)bad_bits;
        }
    }
}

// any alignment / any box: one thread per output pixel
template <typename BG, typename ACC, bool HAS_BG>
__global__ void __launch_bounds__(kPrepThreads)
prep_scalar_kernel(const int16_t *__restrict__ frames, int n, int H, int W,
                   const BG *__restrict__ bground, const uint8_t *__restrict__ roi,
                   int y0, int x0, int h, int w, PrepMath<ACC> math, int flags,
                   uint8_t *__restrict__ out, int32_t *__restrict__ invalid_count, uint32_t *__restrict__ invalid_bits_words,
                   uint32_t *__restrict__ positive_words) {
    const size_t total = (size_t)n * h * w;
    const int bytes_per_row = (w + 7) >> 3;
    const int wpr = (w + 31) >> 5;
    for (size_t i = (size_t)blockIdx.x * kPrepThreads + threadIdx.x; i < total;
         i += (size_t)gridDim.x * kPrepThreads) {
        const int c = (int)(i % w);
        const size_t t = i / w;
        const int r = (int)(t % h);
        const int f = (int)(t / h);
        const size_t pix = (size_t)(y0 + r) * W + (x0 + c);
        const int v = frames[(size_t)f * H * W + pix];
        const ACC rm = roi ? (ACC)(roi[pix] != 0) : (ACC)1;
        uint32_t o;
        if constexpr (std::is_same<ACC, int>::value)
            o = math.run(HAS_BG ? (int)bground[pix] - v : v, v, (int)rm, flags);
        else
            o = math.run((ACC)bground[pix], v, rm, flags);
        out[i] = (uint8_t)o;
        if (positive_words && (uint8_t)o != 0) atomicOr(positive_words + ((size_t)f * h + r) * wpr + (c >> 5), 1u << (c & 31));
        if (v == 0 && rm != (ACC)0) {
            if (invalid_count) atomicAdd(invalid_count + f, 1);
            if (invalid_bits_words) {           // byte-packed mask, set with word atomics (buffer zeroed by the launcher)
                const size_t byte = ((size_t)f * h + r) * bytes_per_row + (c >> 3);
                atomicOr(invalid_bits_words + (byte >> 2), 1u << ((byte & 3) * 8 + (c & 7)));
            }
        }
    }
}

template <typename BG, typename ACC, bool HAS_BG>
int launch_prep(const int16_t *frames, int n, int H, int W, const void *bground, const uint8_t *roi,
                int y0, int x0, int h, int w, double vmin, double vmax, int flags, uint8_t *out,
                int32_t *invalid, uint8_t *invalid_bits, uint8_t *positive_bits, cudaStream_t st) {
    PrepMath<ACC> math(vmin, vmax);
    const bool vec_ok = (w % 8 == 0) && (x0 % 8 == 0) && (W % 8 == 0) &&
                        ((uintptr_t)frames % 16 == 0) && ((uintptr_t)out % 8 == 0);
    if (vec_ok) {
        const int positions = h * (w / 8);
        const int bx = (positions + kPrepThreads - 1) / kPrepThreads;
        int groups = (sm_count() * 8 + bx - 1) / bx;
        groups = max(1, min(groups, (n + kPrepUnroll - 1) / kPrepUnroll));
        const int fpg = (n + groups - 1) / groups;
        groups = (n + fpg - 1) / fpg;
        dim3 grid(bx, groups);
        TimedLaunch timed(K_PREP, st);
        if constexpr (std::is_same<BG, float>::value && HAS_BG) {
            const float hi = (float)vmax;
            if ((flags & MSQ_PREP_HAS_VMIN) && (flags & MSQ_PREP_HAS_VMAX) && (float)vmin == 0.0f && hi >= 0.0f && hi <= 255.0f) {
                prep_vec8_f32_fast_kernel<<<grid, kPrepThreads, 0, st>>>(frames, n, H, W, (const float *)bground, roi, y0, x0, h, w,
                                                                         hi, fpg, out, invalid, invalid_bits, positive_bits);
                MSQ_LAUNCH_OK("prep_frames (float32 fast path)");
                return MSQ_OK;
            }
        }
        prep_vec8_kernel<BG, ACC, HAS_BG><<<grid, kPrepThreads, 0, st>>>(
            frames, n, H, W, (const BG *)bground, roi, y0, x0, h, w, math, flags, fpg, out, invalid, invalid_bits, positive_bits);
    } else {
        const size_t total = (size_t)n * h * w;
        if (invalid_bits) {
            MSQ_REQUIRE((uintptr_t)invalid_bits % 4 == 0, MSQ_EINVAL, "msq_prep_frames: invalid_bits must be 4-byte aligned");
            const size_t bytes = align_up((size_t)n * h * ((w + 7) / 8), 4);
            MSQ_CUDA_OK(cudaMemsetAsync(invalid_bits, 0, bytes, st));
        }
        if (positive_bits) MSQ_CUDA_OK(cudaMemsetAsync(positive_bits, 0, (size_t)n * h * positive_row_bytes(w), st));
        const int blocks = (int)std::min<size_t>((total + kPrepThreads - 1) / kPrepThreads, (size_t)sm_count() * 16);
        TimedLaunch timed(K_PREP, st);
        prep_scalar_kernel<BG, ACC, HAS_BG><<<blocks, kPrepThreads, 0, st>>>(
            frames, n, H, W, (const BG *)bground, roi, y0, x0, h, w, math, flags, out, invalid, reinterpret_cast<uint32_t *>(invalid_bits),
            reinterpret_cast<uint32_t *>(positive_bits));
    }
    MSQ_LAUNCH_OK("prep_frames");
    return MSQ_OK;
}

// ------------------------------- scale ---------------------------------------------------------
// (build_scale_lut lives in common.cuh: the stem kernels share it)
__global__ void __launch_bounds__(256)
scale_u8_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t count, double vmin,
                double vmax, int vmin_is_int, int vec_ok) {
    __shared__ uint8_t lut[256];
    build_scale_lut(lut, vmin, vmax, vmin_is_int);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (vec_ok) {
        const size_t nvec = count / 16;
        for (size_t i = tid; i < nvec; i += nthreads) {
            uint4 v = ldg_stream_u4(in + i * 16);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                w[k] = (uint32_t)lut[w[k] & 0xff] | ((uint32_t)lut[(w[k] >> 8) & 0xff] << 8) |
                       ((uint32_t)lut[(w[k] >> 16) & 0xff] << 16) | ((uint32_t)lut[w[k] >> 24] << 24);
            stg_stream_u4(out + i * 16, make_uint4(w[0], w[1], w[2], w[3]));
        }
        done = nvec * 16;
    }
    for (size_t i = done + tid; i < count; i += nthreads) out[i] = lut[in[i]];
}

__global__ void __launch_bounds__(256)
scale_chw3_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, int n, size_t plane, double vmin,
                  double vmax, int vmin_is_int, int vec_ok) {
    __shared__ uint8_t lut[256];
    build_scale_lut(lut, vmin, vmax, vmin_is_int);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    if (vec_ok) {
        const size_t q = plane / 4;
        for (size_t i = tid; i < (size_t)n * q; i += nthreads) {
            const size_t f = i / q, p = (i - f * q) * 4;
            const uint32_t v = ldg_stream_u1(in + f * plane + p);
            const float4 o = make_float4((float)lut[v & 0xff], (float)lut[(v >> 8) & 0xff],
                                         (float)lut[(v >> 16) & 0xff], (float)lut[v >> 24]);
            float *dst = out + f * 3 * plane + p;
            *reinterpret_cast<float4 *>(dst) = o;
            *reinterpret_cast<float4 *>(dst + plane) = o;
            *reinterpret_cast<float4 *>(dst + 2 * plane) = o;
        }
    } else {
        for (size_t i = tid; i < (size_t)n * plane; i += nthreads) {
            const size_t f = i / plane, p = i - f * plane;
            const float o = (float)lut[in[i]];
            float *dst = out + f * 3 * plane + p;
            dst[0] = o; dst[plane] = o; dst[2 * plane] = o;
        }
    }
}

// a3 + the detector's image transform in one pass: intensity scaling (LUT), 3-channel replication, per-channel
// normalisation (x - mean) / std, bilinear resize (torch upsample_bilinear2d, align_corners = false, scale = in / out) and
// zero padding to the stride-aligned canvas, written channels-last in the dtype the backbone's first convolution reads.
// Replaces six full-tensor passes (scale -> stack -> subtract -> divide -> interpolate -> pad -> cast) by one.
struct NormArg { float mean[3], std[3]; };

template <typename T>
__global__ void __launch_bounds__(256)
detector_input_kernel(const uint8_t *__restrict__ in, T *__restrict__ out, int n, int h, int w, int oh, int ow, int ph, int pw,
                      NormArg norm, double vmin, double vmax, int vmin_is_int) {
    __shared__ uint8_t lut8[256];
    __shared__ float lut[3][256];
    build_scale_lut(lut8, vmin, vmax, vmin_is_int);
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
        const int c = i >> 8, v = i & 255;
        lut[c][v] = ((float)lut8[v] - norm.mean[c]) / norm.std[c];               // torchvision: (image - mean) / std
    }
    __syncthreads();
    const float rh = (float)h / (float)oh, rw = (float)w / (float)ow;
    const size_t total = (size_t)n * ph * pw;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / ((size_t)ph * pw);
        const int p = (int)(i - f * (size_t)ph * pw), y = p / pw, x = p - y * pw;
        float o[3] = {0.f, 0.f, 0.f};
        if (y < oh && x < ow) {
            const float sy = fmaxf(rh * ((float)y + 0.5f) - 0.5f, 0.f), sx = fmaxf(rw * ((float)x + 0.5f) - 0.5f, 0.f);
            const int y1 = (int)sy, x1 = (int)sx;
            const int yp = y1 < h - 1 ? 1 : 0, xp = x1 < w - 1 ? 1 : 0;
            const float ly = sy - (float)y1, lx = sx - (float)x1, hy = 1.f - ly, hx = 1.f - lx;
            const uint8_t *src = in + f * (size_t)h * w + (size_t)y1 * w + x1;
            const int a = src[0], b = src[xp], c = src[(size_t)yp * w], d = src[(size_t)yp * w + xp];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                o[k] = hy * (hx * lut[k][a] + lx * lut[k][b]) + ly * (hx * lut[k][c] + lx * lut[k][d]);
        }
        T *dst = out + i * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) dst[k] = (T)o[k];
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// The R-CNN's stem in one kernel: a3 intensity scaling -> (x - mean) / std -> zero padding -> 7x7 stride-2 convolution
// (+ folded FrozenBN bias) -> ReLU -> 3x3 stride-2 max-pool, from the prepared uint8 chunk to the (n, 64, 64, 64)
// channels-last input of res2 (ref: model/predict.py:74-77 replicates the grey channel three times; detectron2 BasicStem).
// The three input channels are copies of one grey image and PIXEL_MEAN / PIXEL_STD are the same for all of them, so the
// 3-channel convolution equals a 1-channel convolution with the weights summed over the input channels (K = 49 instead of
// 147; exact algebra): cuDNN needs 2.9 ms + 0.9 ms (pool) per 250 frames for the 3-channel form, bound by the C = 3 layout.
// Direct convolution on the FP32 pipe: a CTA owns 8x8 pooled pixels = 17x17 convolution outputs = a 39x39 input patch (kept
// as even / odd column planes so that stride-2 reads are conflict-free); a warp owns 16 output channels (weights are warp
// broadcasts) and half of the pixels, a thread 5 pixels x 16 channels = 80 accumulators.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kStemPool = 8;                         // pooled pixels per tile side
constexpr int kStemConv = 2 * kStemPool + 1;         // 17 convolution outputs per tile side
constexpr int kStemIn = 2 * (kStemConv - 1) + 7;     // 39 input pixels per tile side
constexpr int kStemHalf = (kStemIn + 1) / 2;         // 20 columns per parity plane
constexpr int kStemPix = kStemConv * kStemConv;      // 289
constexpr int kStemPixPerThread = 5;
constexpr int kStemCstride = 72;                     // channels per stored pixel (64 + padding against bank conflicts)

template <typename T> struct StemStore;
template <> struct StemStore<float> {
    static __device__ __forceinline__ float make(float v) { return v; }
    static __device__ __forceinline__ float get(float v) { return v; }
};
template <> struct StemStore<__nv_bfloat16> {
    static __device__ __forceinline__ __nv_bfloat16 make(float v) { return __float2bfloat16_rn(v); }
    static __device__ __forceinline__ float get(__nv_bfloat16 v) { return __bfloat162float(v); }
};

template <typename T>
__global__ void __launch_bounds__(256)
stem_conv_pool_kernel(const uint8_t *__restrict__ in, int h, int w, int conv_h, int conv_w, int pool_h, int pool_w, int tiles_x,
                      float mean, float stdv, double vmin, double vmax, int vmin_is_int, const float *__restrict__ w49x64,
                      const float *__restrict__ bias64, T *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char stem_smem[];
    float *wsm = reinterpret_cast<float *>(stem_smem);                         // [49][64]
    float *even = wsm + 49 * 64;                                                // [39][20] input columns 0, 2, 4, ...
    float *odd = even + kStemIn * kStemHalf;                                    // [39][20] input columns 1, 3, 5, ...
    float *lutf = odd + kStemIn * kStemHalf;                                    // [256] normalised value of every grey level
    T *conv = reinterpret_cast<T *>(lutf + 256);                                // [289][72]
    __shared__ uint8_t lut8[256];
    build_scale_lut(lut8, vmin, vmax, vmin_is_int);
    for (int i = threadIdx.x; i < 256; i += 256) lutf[i] = ((float)lut8[i] - mean) / stdv;
    for (int i = threadIdx.x; i < 49 * 64; i += 256) wsm[i] = w49x64[i];
    __syncthreads();
    const int img = blockIdx.x, ty = blockIdx.y / tiles_x, tx = blockIdx.y - ty * tiles_x;
    const int py0 = ty * kStemPool, px0 = tx * kStemPool;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;                             // first convolution output of the tile
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;                             // first input pixel of the tile
    const uint8_t *src = in + (size_t)img * h * w;
    for (int i = threadIdx.x; i < kStemIn * kStemIn; i += 256) {
        const int j = i / kStemIn, k = i - j * kStemIn;
        const int gy = iy0 + j, gx = ix0 + k;
        const float v = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? lutf[src[(size_t)gy * w + gx]] : 0.f;   // padding is 0 after normalisation
        ((k & 1) ? odd : even)[j * kStemHalf + (k >> 1)] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cg = (warp & 3) * 16;                                             // this warp's 16 output channels
    const int pbase = (warp >> 2) * (kStemPixPerThread * 32);                   // pixels 0..159 / 160..319 (289 in use)
    int off[kStemPixPerThread];
#pragma unroll
    for (int k = 0; k < kStemPixPerThread; ++k) {
        const int p = min(pbase + k * 32 + lane, kStemPix - 1);
        const int oy = p / kStemConv, ox = p - oy * kStemConv;
        off[k] = 2 * oy * kStemHalf + ox;
    }
    float acc[kStemPixPerThread][16];
#pragma unroll
    for (int k = 0; k < kStemPixPerThread; ++k)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[k][c] = 0.f;
#pragma unroll 1
    for (int r = 0; r < 7; ++r) {
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            const float *plane = (t & 1) ? odd : even;
            const int shift = r * kStemHalf + (t >> 1);
            float xin[kStemPixPerThread];
#pragma unroll
            for (int k = 0; k < kStemPixPerThread; ++k) xin[k] = plane[off[k] + shift];
            const float4 *wv = reinterpret_cast<const float4 *>(wsm + (r * 7 + t) * 64 + cg);
            float wt[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const float4 v4 = wv[q]; wt[4 * q] = v4.x; wt[4 * q + 1] = v4.y; wt[4 * q + 2] = v4.z; wt[4 * q + 3] = v4.w; }
#pragma unroll
            for (int k = 0; k < kStemPixPerThread; ++k)
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[k][c] = fmaf(xin[k], wt[c], acc[k][c]);
        }
    }
    // bias + ReLU; convolution outputs outside the map count as 0 (every pool window holds a real output, and those are >= 0)
#pragma unroll
    for (int k = 0; k < kStemPixPerThread; ++k) {
        const int p = pbase + k * 32 + lane;
        if (p >= kStemPix) continue;
        const int oy = p / kStemConv, ox = p - oy * kStemConv;
        const int gy = cy0 + oy, gx = cx0 + ox;
        const bool inside = gy >= 0 && gy < conv_h && gx >= 0 && gx < conv_w;
#pragma unroll
        for (int c = 0; c < 16; ++c)
            conv[p * kStemCstride + cg + c] = StemStore<T>::make(inside ? fmaxf(acc[k][c] + bias64[cg + c], 0.f) : 0.f);
    }
    __syncthreads();
    // 3x3 stride-2 max-pool of the tile, 8 channels per thread, written as contiguous channels-last rows
    for (int i = threadIdx.x; i < kStemPool * kStemPool * 8; i += 256) {
        const int v = i & 7, pp = i >> 3, py = pp / kStemPool, px = pp - py * kStemPool;
        const int gy = py0 + py, gx = px0 + px;
        if (gy >= pool_h || gx >= pool_w) continue;
        float m[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) m[c] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const T *cp = conv + ((2 * py + dy) * kStemConv + 2 * px + dx) * kStemCstride + v * 8;
#pragma unroll
                for (int c = 0; c < 8; ++c) m[c] = fmaxf(m[c], StemStore<T>::get(cp[c]));
            }
        T *dst = out + (((size_t)img * pool_h + gy) * pool_w + gx) * 64 + v * 8;
        if constexpr (sizeof(T) == 2) {
            uint32_t pk[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(m[2 * c], m[2 * c + 1]);
                pk[c] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        } else {
            *reinterpret_cast<float4 *>(dst) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4 *>(dst + 4) = make_float4(m[4], m[5], m[6], m[7]);
        }
    }
}


// Host-side instance masks travel as bit rows (1/8 of the bytes over PCIe); the extract kernels read {0,1} bytes.
// bits (n, h, ceil(w/8)) with bit b of byte B of a row = pixel 8B + b  ->  out (n, h, w) u8.  A thread expands one byte.
__global__ void __launch_bounds__(256)
unpack_mask_bits_kernel(const uint8_t *__restrict__ bits, size_t rows, int w, int wb, uint8_t *__restrict__ out) {
    const size_t total = rows * (size_t)wb;
    const bool vec = (w % 8 == 0) && ((uintptr_t)out % 8 == 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / wb;
        const int B = (int)(i - row * wb);
        const uint32_t v = bits[i];
        uint8_t *dst = out + row * (size_t)w + (size_t)B * 8;
        if (vec) {
            // spread 8 bits to 8 bytes: multiply trick on two nibbles
            const uint32_t lo = ((v & 0xfu) * 0x00204081u) & 0x01010101u, hi = (((v >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
            *reinterpret_cast<uint2 *>(dst) = make_uint2(lo, hi);
        } else {
            for (int b = 0; b < 8 && B * 8 + b < w; ++b) dst[b] = (v >> b) & 1u;
        }
    }
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" int msq_detector_input(const uint8_t *in, void *out, int out_is_bf16, int n, int h, int w, int oh, int ow, int ph, int pw,
                                  const float *mean_host, const float *std_host, double vmin, double vmax, int vmin_is_int,
                                  void *stream) {
    MSQ_REQUIRE(in && out && mean_host && std_host, MSQ_EINVAL, "msq_detector_input: null pointer");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && ph >= oh && pw >= ow, MSQ_EINVAL,
                "msq_detector_input: bad sizes n=%d %dx%d -> %dx%d in %dx%d", n, h, w, oh, ow, ph, pw);
    MSQ_REQUIRE(vmax != vmin, MSQ_EINVAL, "msq_detector_input: vmax == vmin");
    if (n == 0) return MSQ_OK;
    NormArg norm;
    for (int c = 0; c < 3; ++c) {
        MSQ_REQUIRE(std_host[c] != 0.f, MSQ_EINVAL, "msq_detector_input: std[%d] == 0", c);
        norm.mean[c] = mean_host[c]; norm.std[c] = std_host[c];
    }
    const size_t work = (size_t)n * ph * pw;
    const int blocks = (int)std::min<size_t>((work + 255) / 256, (size_t)sm_count() * 16);
    TimedLaunch timed(K_DETECTOR_GLUE, (cudaStream_t)stream);
    if (out_is_bf16)
        detector_input_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, static_cast<__nv_bfloat16 *>(out), n, h, w, oh, ow,
                                                                                       ph, pw, norm, vmin, vmax, vmin_is_int);
    else
        detector_input_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(in, static_cast<float *>(out), n, h, w, oh, ow, ph, pw, norm,
                                                                               vmin, vmax, vmin_is_int);
    MSQ_LAUNCH_OK("detector_input");
    return MSQ_OK;
}

extern "C" int msq_stem_conv_pool(const uint8_t *in, int n, int h, int w, int ph, int pw, double vmin, double vmax, int vmin_is_int,
                                  float mean, float stdv, const float *w49x64, const float *bias64, void *out, int out_is_bf16,
                                  void *stream) {
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0 && ph >= h && pw >= w, MSQ_EINVAL, "msq_stem_conv_pool: bad sizes n=%d %dx%d in %dx%d", n, h, w, ph, pw);
    MSQ_REQUIRE(vmax != vmin && stdv != 0.f, MSQ_EINVAL, "msq_stem_conv_pool: vmax == vmin or std == 0");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(in && w49x64 && bias64 && out, MSQ_EINVAL, "msq_stem_conv_pool: null pointer");
    MSQ_REQUIRE((uintptr_t)out % 16 == 0, MSQ_EINVAL, "msq_stem_conv_pool: output must be 16-byte aligned");
    const int conv_h = (ph + 6 - 7) / 2 + 1, conv_w = (pw + 6 - 7) / 2 + 1;
    const int pool_h = (conv_h + 2 - 3) / 2 + 1, pool_w = (conv_w + 2 - 3) / 2 + 1;
    const int tiles_x = (pool_w + kStemPool - 1) / kStemPool, tiles_y = (pool_h + kStemPool - 1) / kStemPool;
    const size_t fixed = (size_t)(49 * 64 + 2 * kStemIn * kStemHalf + 256) * sizeof(float);
    const size_t smem = fixed + (size_t)kStemPix * kStemCstride * (out_is_bf16 ? 2 : 4);
    cudaStream_t st = (cudaStream_t)stream;
    TimedLaunch timed(K_DETECTOR_GLUE, st);
    if (out_is_bf16) {
        MSQ_CUDA_OK(cudaFuncSetAttribute(stem_conv_pool_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stem_conv_pool_kernel<__nv_bfloat16><<<dim3(n, tiles_x * tiles_y), 256, smem, st>>>(in, h, w, conv_h, conv_w, pool_h, pool_w, tiles_x, mean,
            stdv, vmin, vmax, vmin_is_int, w49x64, bias64, static_cast<__nv_bfloat16 *>(out));
    } else {
        MSQ_CUDA_OK(cudaFuncSetAttribute(stem_conv_pool_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stem_conv_pool_kernel<float><<<dim3(n, tiles_x * tiles_y), 256, smem, st>>>(in, h, w, conv_h, conv_w, pool_h, pool_w, tiles_x, mean, stdv,
            vmin, vmax, vmin_is_int, w49x64, bias64, static_cast<float *>(out));
    }
    MSQ_LAUNCH_OK("stem_conv_pool");
    return MSQ_OK;
}

extern "C" int msq_unpack_mask_bits(const uint8_t *bits, int n, int h, int w, uint8_t *out, void *stream) {
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0, MSQ_EINVAL, "msq_unpack_mask_bits: bad sizes");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(bits && out, MSQ_EINVAL, "msq_unpack_mask_bits: null pointer");
    const int wb = (w + 7) / 8;
    const size_t rows = (size_t)n * h, work = rows * wb;
    const int blocks = (int)std::min<size_t>((work + 255) / 256, (size_t)sm_count() * 16);
    TimedLaunch timed(K_PREP, (cudaStream_t)stream);
    unpack_mask_bits_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(bits, rows, w, wb, out);
    MSQ_LAUNCH_OK("unpack_mask_bits");
    return MSQ_OK;
}

// Host -> device copy of the ROI box only (the part of a raw frame the path needs, SURVEY section 8d: 2A of R bytes): one strided
// 3-D DMA transfer per chunk -- rows of w int16 out of every frame's rows [y0, y0 + h) -- into a dense (n, h, w) int16 device
// array, which msq_prep_frames then takes with H = h, W = w, y0 = x0 = 0 and the cropped background / ROI.  The copy engine
// moves these rows at close to its plain-copy rate, while a kernel that reads them from pinned memory itself (the zero-copy
// path) gets ~70 % of it; neither touches the 74 % of the frame outside the box.
extern "C" int msq_copy_roi_rows(const int16_t *frames_host, int n, int H, int W, int y0, int x0, int h, int w, int16_t *out_dev, void *stream) {
    MSQ_REQUIRE(n >= 0 && H > 0 && W > 0 && h > 0 && w > 0 && y0 >= 0 && x0 >= 0 && y0 + h <= H && x0 + w <= W, MSQ_EINVAL,
                "msq_copy_roi_rows: box (y0=%d,x0=%d,h=%d,w=%d) outside the %dx%d frame", y0, x0, h, w, H, W);
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(frames_host && out_dev, MSQ_EINVAL, "msq_copy_roi_rows: null pointer");
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(const_cast<int16_t *>(frames_host), (size_t)W * sizeof(int16_t), (size_t)W, (size_t)H);
    p.srcPos = make_cudaPos((size_t)x0 * sizeof(int16_t), (size_t)y0, 0);
    p.dstPtr = make_cudaPitchedPtr(out_dev, (size_t)w * sizeof(int16_t), (size_t)w, (size_t)h);
    p.dstPos = make_cudaPos(0, 0, 0);
    p.extent = make_cudaExtent((size_t)w * sizeof(int16_t), (size_t)h, (size_t)n);
    p.kind = cudaMemcpyHostToDevice;
    MSQ_CUDA_OK(cudaMemcpy3DAsync(&p, (cudaStream_t)stream));
    return MSQ_OK;
}

// The same for a ROI that is not a rectangle (the reference's bucket floor is a disc: 78 % of its bounding box): the box is cut
// into n_bands horizontal bands, band b = rows [band_y[b], band_y[b + 1]) x columns [band_x0[b], band_x1[b]) (all relative to
// the box; the caller makes every band cover the ROI pixels of its rows), one strided DMA transfer per band into the same dense
// (n, h, w) device array.  Pixels outside the bands are not written: msq_prep_frames multiplies by the ROI mask, so whatever
// finite int16 values they hold never reach its output (keep the array zero-initialised).  16 bands of a disc move 15 % fewer
// bytes than the box.
extern "C" int msq_copy_roi_bands(const int16_t *frames_host, int n, int H, int W, int y0, int x0, int h, int w, const int *band_y,
                                  const int *band_x0, const int *band_x1, int n_bands, int16_t *out_dev, void *stream) {
    MSQ_REQUIRE(n >= 0 && H > 0 && W > 0 && h > 0 && w > 0 && y0 >= 0 && x0 >= 0 && y0 + h <= H && x0 + w <= W, MSQ_EINVAL,
                "msq_copy_roi_bands: box (y0=%d,x0=%d,h=%d,w=%d) outside the %dx%d frame", y0, x0, h, w, H, W);
    MSQ_REQUIRE(n_bands >= 1 && band_y && band_x0 && band_x1, MSQ_EINVAL, "msq_copy_roi_bands: no bands");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(frames_host && out_dev, MSQ_EINVAL, "msq_copy_roi_bands: null pointer");
    for (int b = 0; b < n_bands; ++b) {
        const int ya = band_y[b], yb = band_y[b + 1], xa = band_x0[b], xb = band_x1[b];
        MSQ_REQUIRE(ya >= 0 && ya <= yb && yb <= h && xa >= 0 && xa <= xb && xb <= w, MSQ_EINVAL,
                    "msq_copy_roi_bands: band %d (rows %d..%d, columns %d..%d) outside the %dx%d box", b, ya, yb, xa, xb, h, w);
        if (ya == yb || xa == xb) continue;
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr(const_cast<int16_t *>(frames_host), (size_t)W * sizeof(int16_t), (size_t)W, (size_t)H);
        p.srcPos = make_cudaPos((size_t)(x0 + xa) * sizeof(int16_t), (size_t)(y0 + ya), 0);
        p.dstPtr = make_cudaPitchedPtr(out_dev, (size_t)w * sizeof(int16_t), (size_t)w, (size_t)h);
        p.dstPos = make_cudaPos((size_t)xa * sizeof(int16_t), (size_t)ya, 0);
        p.extent = make_cudaExtent((size_t)(xb - xa) * sizeof(int16_t), (size_t)(yb - ya), (size_t)n);
        p.kind = cudaMemcpyHostToDevice;
        MSQ_CUDA_OK(cudaMemcpy3DAsync(&p, (cudaStream_t)stream));
    }
    return MSQ_OK;
}

extern "C" int msq_prep_frames_bits(const int16_t *frames, int n, int H, int W, const void *bground, int bg_dtype,
                                    const uint8_t *roi, int y0, int x0, int h, int w, double vmin, double vmax,
                                    int flags, uint8_t *out, int32_t *invalid, uint8_t *invalid_bits, uint32_t *positive_bits,
                                    void *stream) {
    MSQ_REQUIRE(n == 0 || (frames && out), MSQ_EINVAL, "msq_prep_frames: null frames/out pointer");
    MSQ_REQUIRE(n >= 0 && H > 0 && W > 0 && h > 0 && w > 0, MSQ_EINVAL,
                "msq_prep_frames: bad sizes n=%d H=%d W=%d h=%d w=%d", n, H, W, h, w);
    MSQ_REQUIRE(y0 >= 0 && x0 >= 0 && y0 + h <= H && x0 + w <= W, MSQ_EINVAL,
                "msq_prep_frames: box (y0=%d,x0=%d,h=%d,w=%d) outside the %dx%d frame", y0, x0, h, w, H, W);
    MSQ_REQUIRE(bg_dtype >= MSQ_BG_NONE && bg_dtype <= MSQ_BG_U16, MSQ_EINVAL, "msq_prep_frames: bad bg_dtype %d", bg_dtype);
    MSQ_REQUIRE(bg_dtype == MSQ_BG_NONE || bground, MSQ_EINVAL, "msq_prep_frames: bground is null but bg_dtype=%d", bg_dtype);
    MSQ_REQUIRE((uintptr_t)positive_bits % 4 == 0, MSQ_EINVAL, "msq_prep_frames: positive_bits must be 4-byte aligned");
    if (n == 0) return MSQ_OK;
    uint8_t *pos = reinterpret_cast<uint8_t *>(positive_bits);
    cudaStream_t st = (cudaStream_t)stream;
    if (invalid) MSQ_CUDA_OK(cudaMemsetAsync(invalid, 0, sizeof(int32_t) * (size_t)n, st));
    switch (bg_dtype) {
        case MSQ_BG_F32: return launch_prep<float, float, true>(frames, n, H, W, bground, roi, y0, x0, h, w, vmin, vmax, flags, out, invalid, invalid_bits, pos, st);
        case MSQ_BG_F64: return launch_prep<double, double, true>(frames, n, H, W, bground, roi, y0, x0, h, w, vmin, vmax, flags, out, invalid, invalid_bits, pos, st);
        case MSQ_BG_U16: return launch_prep<uint16_t, int, true>(frames, n, H, W, bground, roi, y0, x0, h, w, vmin, vmax, flags, out, invalid, invalid_bits, pos, st);
        default:         return launch_prep<uint16_t, int, false>(frames, n, H, W, nullptr, roi, y0, x0, h, w, vmin, vmax, flags, out, invalid, invalid_bits, pos, st);
    }
}

extern "C" size_t msq_positive_bits_bytes(int n, int h, int w) {
    return n > 0 && h > 0 && w > 0 ? (size_t)n * h * msq::positive_row_bytes(w) : 0;
}

extern "C" int msq_prep_frames(const int16_t *frames, int n, int H, int W, const void *bground, int bg_dtype,
                               const uint8_t *roi, int y0, int x0, int h, int w, double vmin, double vmax,
                               int flags, uint8_t *out, int32_t *invalid, uint8_t *invalid_bits, void *stream) {
    return msq_prep_frames_bits(frames, n, H, W, bground, bg_dtype, roi, y0, x0, h, w, vmin, vmax, flags, out, invalid, invalid_bits,
                                nullptr, stream);
}

extern "C" int msq_scale_frames(const uint8_t *in, uint8_t *out, size_t count, double vmin, double vmax,
                                int vmin_is_int, void *stream) {
    MSQ_REQUIRE(count == 0 || (in && out), MSQ_EINVAL, "msq_scale_frames: null pointer");
    MSQ_REQUIRE(vmax != vmin, MSQ_EINVAL, "msq_scale_frames: vmax == vmin");
    if (count == 0) return MSQ_OK;
    const int vec_ok = ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const size_t work = (count + 15) / 16;
    const int blocks = (int)std::min<size_t>((work + 255) / 256, (size_t)sm_count() * 8);
    TimedLaunch timed(K_SCALE, (cudaStream_t)stream);
    scale_u8_kernel<<<max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(in, out, count, vmin, vmax, vmin_is_int, vec_ok);
    MSQ_LAUNCH_OK("scale_frames");
    return MSQ_OK;
}

extern "C" int msq_scale_frames_chw3_f32(const uint8_t *in, float *out, int n, int h, int w, double vmin,
                                         double vmax, int vmin_is_int, void *stream) {
    MSQ_REQUIRE(in && out, MSQ_EINVAL, "msq_scale_frames_chw3_f32: null pointer");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0, MSQ_EINVAL, "msq_scale_frames_chw3_f32: bad sizes");
    MSQ_REQUIRE(vmax != vmin, MSQ_EINVAL, "msq_scale_frames_chw3_f32: vmax == vmin");
    if (n == 0) return MSQ_OK;
    const size_t plane = (size_t)h * w;
    const int vec_ok = (plane % 4 == 0) && ((uintptr_t)in % 4 == 0) && ((uintptr_t)out % 16 == 0);
    const size_t work = (size_t)n * plane / 4 + 1;
    const int blocks = (int)std::min<size_t>((work + 255) / 256, (size_t)sm_count() * 8);
    TimedLaunch timed(K_SCALE, (cudaStream_t)stream);
    scale_chw3_kernel<<<max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(in, out, n, plane, vmin, vmax, vmin_is_int, vec_ok);
    MSQ_LAUNCH_OK("scale_frames_chw3_f32");
    return MSQ_OK;
}
