// Shared host/device helpers for libmoseq_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/moseq_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmoseq_b200 is written for sm_100a (B200) only"
#endif

namespace msq {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);          // api.cu; thread-local message for msq_last_error()

#define MSQ_REQUIRE(cond, code, ...)                                   \
    do {                                                               \
        if (!(cond)) {                                                 \
            ::msq::set_error(__VA_ARGS__);                             \
            return (code);                                             \
        }                                                              \
    } while (0)

#define MSQ_CUDA_OK(expr)                                                                 \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::msq::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                             __FILE__, __LINE__);                                         \
            return MSQ_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

#define MSQ_LAUNCH_OK(name)                                                               \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            ::msq::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));    \
            return MSQ_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

int sm_count();                                 // api.cu; cached per process (current device)

// ---- per-kernel device timing (CUDA events on the launching stream; off by default) ----------
enum KernelId { K_PREP = 0, K_SCALE, K_CLEAN, K_FEATURES, K_ANGLES, K_MASKED_SUMS, K_SCALARS_KPTS, K_CROP, K_PASTE,
                K_INPAINT, K_KALMAN, K_BGROUND, K_ROI, K_DETECTOR_GLUE, K_COUNT };
struct TimedLaunch {                            // RAII: records an event pair around one step's kernel launch(es)
    int slot;
    cudaStream_t st;
    TimedLaunch(int kernel_id, cudaStream_t stream, int kernels = 1);      // kernels: how many launches the scope makes (the counter)
    ~TimedLaunch();
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__

// streaming 128-bit global load / store that do not pollute L1 (inputs/outputs are touched once)
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u1(const void *p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_u4(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_u2(void *p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream_u8(void *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u8 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg_stream_u1(void *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- SWAR byte tests (shared by features.cu and epilogue.cu) ------------------------------------
// SWAR byte tests on 4 packed pixels, result in bit 7 of every byte, no cross-byte carries:
//   c >= g  (1 <= g <= 255):  low 7 bits compared by adding (128 - (g & 127)); the top bit decides the rest
//   m != 0
struct ByteTest {
    uint32_t k;        // (0x80 - (g & 0x7f)) replicated
    uint32_t hi_or;    // g < 128: a set top bit alone passes;  g >= 128: the top bit is required
    int mode;          // 0: everything passes (g <= 0), 1: g in 1..127, 2: g in 128..255, 3: nothing passes
};
__device__ __forceinline__ ByteTest make_byte_test(int ge) {
    ByteTest t;
    t.mode = ge <= 0 ? 0 : (ge > 255 ? 3 : (ge < 128 ? 1 : 2));
    t.k = (uint32_t)(0x80 - (ge & 0x7f)) * 0x01010101u;
    t.hi_or = 0u;
    return t;
}
__device__ __forceinline__ uint32_t bytes_ge(uint32_t c, const ByteTest &t) {
    const uint32_t low = ((c & 0x7f7f7f7fu) + t.k);
    if (t.mode == 1) return (low | c) & 0x80808080u;
    if (t.mode == 2) return (low & c) & 0x80808080u;
    return t.mode == 0 ? 0x80808080u : 0u;
}
__device__ __forceinline__ uint32_t bytes_nonzero(uint32_t m) {
    return (((m & 0x7f7f7f7fu) + 0x7f7f7f7fu) | m) & 0x80808080u;
}
// ref proc/proc.py:214-234: float64 affine map then truncation.  256 possible inputs -> each CTA
// builds the table in shared memory with the reference's exact float64 operation order.
__device__ __forceinline__ void build_scale_lut(uint8_t *lut, double vmin, double vmax, int vmin_is_int) {
    for (int x = threadIdx.x; x < 256; x += blockDim.x) {
        const double gain = __ddiv_rn(255.0 - 0.0, __dsub_rn(vmax, vmin));
        double xv;
        if (vmin_is_int) xv = (double)(uint8_t)(x - (int)vmin);   // uint8 array - Python int wraps in uint8
        else xv = __dsub_rn((double)x, vmin);
        const double v = __dadd_rn(__dmul_rn(xv, gain), 0.0);
        lut[x] = (uint8_t)(long long)v;
    }
    __syncthreads();
}
#endif  // __CUDACC__
}  // namespace msq

// ---- internal launchers shared with the whole-chunk pipeline (pipeline.cu) ---------------------
namespace msq {
// Rows of a cleaned frame that can hold a non-zero pixel, as left behind by launch_clean() in its scratch: per (frame, 240-column
// tile) the rows [x, y] on which the opening's erosion is non-zero (y < 0: none); the opened frame is zero outside [x - 4, y + 4].
// bands == nullptr: not known, every row has to be read.
struct RowBands {
    const int2 *bands;
    int tiles_x;
};
int launch_clean(const uint8_t *in, uint8_t *out, int n, int h, int w, cudaStream_t st, int2 *bands = nullptr,
                 RowBands *written = nullptr, const uint32_t *positive_bits = nullptr);
// bands: clean_scratch_bytes() or null; *written: what the later kernels may use; positive_bits: msq_prep_frames' bit rows of
// `in` (a superset of its positive pixels is enough), only used together with bands
int launch_clean_stream(const uint8_t *in, uint8_t *out, int n, int h, int w, cudaStream_t st, int2 *bands, RowBands *written,
                        const uint32_t *positive_bits);   // -100: shape not served
size_t clean_scratch_bytes(int n, int w);
int launch_frame_features(const uint8_t *cleaned, const uint8_t *mask, int n, int h, int w, double frame_threshold,
                          double *centroid, double *orientation, double *axis, int64_t *sums24, int *fallback,
                          cudaStream_t st, cudaEvent_t after_stream = nullptr, RowBands rows = {nullptr, 0});   // after_stream: recorded on st after the streaming pass
int launch_masked_sums(const uint8_t *chunk_frames, const uint8_t *mask, int n, int h, int w, double min_h, double max_h,
                       int2 *sums_scratch, cudaStream_t st);
int launch_angles_and_flips(const double *orientation, const double *axis, const double *centroid, const float *kpts,
                            int n, int chunk, double *angle_out, uint8_t *flips, double *conf, int32_t *passes,
                            cudaStream_t st);
int launch_scalars_and_keypoints(const uint8_t *chunk_frames, const uint8_t *mask, const uint8_t *cleaned,
                                 const double *centroid, const double *angle_deg, const double *axis,
                                 const void *kpts, bool kpts_f64, int n, int h, int w, int chunk, double min_h,
                                 double max_h, double true_depth, double *scalars, double *kcols, int2 *sums_scratch,
                                 cudaStream_t st, bool sums_done = false);
int launch_crop_rotate(const uint8_t *src, const uint8_t *src2, int n, int h, int w, const double *centroid,
                       const double *angle_deg, int cw, int ch, uint8_t *out, uint8_t *out2, void *scratch, cudaStream_t st);
}  // namespace msq
