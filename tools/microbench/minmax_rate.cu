// Throughput of candidate min/max instructions on sm_100a: ops per clock per SM.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int MODE>
__global__ void bench(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t a[ILP], b[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed * (threadIdx.x + 1 + i); b[i] = seed ^ (0x9e3779b9u * (i + 1 + threadIdx.x)); }
    uint32_t c = seed * 7u + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) a[i] = __vminu2(a[i], b[i]) + 0;                       // VIMNMX.U16x2
            if (MODE == 1) a[i] = __vimin3_u16x2(a[i], b[i], c);                  // VIMNMX3.U16x2
            if (MODE == 2) a[i] = (uint32_t)min((int)a[i], (int)b[i]);            // VIMNMX s32
            if (MODE == 3) a[i] = (uint32_t)__vimin3_s32((int)a[i], (int)b[i], (int)c);   // VIMNMX3 s32
            if (MODE == 4) { __half2 x = *reinterpret_cast<__half2 *>(&a[i]), y = *reinterpret_cast<__half2 *>(&b[i]);
                             x = __hmin2(x, y); a[i] = *reinterpret_cast<uint32_t *>(&x); }   // HMNMX2
            if (MODE == 5) { float x = __uint_as_float(a[i]); x = fminf(x, __uint_as_float(b[i])); a[i] = __float_as_uint(x); }  // FMNMX
            if (MODE == 6) a[i] = __vminu4(a[i], b[i]);                           // emulated byte SIMD
            if (MODE == 7) a[i] = __funnelshift_r(a[i], b[i], 16);                // SHF
            if (MODE == 8) a[i] = __byte_perm(a[i], b[i], 0x5432);                // PRMT
            if (MODE == 9) a[i] = (a[i] & b[i]) | c;                              // LOP3
            if (MODE == 10) a[i] = a[i] * b[i] + c;                               // IMAD
            if (MODE == 11) { __half2 x = *reinterpret_cast<__half2 *>(&a[i]), y = *reinterpret_cast<__half2 *>(&b[i]);
                              x = __hmax2(__hmin2(x, y), *reinterpret_cast<__half2 *>(&c)); a[i] = *reinterpret_cast<uint32_t *>(&x); }
            if (MODE == 12) { __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162 *>(&a[i]), y = *reinterpret_cast<__nv_bfloat162 *>(&b[i]);
                              x = __hmin2(x, y); a[i] = *reinterpret_cast<uint32_t *>(&x); }   // HMNMX2.BF16
            if (MODE == 13) a[i] = (uint32_t)__viaddmin_s16x2(a[i], b[i], c);     // VIADDMNMX 16x2
            if (MODE == 14) a[i] = __vmaxu2(__vminu2(a[i], b[i]), c);
            b[i] ^= a[i] >> 31 ? 0u : 0u;
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char *name, int ops_per_iter) {
    uint32_t *out; long long *cyc, h;
    const int threads = 1024, blocks = 148 * 2;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, 8);
    bench<MODE><<<blocks, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // per SM: 2 blocks x 1024 threads = 64 warps; instr per warp = ITERS*ILP*ops_per_iter
    double warp_instr_per_sm = 64.0 * ITERS * ILP * ops_per_iter;
    printf("%-28s %8.3f ms  block0 cycles %10lld  -> %.2f warp-instr/clk/SM (%.0f lanes/clk/SM)\n", name, ms, h,
           warp_instr_per_sm / (double)h, 32.0 * warp_instr_per_sm / (double)h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("VIMNMX.U16x2 (vminu2)", 1);
    run<1>("VIMNMX3.U16x2", 1);
    run<2>("VIMNMX.S32", 1);
    run<3>("VIMNMX3.S32", 1);
    run<4>("HMNMX2 (half2 min)", 1);
    run<5>("FMNMX (f32 min)", 1);
    run<6>("vminu4 (emulated)", 1);
    run<7>("SHF funnelshift", 1);
    run<8>("PRMT", 1);
    run<9>("LOP3", 1);
    run<10>("IMAD", 1);
    run<11>("HMNMX2 min+max pair", 2);
    run<12>("HMNMX2.BF16", 1);
    run<13>("VIADDMNMX.S16x2", 1);
    run<14>("vminu2+vmaxu2 pair", 2);
    return 0;
}
