// Which instructions share an execution pipe on sm_100a?  Rate of instruction mixes, warp-instr/clk/SM.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
constexpr int ITERS = 4096;
constexpr int ILP = 6;
__device__ __forceinline__ uint32_t hmin2u(uint32_t a, uint32_t b) {
    __half2 x = *reinterpret_cast<__half2 *>(&a), y = *reinterpret_cast<__half2 *>(&b);
    x = __hmin2(x, y); return *reinterpret_cast<uint32_t *>(&x);
}
template <int MODE>
__global__ void bench(uint32_t *out, uint32_t seed, uint32_t one, uint32_t neg1, uint32_t k16) {
    uint32_t a[ILP], b[ILP], c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed * (threadIdx.x + 1 + i); b[i] = seed ^ (0x9e3779b9u * (i + 1 + threadIdx.x)); c[i] = a[i] ^ b[i]; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            // op X on a[i], op Y on c[i]; independent chains
            if (MODE == 0) { a[i] = __vimin3_u16x2(a[i], b[i], c[i]); }
            if (MODE == 1) { a[i] = __vminu2(a[i], b[i]); }
            if (MODE == 2) { a[i] = __vimin3_u16x2(a[i], b[i], b[i] + 1); c[i] = __vminu2(c[i], b[i]); }
            if (MODE == 3) { a[i] = __vimin3_u16x2(a[i], b[i], b[i] + 1); c[i] = __funnelshift_r(c[i], b[i], 16); }
            if (MODE == 4) { a[i] = __vminu2(a[i], b[i]); c[i] = __funnelshift_r(c[i], b[i], 16); }
            if (MODE == 5) { a[i] = hmin2u(a[i], b[i]); c[i] = __funnelshift_r(c[i], b[i], 16); }
            if (MODE == 6) { a[i] = hmin2u(a[i], b[i]); c[i] = __vimin3_u16x2(c[i], b[i], b[i] + 1); }
            if (MODE == 7) { a[i] = __vminu2(a[i], b[i]); c[i] = c[i] * b[i] + a[i]; }
            if (MODE == 8) { a[i] = hmin2u(a[i], b[i]); c[i] = c[i] * b[i] + 7u; }
            if (MODE == 9) { a[i] = __vimin3_u16x2(a[i], b[i], b[i] + 1); c[i] = c[i] * b[i] + 7u; }
            if (MODE == 10) { a[i] = hmin2u(a[i], b[i]); c[i] = __vminu2(c[i], b[i]); }
            if (MODE == 11) { a[i] = __funnelshift_r(a[i], b[i], 16); c[i] = c[i] * b[i] + 7u; }
            if (MODE == 12) { a[i] = hmin2u(a[i], b[i]); }
            if (MODE == 13) { a[i] = __funnelshift_r(a[i], b[i], 16); }
            if (MODE == 14) { a[i] = a[i] + b[i] + c[i]; }     // IADD3
            if (MODE == 15) { a[i] = a[i] + b[i] + 3; c[i] = __vimin3_u16x2(c[i], b[i], b[i]+1); }
            if (MODE == 16) { a[i] = __vmaxu2(__vminu2(a[i], b[i]), __vminu2(__vmaxu2(a[i], b[i]), c[i])); }   // med3, 4 x VIMNMX2
            if (MODE == 17) { uint32_t lo = __vimin3_u16x2(a[i], b[i], c[i]), hi = __vimax3_u16x2(a[i], b[i], c[i]);
                              a[i] = (a[i] + b[i] + c[i]) - lo - hi; }                                        // med3, 2 x VIMNMX3 + 2 x IADD3
            if (MODE == 18) { uint32_t lo = __vimin3_u16x2(a[i], b[i], c[i]), hi = __vimax3_u16x2(a[i], b[i], c[i]);
                              uint32_t s = a[i] * one + b[i]; s = c[i] * one + s; s = lo * neg1 + s; a[i] = hi * neg1 + s; }  // 2 x VIMNMX3 + 4 x IMAD
            if (MODE == 19) { a[i] = a[i] * one + b[i]; }                                                      // IMAD alone
            if (MODE == 20) { a[i] = a[i] * one + b[i]; c[i] = __vminu2(c[i], b[i]); }
            if (MODE == 21) { a[i] = __umulhi(a[i], k16) + b[i]; }                                                // IMAD.HI alone
            if (MODE == 22) { c[i] = b[i] * k16 + __umulhi(c[i], k16); a[i] = __vimin3_u16x2(a[i], b[i], b[i] + 1); }   // funnel shift on the FMA pipe + VIMNMX3
            if (MODE == 23) { c[i] = __funnelshift_r(c[i], b[i], 16); a[i] = __vimin3_u16x2(a[i], b[i], b[i] + 1); }       // SHF + VIMNMX3
            if (MODE == 24) { c[i] = b[i] * k16 + __umulhi(c[i], k16); }                                              // IMAD.HI + IMAD
            asm volatile("" : "+r"(a[i]), "+r"(c[i]));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= a[i] ^ c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, int instr_per_op) {
    uint32_t *out;
    const int threads = 1024, blocks = 148 * 2;
    cudaMalloc(&out, blocks * threads * 4);
    bench<MODE><<<blocks, threads>>>(out, 12345u, 1u, 0xffffffffu, 65536u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, threads>>>(out, 12345u, 1u, 0xffffffffu, 65536u);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double warp_instr = (double)blocks * 32 * ITERS * ILP * instr_per_op;
    printf("%-34s %7.3f ms -> %.2f warp-instr/clk/SM (assuming %.0f MHz)\n", name, ms,
           warp_instr / (ms * 1e-3) / 148.0 / (clk_khz * 1e3), clk_khz / 1e3);
    cudaFree(out);
}
int main() {
    run<0>("VIMNMX3.U16x2", 1);
    run<1>("VIMNMX.U16x2 (2-in)", 1);
    run<12>("HMNMX2", 1);
    run<13>("SHF", 1);
    run<14>("IADD3", 1);
    run<2>("VIMNMX3 + VIADD + VIMNMX2", 3);
    run<3>("VIMNMX3 + VIADD + SHF", 3);
    run<4>("VIMNMX2 + SHF", 2);
    run<5>("HMNMX2 + SHF", 2);
    run<6>("HMNMX2 + VIMNMX3 + VIADD", 3);
    run<7>("VIMNMX2 + IMAD", 2);
    run<8>("HMNMX2 + IMAD", 2);
    run<9>("VIMNMX3 + VIADD + IMAD", 3);
    run<10>("HMNMX2 + VIMNMX2", 2);
    run<11>("SHF + IMAD", 2);
    run<15>("IADD + VIMNMX3 + VIADD", 3);
    run<16>("med3 = 4 x VIMNMX2 (per med3)", 1);
    run<17>("med3 = 2 x VIMNMX3 + 2 x IADD3 (per med3)", 1);
    run<18>("med3 = 2 x VIMNMX3 + 4 x IMAD (per med3)", 1);
    run<19>("IMAD", 1);
    run<20>("IMAD + VIMNMX2", 2);
    run<21>("IMAD.HI (+IADD)", 1);
    run<24>("funnel shift as IMAD.HI + IMAD (per shift)", 1);
    run<23>("SHF + VIMNMX3 (per pair)", 1);
    run<22>("IMAD.HI + IMAD + VIMNMX3 (per pair)", 1);
    return 0;
}
