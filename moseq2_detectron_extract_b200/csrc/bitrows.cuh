// Bit-row helpers shared by the frame-feature kernels and the session ROI kernel: an image row is kept as 32-bit words,
// lane k of a warp owns word k (rows up to 1024 pixels).
#pragma once
#include <stdint.h>

namespace msq {

__device__ __forceinline__ uint32_t fill_up(uint32_t seed, uint32_t open) {
    return (((open + seed) ^ open) & open) | seed;
}
// all bits of every run of `open` that contains a seed bit, within one 32-bit word
__device__ __forceinline__ uint32_t fill_local(uint32_t seed, uint32_t open) {
    return fill_up(seed, open) | __brev(fill_up(__brev(seed), __brev(open)));
}
__device__ __forceinline__ uint32_t trailing_ones(uint32_t open) {
    return open & ~(open + 1u);
}
__device__ __forceinline__ uint32_t leading_ones(uint32_t open) {
    return __brev(trailing_ones(__brev(open)));
}

// Warp-wide: row = 32 lanes x 32 bits.  Returns, per lane, the bits of the runs of `open` (runs may
// span words) that contain at least one bit of `seed`.
__device__ __forceinline__ uint32_t fill_row(uint32_t seed, uint32_t open, int lane) {
    uint32_t f = fill_local(seed & open, open);
    const uint32_t full = __ballot_sync(0xffffffffu, open == 0xffffffffu);
    const uint32_t g_up = __ballot_sync(0xffffffffu, (f >> 31) != 0u);     // word filled up to its top bit
    const uint32_t g_dn = __ballot_sync(0xffffffffu, (f & 1u) != 0u);      // word filled down to bit 0
    // carry-lookahead with one integer add: generate = g, propagate = word entirely open
    const uint32_t xu = g_up | full;
    const uint32_t cin_up = (xu + g_up) ^ xu ^ g_up;                       // bit k: carry enters word k from k-1
    const uint32_t gr = __brev(g_dn), xr = gr | __brev(full);
    const uint32_t cin_dn = __brev((xr + gr) ^ xr ^ gr);                   // bit k: carry enters word k from k+1
    if ((cin_up >> lane) & 1u) f |= trailing_ones(open);
    if ((cin_dn >> lane) & 1u) f |= leading_ones(open);
    return f;
}

}  // namespace msq
