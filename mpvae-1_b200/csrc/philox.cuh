// Philox4x32-10 (Salmon et al., SC'11) + Box-Muller, shared by the noise kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpv {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// Box-Muller on the special-function unit (MUFU lg2 / sqrt / sin / cos).  The outputs are rounded to the fp16 grid
// right after (philox_normal4), a 2.4e-4 relative step, so the ~1e-6 error of the approximate units is invisible
// there; what it buys is a third of the instructions of logf / sqrtf / sincospif.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    // 23 random bits + 0.5 is exact in fp32, so u is never 0 or 1 and has no rounding step
    const float u1 = ((float)(a >> 9) + 0.5f) * (1.0f / 8388608.0f);   // (0, 1)
    const float u2 = ((float)(b >> 9) + 0.5f) * (1.0f / 8388608.0f);
    // r^2 = -2 ln u1 = -2 ln2 * log2 u1; clamped at 0 because lg2.approx may return +1 ulp for u1 -> 1
    const float r2 = fmaxf(-1.3862943611198906f * __log2f(u1), 0.0f);
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    n0 = r * cs;
    n1 = r * sn;
}

// Stream offset: the host value plus an optional device-side counter (CUDA-graph replays bump the counter).
__device__ __forceinline__ uint2 philox_offset(uint2 off, const unsigned long long* off_dev) {
    if (off_dev == nullptr) return off;
    const unsigned long long o = (((unsigned long long)off.y << 32) | off.x) + *off_dev;
    return make_uint2((uint32_t)o, (uint32_t)(o >> 32));
}

// The four standard normals of counter c (flat indices 4c .. 4c+3 of the GLOBAL (S, B_global, Z) tensor).
__device__ __forceinline__ void philox_normal4(unsigned long long c, uint2 key, uint2 off, float (&n)[4]) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), off.x, off.y), key);
    box_muller(r.x, r.y, n[0], n[1]);
    box_muller(r.z, r.w, n[2], n[3]);
    // The library's noise is DEFINED on the fp16 grid (11-bit mantissa): still standard normal to ~2e-4 relative
    // per sample, identical whichever engine consumes it, and exactly representable as a single tensor-core
    // operand piece -- the dense products with it need two MMA passes instead of three (contract_tc.cu).
#pragma unroll
    for (int j = 0; j < 4; ++j) n[j] = __half2float(__float2half_rn(n[j]));
}

}  // namespace mpv
