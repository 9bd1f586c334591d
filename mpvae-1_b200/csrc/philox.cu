// Counter-based standard normal noise: replaces torch.normal(0, 1, (S, B, Z)) of mpvae.py:162.
//
// Philox4x32-10 (Salmon et al., SC'11) keyed by `seed`; counter = (flat_index / 4, offset) where
// flat_index addresses the GLOBAL (S, B_global, Z) tensor, so a rank that owns rows [row0, row0+B)
// draws exactly the numbers a single GPU would draw for those rows.  Four uniforms per counter become
// four normals by two Box-Muller transforms.
#include "common.cuh"
#include "rows.h"

namespace mpv {
namespace {

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

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    // 23 random bits + 0.5 is exact in fp32, so u is never 0 or 1 and has no rounding step
    const float u1 = ((float)(a >> 9) + 0.5f) * (1.0f / 8388608.0f);   // (0, 1)
    const float u2 = ((float)(b >> 9) + 0.5f) * (1.0f / 8388608.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    n0 = r * cs;
    n1 = r * sn;
}

// One thread per Philox counter.  For sample s the local rows form one contiguous span of the global
// flat index space: [ (s*Bg + row0) * Z, (s*Bg + row0 + B) * Z ).
__global__ void __launch_bounds__(256)
philox_normal_kernel(float* __restrict__ noise, int S, int B, int Z, int Bg, int row0, uint2 key, uint2 off) {
    const int s = blockIdx.y;
    const unsigned long long span_beg = ((unsigned long long)s * Bg + row0) * Z;
    const unsigned long long span_end = span_beg + (unsigned long long)B * Z;
    const unsigned long long c_first = span_beg >> 2;
    const unsigned long long c = c_first + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if ((c << 2) >= span_end) return;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), off.x, off.y), key);
    float n[4];
    box_muller(r.x, r.y, n[0], n[1]);
    box_muller(r.z, r.w, n[2], n[3]);
    float* __restrict__ dst = noise + (size_t)s * B * Z;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned long long g = (c << 2) + j;
        if (g >= span_beg && g < span_end) dst[g - span_beg] = n[j];
    }
}

}  // namespace

int launch_philox_normal(float* noise, int S, int B, int Z, int Bg, int row0, uint64_t seed, uint64_t offset,
                         cudaStream_t stream) {
    if (S <= 0 || B <= 0 || Z <= 0) return 0;
    if (S > 65535) { set_error("philox: S=%d exceeds grid.y limit", S); return 6; }
    const unsigned long long counters = ((unsigned long long)B * Z + 3) / 4 + 1;
    dim3 grid((unsigned)((counters + 255) / 256), S);
    philox_normal_kernel<<<grid, 256, 0, stream>>>(noise, S, B, Z, Bg, row0,
                                                   make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
                                                   make_uint2((uint32_t)offset, (uint32_t)(offset >> 32)));
    return check_launch("philox_normal_kernel");
}

}  // namespace mpv
