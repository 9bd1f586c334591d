// Counter-based standard normal noise: replaces torch.normal(0, 1, (S, B, Z)) of mpvae.py:162.
//
// Philox4x32-10 (Salmon et al., SC'11) keyed by `seed`; counter = (flat_index / 4, offset) where
// flat_index addresses the GLOBAL (S, B_global, Z) tensor, so a rank that owns rows [row0, row0+B)
// draws exactly the numbers a single GPU would draw for those rows.  Four uniforms per counter become
// four normals by two Box-Muller transforms.
#include "common.cuh"
#include "philox.cuh"
#include "rows.h"

namespace mpv {
namespace {

// One thread per Philox counter.  For sample s the local rows form one contiguous span of the global
// flat index space: [ (s*Bg + row0) * Z, (s*Bg + row0 + B) * Z ).
__global__ void __launch_bounds__(256)
philox_normal_kernel(float* __restrict__ noise, int S, int B, int Z, int Bg, int row0, uint2 key, uint2 off,
                     const unsigned long long* __restrict__ off_dev) {
    const int s = blockIdx.y;
    const unsigned long long span_beg = ((unsigned long long)s * Bg + row0) * Z;
    const unsigned long long span_end = span_beg + (unsigned long long)B * Z;
    const unsigned long long c_first = span_beg >> 2;
    const unsigned long long c = c_first + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if ((c << 2) >= span_end) return;
    float n[4];
    philox_normal4(c, key, philox_offset(off, off_dev), n);
    float* __restrict__ dst = noise + (size_t)s * B * Z;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned long long g = (c << 2) + j;
        if (g >= span_beg && g < span_end) dst[g - span_beg] = n[j];
    }
}

}  // namespace

int launch_philox_normal(float* noise, int S, int B, int Z, int Bg, int row0, uint64_t seed, uint64_t offset,
                         const uint64_t* offset_dev, cudaStream_t stream) {
    if (S <= 0 || B <= 0 || Z <= 0) return 0;
    if (S > 65535) { set_error("philox: S=%d exceeds grid.y limit", S); return 6; }
    const unsigned long long counters = ((unsigned long long)B * Z + 3) / 4 + 1;
    dim3 grid((unsigned)((counters + 255) / 256), S);
    philox_normal_kernel<<<grid, 256, 0, stream>>>(noise, S, B, Z, Bg, row0,
                                                   make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
                                                   make_uint2((uint32_t)offset, (uint32_t)(offset >> 32)),
                                                   reinterpret_cast<const unsigned long long*>(offset_dev));
    return check_launch("philox_normal_kernel");
}

}  // namespace mpv
