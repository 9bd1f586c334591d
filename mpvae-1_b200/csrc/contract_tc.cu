// tcgen05 contraction engine (sm_100a): split-precision 3xTF32 GEMMs for the dense regime of the probit ELBO
// (label / rank sets >= 128, e.g. the delicious and eurlex shapes).
//
//   nt: nr[m, l]  = sum_z noise[m, z] * R[l, z]        both operands K-major     (mpvae.py:168)
//   tn: g_R[l, z] = sum_m gxs[m, l]  * noise[m, z]     both operands MN-major    
//
// Plain TF32 (10-bit mantissa) cannot hold the 1e-5 parity bar, so every fp32 operand x is split on the fly
// into hi = tf32(x), lo = tf32(x - hi) (one streaming pre-pass) and the product is accumulated in fp32 in
// tensor memory as  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  (error ~2^-21 relative, below fp32 SGEMM order noise).
//
// The tensor core adds into its fp32 accumulator with truncation, so a K = 3993 chain drifts by ~K * 2^-24
// (measured 3.4e-5 relative).  The accumulation is therefore chunked: tensor memory only ever holds the sum of
// KC k-blocks (128 values of K); the epilogue warps promote each chunk into fp32 REGISTER accumulators with
// round-to-nearest adds while the tensor core fills the other TMEM buffer.
//
// Kernel anatomy (one CTA per SM, persistent over 128 x 256 output tiles, 640 threads):
//   warp 0    : TMA producer  -- cp.async.bulk.tensor (128B-swizzled boxes) into a 2-stage smem ring
//   warp 1    : MMA issuer    -- one lane issues tcgen05.mma.kind::tf32 (M128 N256 K8), 12 per k-block
//   warp 2    : TMEM allocator (512 columns = two 128x256 fp32 chunk accumulators, ping-pong)
//   warps 4-19: promotion + epilogue -- tcgen05.ld a chunk (32 rows x 64 columns per warp), add into registers,
//               store the finished tile
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), tmem full/empty mbarriers (MMA <-> promotion).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc.h"

namespace mpv {
namespace {

constexpr int BM = 128, BN = 256, BK = 32;              // fp32 elements; BK * 4 B = one 128-byte swizzle row
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 4;                     // 16 KiB per (hi | lo) tile
constexpr int B_BYTES = BN * BK * 4;                     // 32 KiB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 96 KiB
constexpr int EPI_BYTES = 0;
constexpr int kDefaultKC = 2;                             // k-blocks accumulated in tensor memory per chunk
// The tensor core truncates (toward zero) when it adds into the fp32 accumulator, which shrinks a chunk sum by
// a measured 1.85e-7 of its value per k-block (12 MMAs of K=8; profiles/r1_tc_chunk_experiment.txt).  The
// promotion multiplies the chunk by (1 + this * k-blocks) to take the systematic part out again; what is left
// is random and ~2x below an fp32 SGEMM's own rounding.
constexpr float kTruncBiasPerKBlock = 1.85e-7f;
constexpr int kEpiWarps = 16;                            // 4 TMEM lane quadrants x 4 column quarters
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
constexpr int TMEM_COLS = 512;
constexpr int kThreads = 128 + 32 * kEpiWarps;

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();   // ~2 s: a protocol bug must not hang the GPU
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor, 128B swizzle (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (sm_100) | [61,64) layout type: 2 = SWIZZLE_128B (16 B atoms), 1 = SWIZZLE_128B_BASE32B
//   (32 B atoms -- the only layout the hardware accepts for MN-major 32-bit operands)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)(sbo16 & 0x3FFFu) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, majors, N >> 3, M >> 4.
template <bool MN>
__device__ __forceinline__ constexpr uint32_t umma_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | (MN ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(BN >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ GEMM
// C[Mc, Nc] (row-major, pitch ldc) = sum_k A(m, k) * B(n, k); operands come pre-split as [2][.][.] fp32
// (plane 0 = hi, plane 1 = lo) through 3-D TMA maps.
//   MN == false: A is [Mc][K], B is [Nc][K] (K contiguous);  box {BK, rows, 1}
//   MN == true : A is [K][Mc], B is [K][Nc] (Mc / Nc contiguous); boxes {32, BK, 1}, 4 per A tile, 8 per B tile
template <bool MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   float* __restrict__ C, int Mc, int Nc, int K, int ldc, int tiles_m, int tiles_n, int kc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;             // 128B-swizzled tiles need 1024 B alignment
    uint8_t* gen = smem_raw + (base - raw);
    const uint32_t bars = base + STAGES * STAGE_BYTES + EPI_BYTES;
    // barrier slots (8 B each): full[STAGES], empty[STAGES], tfull[2], tempty[2]; then the TMEM base address
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + STAGES * STAGE_BYTES + EPI_BYTES + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (K + BK - 1) / BK;

    if (warp == 0) {
        if (lane == 0) {   // ------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t fb = full_bar(stage);
                    mbar_arrive_expect_tx(fb, STAGE_BYTES);
                    const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + 2 * A_BYTES;
                    if (!MN) {
                        tma_load_3d(sa, &tmA, fb, kb * BK, m0, 0);
                        tma_load_3d(sa + A_BYTES, &tmA, fb, kb * BK, m0, 1);
                        tma_load_3d(sb, &tmB, fb, kb * BK, n0, 0);
                        tma_load_3d(sb + B_BYTES, &tmB, fb, kb * BK, n0, 1);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 32; ++j) {
                            tma_load_3d(sa + j * (BK * 128), &tmA, fb, m0 + j * 32, kb * BK, 0);
                            tma_load_3d(sa + A_BYTES + j * (BK * 128), &tmA, fb, m0 + j * 32, kb * BK, 1);
                        }
#pragma unroll
                        for (int j = 0; j < BN / 32; ++j) {
                            tma_load_3d(sb + j * (BK * 128), &tmB, fb, n0 + j * 32, kb * BK, 0);
                            tma_load_3d(sb + B_BYTES + j * (BK * 128), &tmB, fb, n0 + j * 32, kb * BK, 1);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = umma_idesc<MN>();
            // K-major (SWIZZLE_128B): rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1); a k-step of
            //   8 elements is +32 B inside the swizzle row.
            // MN-major (SWIZZLE_128B_BASE32B): each k-row holds 128 B of MN, the next 32-wide MN chunk is one TMA
            //   box (BK*128 B) away (LBO), k-rows come in groups of 4 that are 512 B apart (SBO); a k-step of 8 is
            //   +1024 B.
            constexpr uint32_t lbo = MN ? (BK * 128) / 16 : 1, sbo = MN ? 512 / 16 : 1024 / 16;
            constexpr uint32_t kstep = MN ? 1024u : 32u, layout = MN ? 1u : 2u;
            int stage = 0, buf = 0;
            uint32_t phase = 0, bphase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb0 = 0; kb0 < num_kb; kb0 += kc) {
                    mbar_wait(tempty_bar(buf), bphase ^ 1u);     // promotion warps have drained this TMEM buffer
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(buf * BN);
                    const int kb1 = min(kb0 + kc, num_kb);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + 2 * A_BYTES;
#pragma unroll
                        for (int kk = 0; kk < BK / 8; ++kk) {
                            const uint64_t a_hi = umma_desc(sa + kk * kstep, lbo, sbo, layout);
                            const uint64_t a_lo = umma_desc(sa + A_BYTES + kk * kstep, lbo, sbo, layout);
                            const uint64_t b_hi = umma_desc(sb + kk * kstep, lbo, sbo, layout);
                            const uint64_t b_lo = umma_desc(sb + B_BYTES + kk * kstep, lbo, sbo, layout);
                            tc_mma_tf32(d, a_lo, b_hi, idesc, (kb != kb0 || kk != 0) ? 1u : 0u);   // small terms first
                            tc_mma_tf32(d, a_hi, b_lo, idesc, 1u);
                            tc_mma_tf32(d, a_hi, b_hi, idesc, 1u);
                        }
                        tc_commit(empty_bar(stage));            // smem stage reusable once these MMAs retire
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit(tfull_bar(buf));                   // chunk complete -> promotion warps
                    if (++buf == 2) { buf = 0; bphase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {   // ------------------------------- promotion + epilogue: TMEM lanes 32q.., columns 64h..
        const int q = warp & 3, h = (warp - 4) >> 2;
        float acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.0f;
        int buf = 0;
        uint32_t bphase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int kb0 = 0; kb0 < num_kb; kb0 += kc) {
                mbar_wait(tfull_bar(buf), bphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * 64);
                const float unbias = 1.0f + kTruncBiasPerKBlock * (float)(min(kb0 + kc, num_kb) - kb0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t v[16];
                    tmem_ld_32x16(taddr + i * 16, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j)   // round-to-nearest promotion
                        acc[i * 16 + j] = fmaf(__uint_as_float(v[j]), unbias, acc[i * 16 + j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(buf));
                if (++buf == 2) { buf = 0; bphase ^= 1u; }
            }
            const int row = (tile / tiles_n) * BM + q * 32 + lane;
            const int col0 = (tile % tiles_n) * BN + h * 64;
            if (row < Mc) {
                float* __restrict__ crow = C + (size_t)row * ldc;
#pragma unroll
                for (int j = 0; j < 64; ++j)
                    if (col0 + j < Nc) crow[col0 + j] = acc[j];
            }
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[j] = 0.0f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// hi = tf32(x) (round to nearest, ties away), lo = tf32(x - hi); dst planes are [rows][dpitch], pad columns zero.
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int spitch, int dpitch,
                  size_t plane) {
    const size_t n = (size_t)rows * dpitch;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / dpitch), c = (int)(i % dpitch);
        float hi = 0.0f, lo = 0.0f;
        if (c < cols) {
            const float x = src[(size_t)r * spitch + c];
            uint32_t h, l;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
            hi = __uint_as_float(h);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
            lo = __uint_as_float(l);
        }
        dst[i] = hi;
        dst[plane + i] = lo;
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D map over a split operand [2][rows][pitch] (fp32): dims {cols, rows, 2}, 128B swizzle (16 B atoms for
// K-major tiles, 32 B atoms for MN-major tiles), zero OOB fill.
int make_map(CUtensorMap* map, const float* ptr, int cols, int rows, int pitch, int box_cols, int box_rows,
             CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled unavailable from the driver"); return 8; }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)rows * pitch * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) cols=%d rows=%d pitch=%d", (int)r, cols, rows, pitch); return 8; }
    return 0;
}

int pitch_of(int cols) { return ceil_div(cols, 32) * 32; }

// k-blocks per tensor-memory chunk (MPVAE_TC_KC overrides it for accuracy / speed experiments)
int chunk_kblocks() {
    static int kc = 0;
    if (kc == 0) {
        const char* e = getenv("MPVAE_TC_KC");
        kc = e ? atoi(e) : kDefaultKC;
        if (kc < 1) kc = kDefaultKC;
    }
    return kc;
}

int split(const float* src, float* dst, int rows, int cols, int pitch, cudaStream_t stream) {
    const size_t n = (size_t)rows * pitch;
    const int blocks = (int)((n + 255) / 256 < (size_t)(8 * kNumSMs) ? (n + 255) / 256 : 8 * kNumSMs);
    split_tf32_kernel<<<blocks, 256, 0, stream>>>(src, dst, rows, cols, cols, pitch, n);
    return check_launch("split_tf32_kernel");
}

template <bool MN>
int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, float* C, int Mc, int Nc, int K, int ldc, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        const cudaError_t e = cudaFuncSetAttribute(gemm_3xtf32_kernel<MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(gemm_3xtf32): %s", cudaGetErrorString(e)); return 4; }
        configured = true;
    }
    const int tiles_m = ceil_div(Mc, BM), tiles_n = ceil_div(Nc, BN);
    const int tiles = tiles_m * tiles_n;
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    gemm_3xtf32_kernel<MN><<<grid, kThreads, SMEM_BYTES, stream>>>(a, b, C, Mc, Nc, K, ldc, tiles_m, tiles_n, chunk_kblocks());
    return check_launch("gemm_3xtf32_kernel");
}

}  // namespace

bool tc_available() { return true; }

size_t tc_workspace_nt(int M, int N, int K) {
    const size_t kp = pitch_of(K);
    return align_up(2 * (size_t)M * kp * 4, 1024) + align_up(2 * (size_t)N * kp * 4, 1024);
}

size_t tc_workspace_tn(int M, int N1, int N2) {
    return align_up(2 * (size_t)M * pitch_of(N1) * 4, 1024) + align_up(2 * (size_t)M * pitch_of(N2) * 4, 1024);
}

int tc_contract_nt(const float* A, const float* Bm, float* C, int M, int N, int K, void* ws, size_t ws_bytes,
                   cudaStream_t stream) {
    if (!ws || ws_bytes < tc_workspace_nt(M, N, K)) { set_error("tc_contract_nt: workspace too small"); return 5; }
    const int kp = pitch_of(K);
    float* a2 = static_cast<float*>(ws);
    float* b2 = reinterpret_cast<float*>(static_cast<char*>(ws) + align_up(2 * (size_t)M * kp * 4, 1024));
    if (int rc = split(A, a2, M, K, kp, stream)) return rc;
    if (int rc = split(Bm, b2, N, K, kp, stream)) return rc;
    CUtensorMap ma, mb;
    if (int rc = make_map(&ma, a2, K, M, kp, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    if (int rc = make_map(&mb, b2, K, N, kp, BK, BN, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    return launch_gemm<false>(ma, mb, C, M, N, K, N, stream);
}

int tc_contract_tn(const float* A, const float* Bm, float* C, int M, int N1, int N2, void* ws, size_t ws_bytes,
                   cudaStream_t stream) {
    if (!ws || ws_bytes < tc_workspace_tn(M, N1, N2)) { set_error("tc_contract_tn: workspace too small"); return 5; }
    const int p1 = pitch_of(N1), p2 = pitch_of(N2);
    float* a2 = static_cast<float*>(ws);
    float* b2 = reinterpret_cast<float*>(static_cast<char*>(ws) + align_up(2 * (size_t)M * p1 * 4, 1024));
    if (int rc = split(A, a2, M, N1, p1, stream)) return rc;
    if (int rc = split(Bm, b2, M, N2, p2, stream)) return rc;
    CUtensorMap ma, mb;
    if (int rc = make_map(&ma, a2, N1, M, p1, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    if (int rc = make_map(&mb, b2, N2, M, p2, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
    return launch_gemm<true>(ma, mb, C, N1, N2, M, N2, stream);
}

}  // namespace mpv
