// tcgen05 contraction engine (sm_100a): split-precision tensor-core GEMMs for the dense regime of the probit ELBO
// (label / rank sets >= 128, e.g. the delicious and eurlex shapes).
//
//   nt: nr[m, l]  = sum_z noise[m, z] * R[l, z]        both operands K-major     (mpvae.py:168)
//   tn: g_R[l, z] = sum_m gxs[m, l]  * noise[m, z]     both operands MN-major    
//
// A 10/11-bit mantissa on the operands cannot hold the 1e-5 parity bar, so every fp32 operand x is split by a
// streaming pre-pass into hi + lo (two 11-bit pieces, ~22 bits together) and the product is accumulated in fp32
// in tensor memory as  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  (the dropped lo.lo term is ~2^-22 relative):
//   kind::tf32  hi = tf32(x),     lo = tf32(x - hi)                       fp32 containers, K = 8 per MMA
//   kind::f16   hi = fp16(x * s), lo = fp16(x * s - hi), s = 2^k chosen per tensor so that max|x * s| is in
//               [1, 2); half the bytes and twice the MMA rate of tf32, result rescaled by 1 / (s_a * s_b)
//
// The tensor core adds into its fp32 accumulator with truncation toward zero, so a long K chain drifts
// (K = 3993: -2.4e-5 relative, measured).  The accumulation is therefore chunked: tensor memory only ever holds
// `kc` k-blocks; sixteen promotion warps add each chunk into fp32 REGISTER accumulators (round-to-nearest) while
// the tensor core fills the other TMEM buffer, and scale the chunk by (1 + bias * k-blocks) to take the systematic
// part of the truncation out again (profiles/r1_tc_chunk_experiment.txt).
//
// Kernel anatomy (one CTA per SM, persistent over 128 x 256 output tiles, 640 threads):
//   warp 0    : TMA producer  -- cp.async.bulk.tensor (128B-swizzled boxes) into a 2-stage, 96 KiB-per-stage smem ring
//   warp 1    : MMA issuer    -- one lane issues tcgen05.mma (M128 N256), 12 per k-block
//   warp 2    : TMEM allocator (512 columns = two 128x256 fp32 chunk accumulators, ping-pong)
//   warps 4-19: promotion + epilogue -- tcgen05.ld a chunk (32 rows x 64 columns per warp), add into registers,
//               store the finished tile
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), tmem full/empty mbarriers (MMA <-> promotion).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "philox.cuh"
#include "tc.h"

namespace mpv {
namespace {

constexpr int BM = 128, BN = 256;
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * 128;                        // 16 KiB per (hi | lo) tile: BM rows x one 128-byte swizzle row
constexpr int B_BYTES = BN * 128;                        // 32 KiB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 96 KiB
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES;
constexpr int TMEM_COLS = 512;
constexpr int kEpiWarps = 16;                            // 4 TMEM lane quadrants x 4 column quarters
constexpr int kThreads = 128 + 32 * kEpiWarps;

// Per-kind geometry.  One k-block is one 128-byte swizzle row of K (K-major) or one TMA box of k-rows (MN-major).
template <bool MN, bool F16>
struct Geo {
    static constexpr int ELT = F16 ? 2 : 4;
    static constexpr int BK = 128 / ELT;                 // K elements per k-block: 32 (tf32) / 64 (f16)
    static constexpr int UK = F16 ? 16 : 8;              // K per tcgen05.mma
    static constexpr int BOX_MN = 128 / ELT;             // MN elements per 128-byte row of an MN-major box
    // K-major: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (= 1); a k-step is +32 B in the row.
    // MN-major: each k-row holds 128 B of MN; the next MN chunk is one TMA box (BK * 128 B) away (LBO);
    //   16-bit: k-rows in groups of 8, 1024 B apart (SWIZZLE_128B);
    //   32-bit: k-rows in groups of 4,  512 B apart (SWIZZLE_128B_BASE32B -- the only layout the hardware takes
    //           for MN-major 32-bit operands; TMA side: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    //   a k-step of UK rows is + UK * 128 B.
    static constexpr uint32_t kstep = MN ? UK * 128u : 32u;
    static constexpr uint32_t lbo = MN ? (BK * 128u) / 16u : 1u;
    static constexpr uint32_t sbo = (MN && !F16) ? 512u / 16u : 1024u / 16u;
    static constexpr uint32_t layout = (MN && !F16) ? 1u : 2u;
    // Instruction descriptor (cute::UMMA::InstrDescriptor): [4,6) D = F32 | [7,10) A fmt | [10,13) B fmt
    // (0 = F16, 2 = TF32) | 15 A major | 16 B major (1 = MN) | [17,23) N >> 3 | [24,29) M >> 4
    static constexpr uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) |
                                      (MN ? ((1u << 15) | (1u << 16)) : 0u) | ((uint32_t)(BN >> 3) << 17) |
                                      ((uint32_t)(BM >> 4) << 24);
    // measured shrink of a TMEM chunk sum per k-block (12 MMAs), see header
    static constexpr float trunc_bias = F16 ? 1.85e-7f : 1.85e-7f;
    // k-blocks per TMEM chunk: every promotion costs 128 KiB of tcgen05.ld per CTA at ~64 B/clk, which stalls the
    // MMA stream; 4 keeps the rms error at 5e-7 (fp32 SGEMM: 1.1e-6 at K = 3993) for half the promotion traffic of 2
    static constexpr int default_kc = F16 ? 4 : 4;
};

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
template <bool F16>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (F16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    }
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

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 (sm_100) | [61,64) layout type: 2 = SWIZZLE_128B (16 B atoms), 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)(sbo16 & 0x3FFFu) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}

// ------------------------------------------------------------------------------------------------ GEMM
// C[Mc, Nc] (row-major, pitch ldc) = inv_scale * sum_k A(m, k) * B(n, k); operands come pre-split as [2][.][.]
// planes (0 = hi, 1 = lo) through 3-D TMA maps.
//   MN == false: A is [Mc][K], B is [Nc][K] (K contiguous);   one box {BK, rows, 1} per tile and plane
//   MN == true : A is [K][Mc], B is [K][Nc] (Mc / Nc contiguous); boxes {BOX_MN, BK, 1}
template <bool MN, bool F16>
__global__ void __launch_bounds__(kThreads, 1)
gemm_split_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  float* __restrict__ C, int Mc, int Nc, int K, int ldc, int tiles_m, int tiles_n, int kc,
                  const uint32_t* __restrict__ absmax_a, const uint32_t* __restrict__ absmax_b, int dbg) {
    using G = Geo<MN, F16>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;             // 128B-swizzled tiles need 1024 B alignment
    uint8_t* gen = smem_raw + (base - raw);
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    // barrier slots (8 B each): full[STAGES], empty[STAGES], tfull[2], tempty[2]; then the TMEM base address
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + STAGES * STAGE_BYTES + 128);

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
    const int num_kb = (K + G::BK - 1) / G::BK;

    if (warp == 0) {
        if (lane == 0) {   // ------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t fb = full_bar(stage);
                    if (dbg == 1) { mbar_arrive(fb); if (++stage == STAGES) { stage = 0; phase ^= 1u; } continue; }   // timing probe: no loads
                    mbar_arrive_expect_tx(fb, STAGE_BYTES);
                    const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + 2 * A_BYTES;
                    if (!MN) {
                        tma_load_3d(sa, &tmA, fb, kb * G::BK, m0, 0);
                        tma_load_3d(sa + A_BYTES, &tmA, fb, kb * G::BK, m0, 1);
                        tma_load_3d(sb, &tmB, fb, kb * G::BK, n0, 0);
                        tma_load_3d(sb + B_BYTES, &tmB, fb, kb * G::BK, n0, 1);
                    } else {
                        constexpr int box = G::BK * 128;
#pragma unroll
                        for (int j = 0; j < BM / G::BOX_MN; ++j) {
                            tma_load_3d(sa + j * box, &tmA, fb, m0 + j * G::BOX_MN, kb * G::BK, 0);
                            tma_load_3d(sa + A_BYTES + j * box, &tmA, fb, m0 + j * G::BOX_MN, kb * G::BK, 1);
                        }
#pragma unroll
                        for (int j = 0; j < BN / G::BOX_MN; ++j) {
                            tma_load_3d(sb + j * box, &tmB, fb, n0 + j * G::BOX_MN, kb * G::BK, 0);
                            tma_load_3d(sb + B_BYTES + j * box, &tmB, fb, n0 + j * G::BOX_MN, kb * G::BK, 1);
                        }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ------------------------------------------------ MMA issuer
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
                        for (int kk = 0; kk < G::BK / G::UK; ++kk) {
                            if (dbg == 2) break;                                                       // timing probe: no MMAs
                            const uint64_t a_hi = umma_desc(sa + kk * G::kstep, G::lbo, G::sbo, G::layout);
                            const uint64_t a_lo = umma_desc(sa + A_BYTES + kk * G::kstep, G::lbo, G::sbo, G::layout);
                            const uint64_t b_hi = umma_desc(sb + kk * G::kstep, G::lbo, G::sbo, G::layout);
                            const uint64_t b_lo = umma_desc(sb + B_BYTES + kk * G::kstep, G::lbo, G::sbo, G::layout);
                            tc_mma<F16>(d, a_lo, b_hi, G::idesc, (kb != kb0 || kk != 0) ? 1u : 0u);   // small terms first
                            tc_mma<F16>(d, a_hi, b_lo, G::idesc, 1u);
                            tc_mma<F16>(d, a_hi, b_hi, G::idesc, 1u);
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
        float inv_scale = 1.0f;
        if (F16) {
            const float sa = absmax_a ? scale_from_absmax_bits(*absmax_a) : 1.0f;
            const float sb = absmax_b ? scale_from_absmax_bits(*absmax_b) : 1.0f;
            inv_scale = 1.0f / (sa * sb);                        // powers of two: exact
        }
        const bool vec_store = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15u) == 0;
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
                const float unbias = 1.0f + G::trunc_bias * (float)(min(kb0 + kc, num_kb) - kb0);
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
                if (vec_store && col0 + 64 <= ldc) {
                    // 16-byte stores: rows are 16 B aligned (ldc % 4 == 0); columns in [Nc, ldc) are pitch padding
#pragma unroll
                    for (int j = 0; j < 64; j += 4)
                        *reinterpret_cast<float4*>(crow + col0 + j) =
                            make_float4(acc[j] * inv_scale, acc[j + 1] * inv_scale, acc[j + 2] * inv_scale, acc[j + 3] * inv_scale);
                } else {
#pragma unroll
                    for (int j = 0; j < 64; ++j)
                        if (col0 + j < Nc) crow[col0 + j] = acc[j] * inv_scale;
                }
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

// ------------------------------------------------------------------------------------------------ 2-SM GEMM
// Same computation with CTA pairs (cta_group::2): a cluster of two CTAs on neighbouring SMs owns a 256 x 256 tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 columns), so a stage is 64 KiB per
// SM instead of 96 (3 stages instead of 2, and a third less L2->smem traffic per MMA cycle -- the 1-SM kernel's
// tensor pipe was only 63 % busy waiting for operands, profiles/r1_gemm_ncu_summary.txt).  The leader CTA's single
// thread issues tcgen05.mma.cta_group::2 (M256 N256): the hardware reads A and B from both CTAs' shared memory and
// writes each CTA's 128 rows of D into its own tensor memory.  Barriers:
//   full[s]   (leader only)  armed by the leader with the bytes of BOTH CTAs; both CTAs' TMA loads complete_tx on it
//   empty[s]  (per CTA)      tcgen05.commit multicast to both CTAs
//   tfull[b]  (per CTA)      tcgen05.commit multicast to both CTAs
//   tempty[b] (leader only)  one arrive per promotion warp of both CTAs (remote arrive from the peer)
//
// EX names an operand that is EXACTLY representable in the 11-bit piece format (no lo plane): 1 = A, 2 = B.
// The library's own Philox noise is drawn on the fp16 grid, so the forward (A = noise) and the backward
// (B = noise) products need only two MMA passes and three 16 KiB tiles per stage (48 KiB, 4-deep ring).
template <int EX>
struct Ring2 {
    static constexpr int NA = (EX == 1) ? 1 : 2, NB = (EX == 2) ? 1 : 2;       // planes staged per operand
    static constexpr int STAGE_BYTES = (NA + NB) * A_BYTES;                       // 64 KiB (EX = 0) or 48 KiB
    static constexpr int STAGES = (EX == 0) ? 3 : 4;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES;
    // measured truncation bias per k-block relative to the 3-pass constant (profiles/r01_exact_operand.md): 1.30e-7 / 1.85e-7
    static constexpr float BIAS_SCALE = (EX == 0) ? 1.0f : 0.70f;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// same box delivered to every CTA of `mask` (same CTA-relative smem offset); each destination's bytes are counted
// on the barrier of ITS pair's leader (the peer bit of `leader_bar` is 0)
__device__ __forceinline__ void tma_load_3d_2sm_mc(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2,
                                                   uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(mask)
        : "memory");
}
template <bool F16>
__device__ __forceinline__ void tc_mma_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (F16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    }
}

//
// CL = 2 puts TWO pairs in one cluster of four CTAs.  The pairs work on neighbouring tiles that share the operand
// which costs two planes (B = R for A-exact / plain products: tiles stacked along M; A = gxs for the B-exact
// backward: tiles side by side along N).  Each CTA fetches only half of its share of that operand and multicasts
// it to the CTA of the same parity in the other pair, so L2 -> shared-memory traffic per pair drops from 96 to
// 64 KiB per k-block.  (Opt-in, MPVAE_TC_CLUSTER=2: it did not pay on B200, see cluster_pairs().)
//   empty[s] then counts one tcgen05.commit per PAIR (multicast to all four CTAs): a stage is rewritten only
//   after both pairs have consumed it, because either pair's producers write into both pairs' shared memory.
// A cluster whose second tile falls outside the matrix runs it on zero-filled boxes and stores nothing.
// One thread's 64 consecutive output columns.  Rows whose start is not 16-byte aligned (odd ldc: g_R, the 512 -> L
// heads) get LEAD scalar stores up to the next 16-byte boundary, 15 vector stores, and the rest as scalars; every
// register index is a compile-time constant.
template <int LEAD>
__device__ __forceinline__ void store_row64(float* __restrict__ dst, const float (&acc)[64], float k) {
#pragma unroll
    for (int j = 0; j < LEAD; ++j) dst[j] = acc[j] * k;
    constexpr int NV = (LEAD == 0) ? 16 : 15;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int j = LEAD + 4 * v;
        *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j] * k, acc[j + 1] * k, acc[j + 2] * k, acc[j + 3] * k);
    }
#pragma unroll
    for (int j = LEAD + 4 * NV; j < 64; ++j) dst[j] = acc[j] * k;
}

template <bool MN, bool F16, int EX, int CL>
__global__ void __launch_bounds__(kThreads, 1)
gemm_split_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      float* __restrict__ C, int Mc, int Nc, int K, int ldc, int tiles_m, int tiles_n, int kc,
                      const uint32_t* __restrict__ absmax_a, const uint32_t* __restrict__ absmax_b, int dbg,
                      int full_tiles, int ksplit, float* __restrict__ partials) {
    using G = Geo<MN, F16>;
    using R2 = Ring2<EX>;
    constexpr int STAGES2 = R2::STAGES, STAGE2_BYTES = R2::STAGE_BYTES;
    constexpr int HB = BN / 2;                                // B-tile columns staged by each CTA
    // M = 256 across the pair: same descriptor fields as Geo::idesc with the M field set to 256 >> 4
    constexpr uint32_t idesc2 = (G::idesc & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw);
    const uint32_t bars = base + STAGES2 * STAGE2_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES2 + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES2 + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES2 + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gen + STAGES2 * STAGE2_BYTES + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t rank = crank & 1u;                         // position inside the pair (0 = leader)
    const uint32_t cpair = crank >> 1;                        // which pair of the cluster
    const uint32_t leader_rank = crank & ~1u;                 // cluster rank of this pair's leader
    const bool leader = rank == 0;
    const int cluster = blockIdx.x / (2 * CL), num_clusters = gridDim.x / (2 * CL);
    constexpr bool SHARE_A = (EX == 2);                       // operand multicast between the pairs (CL == 2)
    // super-tiles: CL neighbouring pair tiles along M (B shared) or along N (A shared)
    const int sup_m = SHARE_A ? tiles_m : (tiles_m + CL - 1) / CL;
    const int sup_n = SHARE_A ? (tiles_n + CL - 1) / CL : tiles_n;
    const int num_super = sup_m * sup_n;
    auto tile_m_of = [&](int st) { return SHARE_A ? st / sup_n : (st / sup_n) * CL + (int)cpair; };
    auto tile_n_of = [&](int st) { return SHARE_A ? (st % sup_n) * CL + (int)cpair : st % sup_n; };
    // Work items.  Tiles [0, full_tiles) fill whole waves of the persistent grid and run over all of K.  Each tile of
    // the last, partial wave is cut into `ksplit` slices of K so that the wave keeps every pair busy for 1/ksplit of
    // a tile time instead of leaving most of them idle for a whole one: slice 0 writes C, slice j > 0 writes a
    // 256 x 256 scratch tile that tail_fixup_kernel adds to C afterwards (fixed order -> reproducible).
    const int num_items = (ksplit > 1) ? full_tiles + (num_super - full_tiles) * ksplit : num_super;
    struct Item { int st, kb_lo, kb_hi, slice; };
    auto item_of = [&](int w, int num_kb) {
        Item it;
        if (ksplit <= 1 || w < full_tiles) { it.st = w; it.kb_lo = 0; it.kb_hi = num_kb; it.slice = 0; return it; }
        const int r = w - full_tiles;
        it.st = full_tiles + r / ksplit;
        it.slice = r % ksplit;
        it.kb_lo = (int)(((long long)num_kb * it.slice) / ksplit);
        it.kb_hi = (int)(((long long)num_kb * (it.slice + 1)) / ksplit);
        return it;
    };
    const uint16_t mask_all = (uint16_t)((1u << (2 * CL)) - 1u);
    const uint16_t mask_pair = (uint16_t)(3u << (2 * cpair));
    const uint16_t mask_share = (uint16_t)((1u << rank) | (1u << (rank + 2)));   // same parity in both pairs
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES2; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                       // peer barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_kb = (K + G::BK - 1) / G::BK;

    if (warp == 0) {
        if (lane == 0) {   // ------------------------------------------------ TMA producer (every CTA)
            int stage = 0;
            uint32_t phase = 0;
            for (int w = cluster; w < num_items; w += num_clusters) {
                const Item it = item_of(w, num_kb);
                const int st = it.st;
                const int m0 = tile_m_of(st) * 256 + (int)rank * BM;          // this CTA's 128 rows of A
                const int n0 = tile_n_of(st) * BN + (int)rank * HB;           // this CTA's half of the B tile
                for (int kb = it.kb_lo; kb < it.kb_hi; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t fb = map_to_cta(full_bar(stage), leader_rank);   // the pair leader's barrier collects both CTAs' bytes
                    if (dbg == 1) {   // timing probe: no loads
                        if (leader) mbar_arrive(full_bar(stage));
                        if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE2_BYTES);
                    const uint32_t sa = base + stage * STAGE2_BYTES, sb = sa + R2::NA * A_BYTES;
                    if (!MN) {
                        // K-major tiles are [rows][128 B]; the shared operand's map has boxes of rows / CL
                        constexpr int ra = SHARE_A ? BM / CL : BM, rb = SHARE_A ? HB : HB / CL;
#pragma unroll
                        for (int pl = 0; pl < R2::NA; ++pl) {
                            if (CL > 1 && SHARE_A)
                                tma_load_3d_2sm_mc(sa + pl * A_BYTES + cpair * ra * 128, &tmA, fb, kb * G::BK, m0 + cpair * ra, pl, mask_share);
                            else
                                tma_load_3d_2sm(sa + pl * A_BYTES, &tmA, fb, kb * G::BK, m0, pl);
                        }
#pragma unroll
                        for (int pl = 0; pl < R2::NB; ++pl) {
                            if (CL > 1 && !SHARE_A)
                                tma_load_3d_2sm_mc(sb + pl * A_BYTES + cpair * rb * 128, &tmB, fb, kb * G::BK, n0 + cpair * rb, pl, mask_share);
                            else
                                tma_load_3d_2sm(sb + pl * A_BYTES, &tmB, fb, kb * G::BK, n0, pl);
                        }
                    } else {
                        // MN-major tiles are BOX_MN-wide column boxes of BK k-rows; the shared operand's boxes are
                        // dealt out between the pairs
                        constexpr int box = G::BK * 128;
                        constexpr int nba = BM / G::BOX_MN, nbb = HB / G::BOX_MN;
#pragma unroll
                        for (int j = 0; j < nba; ++j) {
                            if (CL > 1 && SHARE_A && (j / (nba / CL)) != (int)cpair) continue;
#pragma unroll
                            for (int pl = 0; pl < R2::NA; ++pl) {
                                if (CL > 1 && SHARE_A)
                                    tma_load_3d_2sm_mc(sa + pl * A_BYTES + j * box, &tmA, fb, m0 + j * G::BOX_MN, kb * G::BK, pl, mask_share);
                                else
                                    tma_load_3d_2sm(sa + pl * A_BYTES + j * box, &tmA, fb, m0 + j * G::BOX_MN, kb * G::BK, pl);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < nbb; ++j) {
                            if (CL > 1 && !SHARE_A && (j / (nbb / CL)) != (int)cpair) continue;
#pragma unroll
                            for (int pl = 0; pl < R2::NB; ++pl) {
                                if (CL > 1 && !SHARE_A)
                                    tma_load_3d_2sm_mc(sb + pl * A_BYTES + j * box, &tmB, fb, n0 + j * G::BOX_MN, kb * G::BK, pl, mask_share);
                                else
                                    tma_load_3d_2sm(sb + pl * A_BYTES + j * box, &tmB, fb, n0 + j * G::BOX_MN, kb * G::BK, pl);
                            }
                        }
                    }
                    if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {   // ------------------------------------- MMA issuer (leader CTA only)
            int stage = 0, buf = 0;
            uint32_t phase = 0, bphase = 0;
            for (int w = cluster; w < num_items; w += num_clusters) {
                const Item it = item_of(w, num_kb);
                for (int kb0 = it.kb_lo; kb0 < it.kb_hi; kb0 += kc) {
                    mbar_wait(tempty_bar(buf), bphase ^ 1u);     // both CTAs' promotion warps have drained this buffer
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(buf * BN);
                    const int kb1 = min(kb0 + kc, it.kb_hi);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = base + stage * STAGE2_BYTES, sb = sa + R2::NA * A_BYTES;
#pragma unroll
                        for (int kk = 0; kk < G::BK / G::UK; ++kk) {
                            if (dbg == 2) break;   // timing probe: no MMAs
                            const uint32_t first = (kb != kb0 || kk != 0) ? 1u : 0u;
                            const uint64_t a_hi = umma_desc(sa + kk * G::kstep, G::lbo, G::sbo, G::layout);
                            const uint64_t b_hi = umma_desc(sb + kk * G::kstep, G::lbo, G::sbo, G::layout);
                            if (EX == 1) {          // A exact: A.B_lo + A.B_hi
                                const uint64_t b_lo = umma_desc(sb + A_BYTES + kk * G::kstep, G::lbo, G::sbo, G::layout);
                                tc_mma_2sm<F16>(d, a_hi, b_lo, idesc2, first);
                            } else if (EX == 2) {   // B exact: A_lo.B + A_hi.B
                                const uint64_t a_lo = umma_desc(sa + A_BYTES + kk * G::kstep, G::lbo, G::sbo, G::layout);
                                tc_mma_2sm<F16>(d, a_lo, b_hi, idesc2, first);
                            } else {                // small terms first
                                const uint64_t a_lo = umma_desc(sa + A_BYTES + kk * G::kstep, G::lbo, G::sbo, G::layout);
                                const uint64_t b_lo = umma_desc(sb + A_BYTES + kk * G::kstep, G::lbo, G::sbo, G::layout);
                                tc_mma_2sm<F16>(d, a_lo, b_hi, idesc2, first);
                                tc_mma_2sm<F16>(d, a_hi, b_lo, idesc2, 1u);
                            }
                            tc_mma_2sm<F16>(d, a_hi, b_hi, idesc2, 1u);
                        }
                        tc_commit_2sm(empty_bar(stage), mask_all);   // this pair is done with the stage: tell every CTA of the cluster
                        if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit_2sm(tfull_bar(buf), mask_pair);    // chunk complete in both CTAs' tensor memory
                    if (++buf == 2) { buf = 0; bphase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {   // ------------------------------- promotion + epilogue (both CTAs, own 128 rows)
        const int q = warp & 3, h = (warp - 4) >> 2;
        float inv_scale = 1.0f;
        if (F16) {
            const float sa = absmax_a ? scale_from_absmax_bits(*absmax_a) : 1.0f;
            const float sb = absmax_b ? scale_from_absmax_bits(*absmax_b) : 1.0f;
            inv_scale = 1.0f / (sa * sb);
        }
        const bool vec_store = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15u) == 0;
        float acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.0f;
        int buf = 0;
        uint32_t bphase = 0;
        for (int w = cluster; w < num_items; w += num_clusters) {
            const Item it = item_of(w, num_kb);
            const int st = it.st;
            for (int kb0 = it.kb_lo; kb0 < it.kb_hi; kb0 += kc) {
                mbar_wait(tfull_bar(buf), bphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + h * 64);
                const float unbias = 1.0f + G::trunc_bias * R2::BIAS_SCALE * (float)(min(kb0 + kc, it.kb_hi) - kb0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t v[16];
                    tmem_ld_32x16(taddr + i * 16, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        acc[i * 16 + j] = fmaf(__uint_as_float(v[j]), unbias, acc[i * 16 + j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty_bar(buf), leader_rank));
                if (++buf == 2) { buf = 0; bphase ^= 1u; }
            }
            const int row = tile_m_of(st) * 256 + (int)rank * BM + q * 32 + lane;
            const int col0 = tile_n_of(st) * BN + h * 64;
            if (it.slice > 0) {
                // K-slice of a tail tile: the whole 256 x 256 scratch tile is written (rows / columns beyond the
                // matrix hold sums of zero-filled boxes, i.e. zeros)
                float* __restrict__ prow = partials + ((size_t)(st - full_tiles) * (ksplit - 1) + (it.slice - 1)) * (256 * BN) +
                                           (size_t)((int)rank * BM + q * 32 + lane) * BN + h * 64;
#pragma unroll
                for (int j = 0; j < 64; j += 4)
                    *reinterpret_cast<float4*>(prow + j) =
                        make_float4(acc[j] * inv_scale, acc[j + 1] * inv_scale, acc[j + 2] * inv_scale, acc[j + 3] * inv_scale);
            } else if (row < Mc && col0 < Nc) {
                float* __restrict__ crow = C + (size_t)row * ldc;
                if (vec_store && col0 + 64 <= ldc) {
                    // 16-byte stores: rows are 16 B aligned (ldc % 4 == 0); columns in [Nc, ldc) are pitch padding
#pragma unroll
                    for (int j = 0; j < 64; j += 4)
                        *reinterpret_cast<float4*>(crow + col0 + j) =
                            make_float4(acc[j] * inv_scale, acc[j + 1] * inv_scale, acc[j + 2] * inv_scale, acc[j + 3] * inv_scale);
                } else if (col0 + 64 <= Nc) {
                    // unaligned row: scalars up to the next 16-byte boundary, then vectors
                    float* __restrict__ d = crow + col0;
                    switch ((4 - (int)((reinterpret_cast<uintptr_t>(d) >> 2) & 3u)) & 3) {
                        case 0: store_row64<0>(d, acc, inv_scale); break;
                        case 1: store_row64<1>(d, acc, inv_scale); break;
                        case 2: store_row64<2>(d, acc, inv_scale); break;
                        default: store_row64<3>(d, acc, inv_scale); break;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 64; ++j)
                        if (col0 + j < Nc) crow[col0 + j] = acc[j] * inv_scale;
                }
            }
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[j] = 0.0f;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                       // the peer may still be reading our smem / signalling us
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// C += the K-slices 1 .. ksplit-1 of the tail tiles (see Item above), slice order fixed.  64 CTAs per tail tile,
// each 4 rows of 256 columns; the loads of a thread's four rows are independent of each other.
__global__ void __launch_bounds__(256)
tail_fixup_kernel(float* __restrict__ C, int Mc, int Nc, int ldc, int tiles_n, int full_tiles, int ksplit,
                  const float* __restrict__ partials) {
    const int t = blockIdx.x >> 6, rows0 = (blockIdx.x & 63) * 4;
    const int st = full_tiles + t;
    const int m0 = (st / tiles_n) * 256, n0 = (st % tiles_n) * BN;
    const float* __restrict__ p = partials + (size_t)t * (ksplit - 1) * (256 * BN);
    const int c = threadIdx.x;                      // BN == 256 columns, one per thread
    if (n0 + c >= Nc) return;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (m0 + rows0 + i < Mc) ? C[(size_t)(m0 + rows0 + i) * ldc + n0 + c] : 0.0f;
    for (int j = 0; j < ksplit - 1; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += p[(size_t)j * (256 * BN) + (rows0 + i) * BN + c];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (m0 + rows0 + i < Mc) C[(size_t)(m0 + rows0 + i) * ldc + n0 + c] = v[i];
}

// ------------------------------------------------------------------------------------------------ pre-passes
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

// max |x| as raw fp32 bits (ordering of non-negative floats == ordering of their bit patterns; NaN sorts highest)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ src, size_t n, uint32_t* __restrict__ out) {
    uint32_t m = 0;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {      // 16-byte loads over the aligned bulk
        const size_t n4 = n / 4;
        const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src);
        for (size_t i = tid; i < n4; i += nthr) {
            const uint4 v = s4[i];
            const uint32_t a = max(max(v.x & 0x7FFFFFFFu, v.y & 0x7FFFFFFFu), max(v.z & 0x7FFFFFFFu, v.w & 0x7FFFFFFFu));
            m = a > m ? a : m;
        }
        done = n4 * 4;
    }
    for (size_t i = done + tid; i < n; i += nthr) {
        const uint32_t b = __float_as_uint(src[i]) & 0x7FFFFFFFu;
        m = b > m ? b : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    // one atomic per CTA: thousands of atomics on one address serialise (10 us for a 5 MB tensor with one per warp)
    __shared__ uint32_t s_m[8];
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = s_m[w] > m ? s_m[w] : m;
        if (m != 0) atomicMax(out, m);
    }
}

// hi = fp16(x * s), lo = fp16(x * s - hi) with s from the tensor's absmax; four elements per thread.
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int rows, int cols, int spitch, int dpitch,
                 size_t plane, const uint32_t* __restrict__ absmax) {
    const float s = absmax ? scale_from_absmax_bits(*absmax) : 1.0f;
    const size_t n4 = (size_t)rows * dpitch / 4;          // dpitch is a multiple of 64: groups of 4 never straddle rows
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = 4 * i;
        const int r = (int)(e / dpitch), c = (int)(e % dpitch);
        const float* __restrict__ sp = src + (size_t)r * spitch + c;
        float x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = (c + j < cols) ? sp[j] * s : 0.0f;
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { h[j] = __float2half_rn(x[j]); l[j] = __float2half_rn(x[j] - __half2float(h[j])); }
        const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
        const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
        lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(dst + e) = hv;
        *reinterpret_cast<uint2*>(dst + plane + e) = lv;
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

// 3-D map over a split operand [2][rows][pitch]: dims {cols, rows, 2}, 128B swizzle, zero OOB fill.
int make_map(CUtensorMap* map, const void* ptr, bool f16, int cols, int rows, int pitch, int box_cols, int box_rows,
             CUtensorMapSwizzle swizzle, int planes = 2) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled unavailable from the driver"); return 8; }
    const cuuint64_t esz = f16 ? 2 : 4;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * esz, (cuuint64_t)rows * pitch * esz};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                          const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) cols=%d rows=%d pitch=%d", (int)r, cols, rows, pitch); return 8; }
    return 0;
}

int pitch_of(int cols) { return ceil_div(cols, 64) * 64; }

// MPVAE_TC_KIND = tf32 | f16 selects the operand kind (default below); MPVAE_TC_KC the TMEM chunk length.
bool use_f16() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MPVAE_TC_KIND");
        v = (e && strcmp(e, "tf32") == 0) ? 0 : 1;   // default: fp16 split (same accuracy, half the bytes, 2x MMA rate)
    }
    return v == 1;
}
int chunk_kblocks(int dflt) {
    static int kc = -1;
    if (kc < 0) {
        const char* e = getenv("MPVAE_TC_KC");
        kc = e ? atoi(e) : 0;
    }
    return kc > 0 ? kc : dflt;
}

int grid_for(size_t n) { return (int)((n + 255) / 256 < (size_t)(8 * kNumSMs) ? (n + 255) / 256 : 8 * kNumSMs); }

// scratch layout: [absmax_a, absmax_b (256 B)] [A planes] [B planes]
struct Scratch { uint32_t* absmax; char* a; char* b; };

size_t planes_bytes(size_t rows, size_t pitch) { return align_up(2 * rows * pitch * 4, 1024); }   // sized for fp32 planes

Scratch carve_scratch(void* ws, size_t rows_a, size_t pitch_a) {
    char* p = static_cast<char*>(ws);
    return {reinterpret_cast<uint32_t*>(p), p + 256, p + 256 + planes_bytes(rows_a, pitch_a)};
}

int split_operand(const float* src, void* dst, int rows, int cols, int pitch, bool f16, uint32_t* absmax, cudaStream_t stream) {
    const size_t n = (size_t)rows * pitch;
    if (!f16) {
        split_tf32_kernel<<<grid_for(n), 256, 0, stream>>>(src, static_cast<float*>(dst), rows, cols, cols, pitch, n);
        return check_launch("split_tf32_kernel");
    }
    if (absmax) {
        absmax_kernel<<<grid_for((size_t)rows * cols), 256, 0, stream>>>(src, (size_t)rows * cols, absmax);
        if (int rc = check_launch("absmax_kernel")) return rc;
    }
    split_f16_kernel<<<grid_for(n / 4), 256, 0, stream>>>(src, static_cast<__half*>(dst), rows, cols, cols, pitch, n, absmax);
    return check_launch("split_f16_kernel");
}

// MPVAE_TC_CTA = 1 | 2: CTA pairs (cta_group::2, 256 x 256 tiles) or single CTAs (128 x 256 tiles)
int cta_group() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MPVAE_TC_CTA");
        v = (e && atoi(e) == 1) ? 1 : 2;   // default: CTA pairs
    }
    return v;
}

// MPVAE_TC_CLUSTER = 1 | 2: pairs per cluster.  2 = four-CTA clusters with the two-plane operand multicast; kept as
// an experiment, NOT the default: measured on B200 (profiles/r01_multicast_probe.txt) the load-only time does not
// move (0.49 -> 0.51 ms: the cap is on bytes DELIVERED to the SMs, ~8 TB/s, not on L2 reads) and only 33 of 37
// four-CTA clusters are co-resident (GPC sizes), so the kernel gets slower (0.645 -> 0.666 ms).
int cluster_pairs() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MPVAE_TC_CLUSTER");
        v = (e && atoi(e) == 2) ? 2 : 1;
    }
    return v;
}

// rows of one TMA box of the K-major B operand (R in noise.R^T): the whole tile for single CTAs, half of it per CTA
// of a pair, a quarter when two pairs share the tile by multicast (the nt products never have an exact B)
int nt_b_box_rows() { return cta_group() == 2 ? BN / 2 / cluster_pairs() : BN; }

template <bool MN, bool F16, int EX, int CL>
int launch_gemm_2sm(const CUtensorMap& a, const CUtensorMap& b, float* C, int Mc, int Nc, int K, int ldc, const uint32_t* ma,
                    const uint32_t* mb, int kc, int dbg, cudaStream_t stream, float* partials, size_t partials_bytes) {
    auto kernel = gemm_split_2sm_kernel<MN, F16, EX, CL>;
    static int max_clusters = 0;     // co-resident clusters (a cluster must fit inside one GPC)
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Ring2<EX>::SMEM_BYTES;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters == 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring2<EX>::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(gemm_split_2sm): %s", cudaGetErrorString(e)); return 4; }
        cfg.gridDim = dim3(kNumSMs / (2 * CL) * (2 * CL));
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
        if (e != cudaSuccess || n <= 0) { set_error("cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(e)); return 4; }
        max_clusters = n < kNumSMs / (2 * CL) ? n : kNumSMs / (2 * CL);
        if (dbg) fprintf(stderr, "[mpvae tc] cluster of %d CTAs: %d co-resident (of %d)\n", 2 * CL, n, kNumSMs / (2 * CL));
    }
    const int tiles_m = ceil_div(Mc, 256), tiles_n = ceil_div(Nc, BN);
    const int supers = (EX == 2) ? tiles_m * ceil_div(tiles_n, CL) : ceil_div(tiles_m, CL) * tiles_n;
    // K-split of the last, partial wave (plain pairs only): worth it when the wave would leave at least half of the
    // pairs idle and the caller gave scratch for the slices
    int full_tiles = supers, ksplit = 1;
    static const bool no_split = getenv("MPVAE_TC_NO_KSPLIT") != nullptr;
    if (CL == 1 && partials != nullptr && !no_split) {
        const int tail = supers % max_clusters, num_kb = ceil_div(K, Geo<MN, F16>::BK);
        if (tail > 0) {
            int f = max_clusters / tail;
            if (f > num_kb / kc) f = num_kb / kc;                      // a slice keeps at least one full TMEM chunk
            if (f > 8) f = 8;
            if (f >= 2 && (size_t)tail * (f - 1) * 256 * BN * sizeof(float) <= partials_bytes) { full_tiles = supers - tail; ksplit = f; }
        }
    }
    const int items = full_tiles + (supers - full_tiles) * ksplit;
    const int clusters = items < max_clusters ? items : max_clusters;
    cfg.gridDim = dim3(2 * CL * clusters);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a, b, C, Mc, Nc, K, ldc, tiles_m, tiles_n, kc, ma, mb, dbg, full_tiles, ksplit,
                                             partials);
    if (e != cudaSuccess) { set_error("gemm_split_2sm_kernel launch: %s", cudaGetErrorString(e)); return 3; }
    if (int rc = check_launch("gemm_split_2sm_kernel")) return rc;
    if (ksplit > 1 && dbg == 0) {
        static_assert(BN == 256, "tail_fixup_kernel maps one thread to one tile column");
        tail_fixup_kernel<<<(supers - full_tiles) * 64, 256, 0, stream>>>(C, Mc, Nc, ldc, tiles_n, full_tiles, ksplit, partials);
        return check_launch("tail_fixup_kernel");
    }
    return 0;
}

template <bool MN, bool F16, int EX>
int launch_gemm_2sm(const CUtensorMap& a, const CUtensorMap& b, float* C, int Mc, int Nc, int K, int ldc, const uint32_t* ma,
                    const uint32_t* mb, int kc, int dbg, cudaStream_t stream, float* partials, size_t partials_bytes) {
    if (cluster_pairs() == 2)
        return launch_gemm_2sm<MN, F16, EX, 2>(a, b, C, Mc, Nc, K, ldc, ma, mb, kc, dbg, stream, partials, partials_bytes);
    return launch_gemm_2sm<MN, F16, EX, 1>(a, b, C, Mc, Nc, K, ldc, ma, mb, kc, dbg, stream, partials, partials_bytes);
}

// ex: 0 = both operands carry hi|lo planes, 1 = A is exact (single plane), 2 = B is exact.  Exact operands need the
// CTA-pair kernel (tc_exact_supported()).
template <bool MN, bool F16>
int launch_gemm(const CUtensorMap& a, const CUtensorMap& b, float* C, int Mc, int Nc, int K, int ldc, const uint32_t* ma,
                const uint32_t* mb, cudaStream_t stream, int ex = 0, float* partials = nullptr, size_t partials_bytes = 0) {
    // k-blocks per TMEM chunk: the same number of truncating MMAs per chunk (48) whether a k-step is 3 or 2 MMAs
    const int kc = chunk_kblocks(ex == 0 ? Geo<MN, F16>::default_kc : (Geo<MN, F16>::default_kc * 3) / 2);
    static int dbg = -1;
    if (dbg < 0) {   // timing probes for kernel development: results are INVALID while one is active
        const char* e = getenv("MPVAE_TC_DEBUG");
        dbg = e ? atoi(e) : 0;
        if (dbg != 0) fprintf(stderr, "[mpvae tc] MPVAE_TC_DEBUG=%d: timing probe active, tensor-engine results are INVALID\n", dbg);
    }
    if (dbg == 3) return 0;                                                               // pre-passes only
    if (cta_group() == 2) {
        if (ex == 1) return launch_gemm_2sm<MN, F16, 1>(a, b, C, Mc, Nc, K, ldc, ma, mb, kc, dbg, stream, partials, partials_bytes);
        if (ex == 2) return launch_gemm_2sm<MN, F16, 2>(a, b, C, Mc, Nc, K, ldc, ma, mb, kc, dbg, stream, partials, partials_bytes);
        return launch_gemm_2sm<MN, F16, 0>(a, b, C, Mc, Nc, K, ldc, ma, mb, kc, dbg, stream, partials, partials_bytes);
    }
    if (ex != 0) { set_error("exact-operand products need the CTA-pair kernel"); return 7; }
    static bool configured = false;
    if (!configured) {
        const cudaError_t e = cudaFuncSetAttribute(gemm_split_kernel<MN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(gemm_split): %s", cudaGetErrorString(e)); return 4; }
        configured = true;
    }
    const int tiles_m = ceil_div(Mc, BM), tiles_n = ceil_div(Nc, BN);
    const int tiles = tiles_m * tiles_n;
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    gemm_split_kernel<MN, F16><<<grid, kThreads, SMEM_BYTES, stream>>>(a, b, C, Mc, Nc, K, ldc, tiles_m, tiles_n, kc, ma, mb, dbg);
    return check_launch("gemm_split_kernel");
}

}  // namespace

bool tc_available() { return true; }

// [absmax slots] [A planes] [B planes] [tail-wave scratch]
size_t tc_workspace_nt(int M, int N, int K) {
    const size_t kp = pitch_of(K);
    return 256 + planes_bytes(M, kp) + planes_bytes(N, kp) + tc_tail_scratch_bytes();
}

size_t tc_workspace_tn(int M, int N1, int N2) {
    return 256 + planes_bytes(M, pitch_of(N1)) + planes_bytes(M, pitch_of(N2)) + tc_tail_scratch_bytes();
}

int tc_contract_nt(const float* A, const float* Bm, float* C, int M, int N, int K, void* ws, size_t ws_bytes,
                   cudaStream_t stream, int reuse_planes, int exact, int ldc, int allow_ksplit) {
    if (!ws || ws_bytes < tc_workspace_nt(M, N, K)) { set_error("tc_contract_nt: workspace too small"); return 5; }
    if (ldc <= 0) ldc = N;
    const int ex = exact ? 1 : 0;
    const bool f16 = use_f16();
    const int kp = pitch_of(K);
    const Scratch s = carve_scratch(ws, M, kp);
    if (!reuse_planes) {
        if (f16 && cudaMemsetAsync(s.absmax, 0, 256, stream) != cudaSuccess) { set_error("cudaMemsetAsync failed"); return 2; }
        if (int rc = split_operand(A, s.a, M, K, kp, f16, s.absmax, stream)) return rc;
        if (int rc = split_operand(Bm, s.b, N, K, kp, f16, s.absmax + 1, stream)) return rc;
    }
    CUtensorMap ma, mb;
    const int bk = f16 ? 64 : 32;
    if (int rc = make_map(&ma, s.a, f16, K, M, kp, bk, BM, CU_TENSOR_MAP_SWIZZLE_128B, ex ? 1 : 2)) return rc;
    if (int rc = make_map(&mb, s.b, f16, K, N, kp, bk, nt_b_box_rows(), CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    // K-sliced partial waves only on request: without them each element's summation order is a function of K alone,
    // which the loss path needs (a row's predictions must not depend on how many rows the call holds)
    float* tail = allow_ksplit ? reinterpret_cast<float*>(s.b + planes_bytes(N, kp)) : nullptr;
    const size_t tail_bytes = allow_ksplit ? tc_tail_scratch_bytes() : 0;
    if (f16) return launch_gemm<false, true>(ma, mb, C, M, N, K, ldc, s.absmax, s.absmax + 1, stream, ex, tail, tail_bytes);
    return launch_gemm<false, false>(ma, mb, C, M, N, K, ldc, nullptr, nullptr, stream, ex, tail, tail_bytes);
}

int tc_contract_tn(const float* A, const float* Bm, float* C, int M, int N1, int N2, void* ws, size_t ws_bytes,
                   cudaStream_t stream, int reuse_planes, int exact) {
    if (!ws || ws_bytes < tc_workspace_tn(M, N1, N2)) { set_error("tc_contract_tn: workspace too small"); return 5; }
    const int ex = exact ? 2 : 0;
    const bool f16 = use_f16();
    const int p1 = pitch_of(N1), p2 = pitch_of(N2);
    const Scratch s = carve_scratch(ws, M, p1);
    if (!reuse_planes) {
        if (f16 && cudaMemsetAsync(s.absmax, 0, 256, stream) != cudaSuccess) { set_error("cudaMemsetAsync failed"); return 2; }
        if (int rc = split_operand(A, s.a, M, N1, p1, f16, s.absmax, stream)) return rc;
        if (int rc = split_operand(Bm, s.b, M, N2, p2, f16, s.absmax + 1, stream)) return rc;
    }
    CUtensorMap ma, mb;
    const int bk = f16 ? 64 : 32, box_mn = f16 ? 64 : 32;
    const CUtensorMapSwizzle sw = f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    if (int rc = make_map(&ma, s.a, f16, N1, M, p1, box_mn, bk, sw)) return rc;
    if (int rc = make_map(&mb, s.b, f16, N2, M, p2, box_mn, bk, sw, ex ? 1 : 2)) return rc;
    float* tail = reinterpret_cast<float*>(s.b + planes_bytes(M, p2));
    if (f16) return launch_gemm<true, true>(ma, mb, C, N1, N2, M, N2, s.absmax, s.absmax + 1, stream, ex, tail, tc_tail_scratch_bytes());
    return launch_gemm<true, false>(ma, mb, C, N1, N2, M, N2, nullptr, nullptr, stream, ex, tail, tc_tail_scratch_bytes());
}

// ------------------------------------------------------------------------------------------------ staged interface
size_t tc_planes_bytes(int rows, int cols) { return planes_bytes((size_t)rows, (size_t)pitch_of(cols)); }

int tc_split(const float* src, int rows, int cols, void* planes, uint32_t* absmax, int compute_absmax, cudaStream_t stream,
             int src_pitch) {
    const bool f16 = use_f16();
    if (src_pitch <= 0) src_pitch = cols;
    const int pitch = pitch_of(cols);
    const size_t n = (size_t)rows * pitch;
    if (!f16) {
        split_tf32_kernel<<<grid_for(n), 256, 0, stream>>>(src, static_cast<float*>(planes), rows, cols, src_pitch, pitch, n);
        return check_launch("split_tf32_kernel");
    }
    if (absmax && compute_absmax) {
        absmax_kernel<<<grid_for((size_t)rows * cols), 256, 0, stream>>>(src, (size_t)rows * cols, absmax);
        if (int rc = check_launch("absmax_kernel")) return rc;
    }
    split_f16_kernel<<<grid_for(n / 4), 256, 0, stream>>>(src, static_cast<__half*>(planes), rows, cols, src_pitch, pitch, n, absmax);
    return check_launch("split_f16_kernel");
}

namespace {

// Philox normals written straight into operand planes (scale 1: |n| < 6 fits fp16).  One thread per counter, same
// counter -> element mapping as philox_normal_kernel, so the numbers are identical to the fp32 tensor it would write.
template <bool F16>
__global__ void __launch_bounds__(256)
philox_planes_kernel(void* __restrict__ planes, int S, int B, int Z, int pitch, int Bg, int row0, uint2 key, uint2 off,
                     const unsigned long long* __restrict__ off_dev, int write_lo) {
    const int s = blockIdx.y;
    const unsigned long long span_beg = ((unsigned long long)s * Bg + row0) * Z;
    const unsigned long long span_end = span_beg + (unsigned long long)B * Z;
    const unsigned long long c = (span_beg >> 2) + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if ((c << 2) >= span_end) return;
    float n[4];
    philox_normal4(c, key, philox_offset(off, off_dev), n);
    const size_t plane = (size_t)S * B * pitch;
    // position of flat index 4c inside this sample's (B, Z) block (B * Z < 2^31 is checked by the launcher);
    // one 32-bit division per thread, then walk the four elements
    const long long e0 = (long long)((c << 2) - span_beg);               // may be -3..-1 for the first counter
    int row = e0 >= 0 ? (int)((unsigned)e0 / (unsigned)Z) : -1;
    int col = e0 >= 0 ? (int)((unsigned)e0 % (unsigned)Z) : Z + (int)e0;
    // n[j] already lies on the fp16 grid (philox_normal4), so hi = n and lo = 0 in both kinds
    if (F16 && !write_lo && row >= 0 && row < B && col + 4 <= Z) {
        // the common case: four neighbours of one row; as few stores as the alignment of the destination allows
        // (it is the same for every thread of a row, so a warp does not diverge here)
        __half* dst = static_cast<__half*>(planes) + ((size_t)s * B + (size_t)row) * pitch + (size_t)col;
        const __half2 p01 = __floats2half2_rn(n[0], n[1]), p23 = __floats2half2_rn(n[2], n[3]);
        const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
        if ((a & 7u) == 0) {
            uint2 v;
            v.x = *reinterpret_cast<const uint32_t*>(&p01);
            v.y = *reinterpret_cast<const uint32_t*>(&p23);
            *reinterpret_cast<uint2*>(dst) = v;
        } else if ((a & 3u) == 0) {
            *reinterpret_cast<__half2*>(dst) = p01;
            *reinterpret_cast<__half2*>(dst + 2) = p23;
        } else {
            dst[0] = __low2half(p01);
            *reinterpret_cast<__half2*>(dst + 1) = __halves2half2(__high2half(p01), __low2half(p23));
            dst[3] = __high2half(p23);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = row, cc = col;
        if (++col == Z) { col = 0; ++row; }
        if (r < 0 || r >= B) continue;
        const size_t o = ((size_t)s * B + (size_t)r) * pitch + (size_t)cc;
        if (F16) {
            static_cast<__half*>(planes)[o] = __float2half_rn(n[j]);
            if (write_lo) static_cast<__half*>(planes)[plane + o] = __float2half_rn(0.0f);
        } else {
            static_cast<float*>(planes)[o] = n[j];
            if (write_lo) static_cast<float*>(planes)[plane + o] = 0.0f;
        }
    }
}

}  // namespace

int tc_philox_planes(void* planes, int S, int B, int Z, int Bg, int row0, uint64_t seed, uint64_t offset,
                     const uint64_t* offset_dev, cudaStream_t stream) {
    if (S > 65535) { set_error("philox: S=%d exceeds grid.y limit", S); return 6; }
    if ((long long)B * Z >= 0x7fffffffLL || Z < 4) { set_error("philox planes: B*Z=%lld, Z=%d out of range", (long long)B * Z, Z); return 6; }
    const unsigned long long counters = ((unsigned long long)B * Z + 3) / 4 + 1;
    dim3 grid((unsigned)((counters + 255) / 256), S);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint2 off = make_uint2((uint32_t)offset, (uint32_t)(offset >> 32));
    const unsigned long long* od = reinterpret_cast<const unsigned long long*>(offset_dev);
    const int write_lo = tc_exact_supported() ? 0 : 1;   // the single-CTA kernel still reads a (zero) lo plane
    if (use_f16()) philox_planes_kernel<true><<<grid, 256, 0, stream>>>(planes, S, B, Z, pitch_of(Z), Bg, row0, key, off, od, write_lo);
    else philox_planes_kernel<false><<<grid, 256, 0, stream>>>(planes, S, B, Z, pitch_of(Z), Bg, row0, key, off, od, write_lo);
    return check_launch("philox_planes_kernel");
}

bool tc_exact_supported() { return cta_group() == 2; }

bool tc_f16_kind() { return use_f16(); }
int tc_pitch(int cols) { return pitch_of(cols); }
int tc_absmax(const float* src, size_t n, uint32_t* out_bits, cudaStream_t stream) {
    absmax_kernel<<<grid_for(n), 256, 0, stream>>>(src, n, out_bits);
    return check_launch("absmax_kernel");
}

size_t tc_tail_scratch_bytes() { return (size_t)(kNumSMs / 2 - 1) * 256 * BN * sizeof(float); }

int tc_gemm_nt(const void* a_planes, const void* b_planes, float* C, int M, int N, int K, const uint32_t* absmax_a,
               const uint32_t* absmax_b, cudaStream_t stream, int ldc, int a_exact, void* tail_scratch, size_t tail_scratch_bytes) {
    if (ldc <= 0) ldc = N;
    const bool f16 = use_f16();
    const int kp = pitch_of(K), bk = f16 ? 64 : 32;
    CUtensorMap ma, mb;
    if (int rc = make_map(&ma, a_planes, f16, K, M, kp, bk, BM, CU_TENSOR_MAP_SWIZZLE_128B, a_exact ? 1 : 2)) return rc;
    if (int rc = make_map(&mb, b_planes, f16, K, N, kp, bk, nt_b_box_rows(), CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    float* ts = static_cast<float*>(tail_scratch);
    if (f16) return launch_gemm<false, true>(ma, mb, C, M, N, K, ldc, absmax_a, absmax_b, stream, a_exact ? 1 : 0, ts, tail_scratch_bytes);
    return launch_gemm<false, false>(ma, mb, C, M, N, K, ldc, nullptr, nullptr, stream, a_exact ? 1 : 0, ts, tail_scratch_bytes);
}

int tc_gemm_tn(const void* a_planes, const void* b_planes, float* C, int M, int N1, int N2, const uint32_t* absmax_a,
               const uint32_t* absmax_b, cudaStream_t stream, int b_exact, void* tail_scratch, size_t tail_scratch_bytes,
               int a_pitch) {
    const bool f16 = use_f16();
    // a_pitch > 0: A is a column range of wider planes (row slab of C): a_planes points at its first column
    const int p1 = a_pitch > 0 ? a_pitch : pitch_of(N1), p2 = pitch_of(N2);
    const int bk = f16 ? 64 : 32, box_mn = f16 ? 64 : 32;
    const CUtensorMapSwizzle sw = f16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    CUtensorMap ma, mb;
    if (int rc = make_map(&ma, a_planes, f16, N1, M, p1, box_mn, bk, sw)) return rc;
    if (int rc = make_map(&mb, b_planes, f16, N2, M, p2, box_mn, bk, sw, b_exact ? 1 : 2)) return rc;
    float* ts = static_cast<float*>(tail_scratch);
    if (f16) return launch_gemm<true, true>(ma, mb, C, N1, N2, M, N2, absmax_a, absmax_b, stream, b_exact ? 2 : 0, ts, tail_scratch_bytes);
    return launch_gemm<true, false>(ma, mb, C, N1, N2, M, N2, nullptr, nullptr, stream, b_exact ? 2 : 0, ts, tail_scratch_bytes);
}

}  // namespace mpv
