// CUDA-core (FFMA) contraction kernels: the low-rank product of mpvae.py:165-170 and its backward.
//
//   nt: nr[m, l]  = sum_z noise[m, z] * R[l, z]            m = s*B + b      (forward, mpvae.py:168)
//   tn: g_R[l, z] = sum_m gxs[m, l]  * noise[m, z]                          (backward)
//
// Classic shared-memory tiled SGEMM with a register micro-tile; exact fp32 FMA accumulation in a fixed
// k order (deterministic).  This is the engine for small label / rank sets and the always-available
// fp32 cross-check for the tcgen05 path (contract_tc.cu) used when Z >= 128.
#include "common.cuh"
#include "rows.h"

namespace mpv {
namespace {

template <int T, int BDIM>
__device__ __forceinline__ int tile_index(int t, int j) {
    // 8-wide micro-tiles are split 4 + 4 across the two halves of the block tile so that the float4
    // shared-memory reads of neighbouring threads stay bank-conflict free.
    if (T == 8) return (j < 4) ? t * 4 + j : BDIM / 2 + t * 4 + (j - 4);
    return t * T + j;
}

template <int BM, int BN, int BK, int TM, int TN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_fma_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, int M, int N, int K,
                int lda, int ldb, int ldc, int k_per_split, size_t split_stride) {
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll 4
        for (int e = tid; e < BM * BK; e += NT) {
            int row, col;
            if (A_KMAJOR) { row = e / BK; col = e % BK; } else { col = e / BM; row = e % BM; }
            const int gm = m0 + row, gk = k0 + col;
            float v = 0.0f;
            if (gm < M && gk < kend) v = A_KMAJOR ? A[(size_t)gm * lda + gk] : A[(size_t)gk * lda + gm];
            As[col][row] = v;
        }
#pragma unroll 4
        for (int e = tid; e < BN * BK; e += NT) {
            int row, col;
            if (B_KMAJOR) { row = e / BK; col = e % BK; } else { col = e / BN; row = e % BN; }
            const int gn = n0 + row, gk = k0 + col;
            float v = 0.0f;
            if (gn < N && gk < kend) v = B_KMAJOR ? Bm[(size_t)gn * ldb + gk] : Bm[(size_t)gk * ldb + gn];
            Bs[col][row] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) av[i] = As[kk][tile_index<TM, BM>(ty, i)];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tile_index<TN, BN>(tx, j)];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    float* __restrict__ Cp = C + (size_t)blockIdx.z * split_stride;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + tile_index<TM, BM>(ty, i);
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tile_index<TN, BN>(tx, j);
            if (gn < N) Cp[(size_t)gm * ldc + gn] = acc[i][j];
        }
    }
}

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, size_t n, int splits) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];   // fixed order
        out[i] = s;
    }
}

template <bool AK, bool BK_>
int launch_gemm(const float* A, const float* Bm, float* C, int M, int N, int K, int lda, int ldb, int ldc, int splits,
                int k_per_split, size_t split_stride, cudaStream_t stream) {
    if (N <= 16) {
        dim3 grid(ceil_div(N, 16), ceil_div(M, 128), splits);
        gemm_fma_kernel<128, 16, 16, 8, 1, AK, BK_><<<grid, 256, 0, stream>>>(A, Bm, C, M, N, K, lda, ldb, ldc, k_per_split, split_stride);
    } else if (N <= 64 || M <= 64) {
        dim3 grid(ceil_div(N, 64), ceil_div(M, 64), splits);
        gemm_fma_kernel<64, 64, 16, 4, 4, AK, BK_><<<grid, 256, 0, stream>>>(A, Bm, C, M, N, K, lda, ldb, ldc, k_per_split, split_stride);
    } else {
        dim3 grid(ceil_div(N, 128), ceil_div(M, 128), splits);
        gemm_fma_kernel<128, 128, 8, 8, 8, AK, BK_><<<grid, 256, 0, stream>>>(A, Bm, C, M, N, K, lda, ldb, ldc, k_per_split, split_stride);
    }
    return check_launch("gemm_fma_kernel");
}

struct SplitPlan { int splits, k_per_split; };

SplitPlan plan_tn(int M, int N1, int N2) {
    // output tiles of the variant launch_gemm will pick (rows = N1, cols = N2, reduction = M)
    int tiles;
    if (N2 <= 16) tiles = ceil_div(N2, 16) * ceil_div(N1, 128);
    else if (N2 <= 64 || N1 <= 64) tiles = ceil_div(N2, 64) * ceil_div(N1, 64);
    else tiles = ceil_div(N2, 128) * ceil_div(N1, 128);
    // the reduction streams gxs from HBM with no reuse when N2 is small: keep ~8 CTAs per SM in flight
    int splits = ceil_div(8 * kNumSMs, tiles);
    const int max_splits = ceil_div(M, 64);
    if (splits > max_splits) splits = max_splits;
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
    int kps = ceil_div(ceil_div(M, splits), 16) * 16;
    splits = ceil_div(M, kps);
    return {splits, kps};
}

}  // namespace

int launch_contract_nt_fma(const float* A, const float* Bm, float* C, int M, int N, int K, cudaStream_t stream, int ldc) {
    return launch_gemm<true, true>(A, Bm, C, M, N, K, K, K, ldc > 0 ? ldc : N, 1, K, 0, stream);
}

size_t contract_tn_fma_workspace(int M, int N1, int N2) {
    const SplitPlan p = plan_tn(M, N1, N2);
    return p.splits > 1 ? (size_t)p.splits * N1 * N2 * sizeof(float) : 0;
}

int launch_contract_tn_fma(const float* A, const float* Bm, float* C, int M, int N1, int N2, void* ws, size_t ws_bytes,
                           cudaStream_t stream, int lda) {
    const SplitPlan p = plan_tn(M, N1, N2);
    const size_t n = (size_t)N1 * N2;
    if (lda <= 0) lda = N1;
    if (p.splits == 1) return launch_gemm<false, false>(A, Bm, C, N1, N2, M, lda, N2, N2, 1, M, 0, stream);
    if (ws == nullptr || ws_bytes < (size_t)p.splits * n * sizeof(float)) {
        set_error("contract_tn: workspace too small (%zu < %zu)", ws_bytes, (size_t)p.splits * n * sizeof(float));
        return 5;
    }
    float* part = static_cast<float*>(ws);
    int rc = launch_gemm<false, false>(A, Bm, part, N1, N2, M, lda, N2, N2, p.splits, p.k_per_split, n, stream);
    if (rc) return rc;
    const int blocks = (int)((n + 255) / 256 < 4 * kNumSMs ? (n + 255) / 256 : 4 * kNumSMs);
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(part, C, n, p.splits);
    return check_launch("splitk_reduce_kernel");
}

}  // namespace mpv
