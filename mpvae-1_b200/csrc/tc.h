// tcgen05 (5th-gen tensor core) contraction engine: 3xTF32 split-precision GEMMs for the dense regime
// (label / rank sets >= 128).  See contract_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace mpv {

bool tc_available();
// scratch for operand staging (hi/lo tf32 splits, K padded to the TMA box)
size_t tc_workspace_nt(int M, int N, int K);
size_t tc_workspace_tn(int M, int N1, int N2);
// C[M,N] = A[M,K] . B[N,K]^T
int tc_contract_nt(const float* A, const float* Bm, float* C, int M, int N, int K, void* ws, size_t ws_bytes,
                   cudaStream_t stream);
// C[N1,N2] = A[M,N1]^T . B[M,N2]
int tc_contract_tn(const float* A, const float* Bm, float* C, int M, int N1, int N2, void* ws, size_t ws_bytes,
                   cudaStream_t stream);

}  // namespace mpv
