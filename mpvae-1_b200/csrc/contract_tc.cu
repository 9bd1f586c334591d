// tcgen05 contraction engine -- placeholder until the 3xTF32 kernels land (tc_available() == false routes
// every shape through the CUDA-core engine in contract_fma.cu).
#include "common.cuh"
#include "tc.h"

namespace mpv {

bool tc_available() { return false; }
size_t tc_workspace_nt(int, int, int) { return 0; }
size_t tc_workspace_tn(int, int, int) { return 0; }
int tc_contract_nt(const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t) {
    set_error("tensor engine not built");
    return 7;
}
int tc_contract_tn(const float*, const float*, float*, int, int, int, void*, size_t, cudaStream_t) {
    set_error("tensor engine not built");
    return 7;
}

}  // namespace mpv
