// CostRegNet tensor-core path (tcgen05 / TMEM implicit GEMM, bf16 operands, fp32 accumulate).
// Placeholder until the kernels land: the entry points report MVS_ERR_UNSUPPORTED.
#include "common.cuh"

namespace mvs {

size_t costreg_tc_workspace_bytes(int, int, int, int) { return 0; }

int costreg_tc(const float *, const mvs_costreg_params *, float *, void *, int, int, int, int, cudaStream_t) {
    return set_error(MVS_ERR_UNSUPPORTED, "MVS_PRECISION_BF16 (tcgen05 path) is not built in this version");
}

}  // namespace mvs
