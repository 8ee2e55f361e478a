// Family L Euler-Maruyama (fixed step and per-trial step doubling with the state in global memory).
#include "odecol_common.cuh"

namespace odecol {

size_t stage_em_fwd_workspace_bytes(const DevProblem&, int) { return 0; }

int stage_em_fwd(const DevProblem&, const float*, int, const float*, float*, const float*, uint64_t, int64_t, float, int,
                 float, float, float, int*, int*, int*, void*, size_t, cudaStream_t) {
    return ODECOL_E_UNSUPPORTED;
}

}  // namespace odecol
