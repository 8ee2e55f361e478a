// Family L reverse sweep (discrete RK4 adjoint with the state in global memory).
#include "odecol_common.cuh"

namespace odecol {

size_t stage_rk4_bwd_workspace_bytes(const DevProblem&, int) { return 0; }

int stage_rk4_bwd(const DevProblem&, const float*, int, const float*, const float*, const int*, int, float*, float*,
                  void*, size_t, cudaStream_t) {
    return ODECOL_E_UNSUPPORTED;
}

}  // namespace odecol
