// Step kernels for float path storage, float state (all degrees); see lsm_step.cuh.
#include "lsm_step.cuh"

namespace amc {

cudaError_t launch_step_f32s(int degree, int grid, const StepArgs& a, cudaStream_t s, bool pdl, int n_batch) {
    return launch_step_d<float, float>(degree, grid, a, s, pdl, n_batch);
}

int step_occupancy_f32s(int degree) { return occupancy_d<float, float>(degree); }

}  // namespace amc
