// step_c144.cu -- fused tick kernels specialised for the reference's config.py grid (10x10 -> 144 cells).
#include "step_dispatch.cuh"

namespace tron {
int tile_envs_c144() { return 128; }
int launch_step_c144(const StepParams& p, int mode, int obs_dtype, int enc_kind, cudaStream_t s) {
    return launch_mode<144, 8>(p, mode, obs_dtype, enc_kind, s);
}
}  // namespace tron
