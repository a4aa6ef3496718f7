// step_c144.cu -- fused tick kernels specialised for the reference's config.py grid (10x10 -> 144 cells).
#include "step_dispatch.cuh"

namespace tron {
// games per CTA for the 144-cell kernels: 128 (one per thread) when there are enough games to fill the GPU, fewer for small
// batches so that at least ~2 CTAs per SM exist (4096 games -> 16 per CTA -> 256 CTAs instead of 32)
int tile_envs_small_grid(int n_envs) {
    int g = n_envs / (2 * sm_count());
    g -= g % 16;
    return g < 16 ? 16 : (g > 128 ? 128 : g);
}
int tile_envs_c144(int n_envs) { return tile_envs_small_grid(n_envs); }
int launch_step_c144(const StepParams& p, int mode, int obs_dtype, int enc_kind, cudaStream_t s) {
    return launch_mode<144, 8>(p, mode, obs_dtype, enc_kind, s);
}
}  // namespace tron
