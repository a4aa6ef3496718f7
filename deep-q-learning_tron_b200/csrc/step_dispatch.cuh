// step_dispatch.cuh -- instantiates step_tile_kernel for one geometry class and dispatches at run time.
#pragma once
#include "launch.h"
#include "step_kernels.cuh"

namespace tron {

constexpr int kThreads = 128;
extern long long g_tile_ctas_per_sm;  // TRON_OPT_TILE_CTAS_PER_SM (abi.cu)
// measured defaults (profiles/r2_cta_cap_sweep.jsonl): 5 at 10x10 (+1.5-2 %), 6 for >= 2048 cells (+2 %), no cap in between (8x8, 32x32 lose)
inline long long default_tile_ctas(int C) { return C == 144 ? 5 : C >= 2048 ? 6 : 32; }

inline size_t tile_smem_bytes(int G, int C) {
    return (size_t)((G * C + 15) & ~15) + (size_t)((C + 15) & ~15) + (size_t)G * 4 + (size_t)G + 8 + 16 + 16 + 6 * sizeof(PlaneTab);
}

template <int C_T, int OD, int LP, bool CP, int CH, int MODE>
int launch_one(const StepParams& p, cudaStream_t s) {
    auto kern = step_tile_kernel<C_T, kThreads, OD, LP, CP, CH, MODE>;
    size_t smem = tile_smem_bytes(p.G, p.C);
    const unsigned n_ctas = (unsigned)(((long long)p.N + p.G - 1) / p.G);
    const long long cap = g_tile_ctas_per_sm > 0 ? g_tile_ctas_per_sm : default_tile_ctas(p.C);
    if (LP > 0 && MODE == MODE_STEP && cap < 32 && n_ctas > 8u * (unsigned)sm_count()) {  // see step_bits.cu: fewer write streams per SM
        const size_t want = (size_t)(228 * 1024) / (size_t)(cap + 1) - 1024 + 256;
        if (want > smem && want <= 200 * 1024) smem = want;
    }
    if (smem > 48 * 1024 && ensure_dynamic_smem((const void*)kern, smem) != TRON_OK) return TRON_ERR_CUDA;
    const unsigned grid = (unsigned)(((long long)p.N + p.G - 1) / p.G);
    kern<<<grid, kThreads, smem, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

template <int C_T, int OD, int CH, int MODE>
int launch_enc(const StepParams& p, int enc_kind, cudaStream_t s) {
    switch (enc_kind) {
        case 1: return launch_one<C_T, OD, 1, false, CH, MODE>(p, s);
        case 2: return launch_one<C_T, OD, 3, false, CH, MODE>(p, s);
        case 3: return launch_one<C_T, OD, 3, true, CH, MODE>(p, s);
        default: return TRON_ERR_INVALID;
    }
}

template <int C_T, int CH>
int launch_mode(const StepParams& p, int mode, int od, int enc_kind, cudaStream_t s) {
    // f32 planes: 4 cells per item so that one warp-wide STG.128 covers whole 32-byte sectors (8 cells would give each
    // lane two half-sector stores; measured 78% -> see profiles/r1_step_f32_lut1_2M.json)
    constexpr int CH32 = CH == 8 ? 4 : CH;
    if (mode == MODE_RESET) return launch_one<C_T, TRON_I8, 0, false, CH, MODE_RESET>(p, s);
    if (mode == MODE_STEP && enc_kind == 0) return launch_one<C_T, TRON_I8, 0, false, CH, MODE_STEP>(p, s);
    if (enc_kind == 0) return TRON_ERR_INVALID;
    if (mode == MODE_STEP) {
        if (od == TRON_BF16) return launch_enc<C_T, TRON_BF16, CH, MODE_STEP>(p, enc_kind, s);
        if (od == TRON_F32) return launch_enc<C_T, TRON_F32, CH32, MODE_STEP>(p, enc_kind, s);
        if (od == TRON_I8) return launch_enc<C_T, TRON_I8, CH, MODE_STEP>(p, enc_kind, s);
    } else if (mode == MODE_OBSERVE) {
        if (od == TRON_BF16) return launch_enc<C_T, TRON_BF16, CH, MODE_OBSERVE>(p, enc_kind, s);
        if (od == TRON_F32) return launch_enc<C_T, TRON_F32, CH32, MODE_OBSERVE>(p, enc_kind, s);
        if (od == TRON_I8) return launch_enc<C_T, TRON_I8, CH, MODE_OBSERVE>(p, enc_kind, s);
    }
    return TRON_ERR_INVALID;
}

}  // namespace tron
