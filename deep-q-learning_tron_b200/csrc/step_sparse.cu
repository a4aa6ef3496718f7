// step_sparse.cu -- pure tick (no observation planes) for LARGE grids: thread-per-game directly on HBM.
//
// A tick touches at most six cells of a game (two old heads, two slide cells, two new heads), so staging a whole
// 64x64 grid (4,356 B) in shared memory would move ~100x more bytes than needed.  Here each thread reads and writes
// only those cells.  Auto-reset would normally rewrite the whole grid; instead every player carries the bounding box of
// the cells it has written this episode (a player only writes along its own path), kept in 8 bytes next to the
// per-game metadata, and a reset restores just the two boxes, warp-cooperatively.  The int8 Tile.value layout is
// unchanged, so observe/export and the fused tile kernel work on the same state.
#include "launch.h"
#include "tick_core.cuh"

namespace tron {

constexpr int kSparseThreads = 256;

__global__ void __launch_bounds__(kSparseThreads) step_sparse_kernel(const StepParams p) {
    const int tid = threadIdx.x, lane = tid & 31;
    const long long env = (long long)blockIdx.x * kSparseThreads + tid;
    const bool valid = env < p.N;
    const int C = p.C, Hc = p.Hc;
    EnvState e = unpack_meta(valid ? p.meta[env] : make_uint2(0, 0));
    BoxRegs bx;
    if (valid && (e.flags & TRON_FLAG_BOXES_VALID)) {
        bx = unpack_boxes(p.boxes[env]);
    } else {  // unknown history: first reset clears the whole board once
        bx.lo_r[0] = -1; bx.hi_r[0] = p.W; bx.lo_c[0] = -1; bx.hi_c[0] = p.H;
        bx.lo_r[1] = 0; bx.hi_r[1] = -1; bx.lo_c[1] = 0; bx.hi_c[1] = -1;
    }
    int8_t* g = p.grid + (size_t)(valid ? env : 0) * C;
    for (int t = 0; t < p.T; ++t) {
        ByteCells cells{g, Hc, C};
        const bool do_reset = valid ? env_tick<MODE_STEP, true>(cells, p, e, env, t, tid, bx) : false;
        unsigned mask = __ballot_sync(0xFFFFFFFFu, do_reset);
        const uint2 pb = pack_boxes(bx);
        while (mask) {  // all 32 lanes restore one finished game's boxes to the template (border WALL, interior EMPTY)
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const long long senv = __shfl_sync(0xFFFFFFFFu, env, src);
            const BoxRegs sb = unpack_boxes(make_uint2(__shfl_sync(0xFFFFFFFFu, pb.x, src), __shfl_sync(0xFFFFFFFFu, pb.y, src)));
            int8_t* sg = p.grid + (size_t)senv * C;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int w = sb.hi_c[i] - sb.lo_c[i] + 1, h = sb.hi_r[i] - sb.lo_r[i] + 1;
                if (w <= 0 || h <= 0) continue;
                const int area = w * h;
                for (int k = lane; k < area; k += 32) {
                    const int rr = k / w, r = sb.lo_r[i] + rr, c = sb.lo_c[i] + (k - rr * w);
                    const bool border = r < 0 || c < 0 || r >= p.W || c >= p.H;
                    sg[(r + 1) * Hc + c + 1] = border ? (int8_t)TRON_TILE_WALL : (int8_t)TRON_TILE_EMPTY;
                }
            }
            __syncwarp();
            if (lane == src) {  // heads of the fresh game go in after the clear
                g[(e.r1 + 1) * Hc + e.c1 + 1] = TRON_TILE_P1_HEAD;
                g[(e.r2 + 1) * Hc + e.c2 + 1] = TRON_TILE_P2_HEAD;
                box_set_spawn(bx, e);
                e.flags |= TRON_FLAG_BOXES_VALID;
            }
            __syncwarp();
        }
    }
    if (valid) {
        p.meta[env] = pack_meta(e);
        p.boxes[env] = pack_boxes(bx);
    }
}

int launch_step_sparse(const StepParams& p, cudaStream_t s) {
    const unsigned grid = (unsigned)(((long long)p.N + kSparseThreads - 1) / kSparseThreads);
    step_sparse_kernel<<<grid, kSparseThreads, 0, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

}  // namespace tron
