// minimax.cu -- batched scripted opponent: the reference's MinimaxPlayer(depth 2, Voronoi heuristic)
// (tron/minimax.py:58-310) for one player of every game, one warp per game, one lane per (my move, enemy move) leaf.
//
// The reference's search is reproduced with its quirks: it runs on the TRANSPOSED observation of the moving player
// (minimax.py:298); heads are found with argmax / argmin over the flattened map (:151-153,219-220); get_shortest_path
// (:64-86) is a FIFO over an ordered SET of (x, y, l) tuples that marks a cell only when it is popped, so a cell can be
// queued again by a same-level neighbour with a larger l and its distance is then overwritten; get_voronoi_value (:88-123)
// counts cells literally as coded (enemy-body cells, value -3, count for player 1); a mover without a free neighbour
// leaves its node at value 0 (:233-234); root ties use first-best or a Philox draw (random.choice / randint, :234,267).
// Each lane keeps its own map, two distance maps and the BFS queue in local memory (L1-resident, ~5 KB); the 16 leaf
// values are reduced with shuffles (min over the enemy move, max over my move).  Grids up to 256 cells (W,H <= 14).
#include "common.cuh"
#include "launch.h"

namespace tron {

constexpr int kMmMaxC = 256;
constexpr int kMmQueue = 4 * kMmMaxC + 8;
enum : uint32_t { TAG_MINIMAX = 8 };

struct MmMap {
    int rows, cols;
    short v[kMmMaxC];
};

__device__ __forceinline__ int mm_wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? n - 1 : i); }
__device__ __forceinline__ int mm_at(const MmMap& m, int i0, int i1) { return mm_wrap(i0, m.rows) * m.cols + mm_wrap(i1, m.cols); }
__device__ int mm_arg(const MmMap& m, bool want_max) {
    int best = 0;
    for (int i = 1; i < m.rows * m.cols; ++i)
        if (want_max ? m.v[i] > m.v[best] : m.v[i] < m.v[best]) best = i;
    return best;
}
__device__ __forceinline__ int mm_d0(int k) { return (k == 1) - (k == 3); }   // (x,y-1) (x+1,y) (x,y+1) (x-1,y)
__device__ __forceinline__ int mm_d1(int k) { return (k == 2) - (k == 0); }

// reference get_shortest_path (minimax.py:64-86): FIFO over a set of (x, y, l); cells are marked when popped
__device__ void mm_shortest_path(const MmMap& gm, int ind, int pl_mi, MmMap& dist, short* qx, short* qy, short* ql) {
    dist.rows = gm.rows; dist.cols = gm.cols;
    for (int i = 0; i < gm.rows * gm.cols; ++i) dist.v[i] = gm.v[i];
    int head = 0, tail = 0;
    qx[0] = (short)(ind / gm.cols); qy[0] = (short)(ind % gm.cols); ql[0] = (short)pl_mi; tail = 1;
    while (head < tail) {
        const int x = qx[head], y = qy[head], l = ql[head];
        ++head;
        dist.v[mm_at(dist, x, y)] = (short)(l + pl_mi);
        for (int k = 0; k < 4; ++k) {
            const int nx = x + mm_d0(k), ny = y + mm_d1(k);
            if (dist.v[mm_at(dist, nx, ny)] != 1) continue;
            bool dup = false;
            for (int q = head; q < tail; ++q) dup |= (qx[q] == nx && qy[q] == ny && ql[q] == l + pl_mi);
            if (!dup && tail < kMmQueue) { qx[tail] = (short)nx; qy[tail] = (short)ny; ql[tail] = (short)(l + pl_mi); ++tail; }
        }
    }
}
// reference get_voronoi_value (minimax.py:88-123)
__device__ int mm_voronoi(const MmMap& gm, MmMap& p1, MmMap& p2, short* qx, short* qy, short* ql) {
    mm_shortest_path(gm, mm_arg(gm, true), 1, p1, qx, qy, ql);
    mm_shortest_path(gm, mm_arg(gm, false), -1, p2, qx, qy, ql);
    int a1 = 0, a2 = 0;
    for (int i = 0; i < gm.rows * gm.cols; ++i) {
        const int u = p1.v[i], w = p2.v[i];
        if (u == -1 || u == 2 || w == -2) continue;
        if (u != 1 && w == 1) a1++;
        else if (u == 1 && w != 1) a2++;
        else if (u + w < 0) a1++;
        else if (u + w > 0) a2++;
    }
    return a1 - a2;
}
// reference get_blocked (minimax.py:168-203): bit k of `free_or_crash` = move k is expanded, returns true when no move is free
__device__ bool mm_blocked(const MmMap& gm, int deo, int* expand_mask) {
    const int ind = mm_arg(gm, deo == 1), x = ind / gm.cols, y = ind % gm.cols;
    bool all = true;
    int mask = 0;
    for (int k = 0; k < 4; ++k) {
        const int v = gm.v[mm_at(gm, x + mm_d0(k), y + mm_d1(k))];
        if (v == 1) { all = false; mask |= 1 << k; }
        else if (v == 10) mask |= 1 << k;  // crash into my head: still expanded
    }
    *expand_mask = mask;
    return all;
}
// reference get_next_map (minimax.py:147-166), in place
__device__ void mm_move(MmMap& gm, int action, int deo) {
    const int ind = mm_arg(gm, deo == 1), x = ind / gm.cols, y = ind % gm.cols;
    gm.v[mm_at(gm, x + mm_d0(action), y + mm_d1(action))] = (short)(10 * deo);
    gm.v[ind] = -1;
}

__global__ void __launch_bounds__(128) minimax_kernel(const int8_t* __restrict__ tiles, int n, int W, int H, int player, int tie_mode,
                                                      unsigned long long seed, unsigned long long counter, const unsigned long long* counter_dev,
                                                      unsigned long long base, uint8_t* actions, int* values, int* child_ties) {
    const int lane = threadIdx.x & 31;
    const long long env = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (env >= n) return;
    const int C = (W + 2) * (H + 2);
    // colour of Tile.value -1..6 for this player (reference map.py:67-81)
    const int colour[8] = {-1, 1, player == 1 ? -2 : -3, player == 1 ? 10 : -10, player == 1 ? -3 : -2, player == 1 ? -10 : 10,
                           player == 1 ? -2 : -3, player == 1 ? -3 : -2};
    MmMap gm, p1, p2;
    short qx[kMmQueue], qy[kMmQueue], ql[kMmQueue];
    gm.rows = H + 2; gm.cols = W + 2;
    const int8_t* t = tiles + (size_t)env * C;
    for (int r = 0; r < W + 2; ++r)  // transposed observation (minimax.py:298)
        for (int c = 0; c < H + 2; ++c) gm.v[c * (W + 2) + r] = (short)colour[(t[r * (H + 2) + c] + 1) & 7];

    int mask0 = 0;
    const bool root_blocked = mm_blocked(gm, 1, &mask0);
    const int a = (lane >> 2) & 3, b = lane & 3;
    int leaf = INT_MAX;
    bool enemy_boxed = false;
    if (!root_blocked && lane < 16 && ((mask0 >> a) & 1)) {
        mm_move(gm, a, 1);
        int mask1 = 0;
        if (mm_blocked(gm, -1, &mask1)) { leaf = 0; enemy_boxed = true; }  // enemy has no free move: the node keeps its initial value 0 (minimax.py:233)
        else if ((mask1 >> b) & 1) {
            mm_move(gm, b, -1);
            leaf = mm_voronoi(gm, p1, p2, qx, qy, ql);
        }
    }
    int v = leaf;
    v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, 1));
    v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, 2));  // min over the enemy's moves
    const unsigned m_min = __ballot_sync(0xFFFFFFFFu, leaf != INT_MAX && leaf == v), m_box = __ballot_sync(0xFFFFFFFFu, enemy_boxed);
    int val[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = __shfl_sync(0xFFFFFFFFu, v, 4 * k);
        val[k] = ((mask0 >> k) & 1) && !root_blocked ? x : INT_MIN;
    }
    if (lane == 0) {
        if (counter_dev) counter += *counter_dev;
        const uint4 rnd = philox(seed, counter, base + (unsigned long long)env, TAG_MINIMAX, (uint32_t)player);
        int act;
        if (root_blocked) {
            act = tie_mode ? (int)(rnd.x >> 30) : 0;  // random.randint(1, 4) (minimax.py:234)
        } else {
            int best = INT_MIN, cnt = 0, list[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) if (val[k] != INT_MIN && val[k] > best) best = val[k];
#pragma unroll
            for (int k = 0; k < 4; ++k) if (val[k] == best) list[cnt++] = k;
            act = list[tie_mode ? (int)__umulhi(rnd.x, (uint32_t)cnt) : 0];  // random.choice (minimax.py:267)
        }
        actions[env] = (uint8_t)act;
        if (values) { values[4 * env] = val[0]; values[4 * env + 1] = val[1]; values[4 * env + 2] = val[2]; values[4 * env + 3] = val[3]; }
        if (child_ties) {  // what the depth-1 node of each root move draws from the reference's global RNG (see the header)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                child_ties[4 * env + k] = val[k] == INT_MIN ? -1 : ((m_box >> (4 * k)) & 1u) ? 0 : __popc((m_min >> (4 * k)) & 0xFu);
        }
    }
}

int launch_minimax(const int8_t* tiles, int n, int W, int H, int player, int tie_mode, uint64_t seed, uint64_t counter, const uint64_t* cdev,
                   uint64_t base, uint8_t* actions, int32_t* values, int32_t* child_ties, cudaStream_t s) {
    minimax_kernel<<<(n + 3) / 4, 128, 0, s>>>(tiles, n, W, H, player, tie_mode, seed, counter, (const unsigned long long*)cdev, base, actions, values, child_ties);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

}  // namespace tron
