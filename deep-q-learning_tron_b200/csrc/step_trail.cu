// step_trail.cu -- pure tick for LARGE grids on a TRAIL-LIST state (TRON_LAYOUT_TRAIL).
//
// A dense 64x64 grid is 4,356 B per game of which a tick touches ~6 cells scattered over several DRAM rows; the int8
// sparse kernel is therefore bound by random 32-byte DRAM operations (~15 per game-tick), not by bandwidth.  Here a game
// is ONE record whose hot part is contiguous:
//     [0:8)   tron_meta (heads, alive/done/winner flags, ep_len)      [8:10) n1   [10:12) n2   [12:16) reserved
//     [16: )  trail cells, interleaved by player: entry (k, player) at 16 + 2*(2k + player), 2 bytes = {p0 | slide<<7, p1}
// Every tile a player leaves behind (body, or slide tile in ice/temper mode) is appended to its list; walls are implicit
// and heads live in the meta.  "Is this cell free?" is a scan of both lists -- a handful of entries for typical episodes --
// and a reset is n1 = n2 = 0.  The first 12 ticks of both players sit in the record's first 64 bytes, so a game-tick is
// normally one 64-byte read and one or two sector writes.  Capacity is W*H entries per player (trail cells are distinct
// interior cells), so the representation is exact for any episode; tron_export_grid renders the Tile.value grid.
// Observations are not produced from this layout (use TRON_LAYOUT_TILE8 / BITS10 for the fused obs path).
#include "launch.h"
#include "step_kernels.cuh"

namespace tron {

constexpr int kTrailThreads = 128;
constexpr int kTrailHot = 24;  // entries held in registers (bytes 16..64 of the record)

// a multiple of 64 bytes: the hot head of a record (header + first 24 entries) is then exactly one 64-byte DRAM atom /
// two L2 sectors, and the kernel may always load the first 64 bytes
__host__ __device__ inline size_t trail_record_bytes(int W, int H) { return (16u + 4u * (size_t)W * (size_t)H + 63u) & ~(size_t)63u; }
size_t trail_record_bytes_host(int W, int H) { return trail_record_bytes(W, H); }

struct TrailCells {
    unsigned char* rec;       // this game's record in HBM
    int n[2];                 // entries per player at the start of the tick (those are in hot[] / HBM)
    uint32_t hot[kTrailHot / 2];  // first 24 entries as loaded
    unsigned short fresh[4];  // entries appended during this tick (two bodies, up to two slide tiles)
    int fresh_owner[4], n_fresh;
    int W, H;

    __device__ __forceinline__ static unsigned short pack(int r, int c, bool slide) { return (unsigned short)((r & 0x7F) | (slide ? 0x80 : 0) | (c << 8)); }

    __device__ __forceinline__ int get(int r, int c) const {
        if (r < 0 || c < 0 || r >= W || c >= H) return TRON_TILE_WALL;
        const unsigned key = (unsigned)(r & 0x7F) | ((unsigned)c << 8);
        bool hit = false;
#pragma unroll
        for (int w = 0; w < kTrailHot / 2; ++w) {  // word w = {entry (k=w, P1), entry (k=w, P2)}
            const unsigned e1 = hot[w] & 0xFF7Fu, e2 = (hot[w] >> 16) & 0xFF7Fu;
            hit |= (w < n[0] && e1 == key) | (w < n[1] && e2 == key);
        }
        const int nmax = max(n[0], n[1]);
        if (nmax > kTrailHot / 2) {  // long episode: the rest of the lists, straight from memory
            const uint32_t* words = (const uint32_t*)(rec + 16);
            for (int w = kTrailHot / 2; w < nmax; ++w) {
                const uint32_t v = words[w];
                hit |= (w < n[0] && (v & 0xFF7Fu) == key) | (w < n[1] && ((v >> 16) & 0xFF7Fu) == key);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) hit |= (i < n_fresh && (fresh[i] & 0xFF7Fu) == key);
        return hit ? TRON_TILE_P1_BODY : TRON_TILE_EMPTY;  // callers only test for EMPTY
    }
    __device__ __forceinline__ void put(int r, int c, int tile) {
        int owner;
        bool slide = false;
        if (tile == TRON_TILE_P1_BODY) owner = 0;
        else if (tile == TRON_TILE_P2_BODY) owner = 1;
        else if (tile == TRON_TILE_P1_SLIDE) { owner = 0; slide = true; }
        else if (tile == TRON_TILE_P2_SLIDE) { owner = 1; slide = true; }
        else return;  // heads are metadata
        if (r < 0 || c < 0 || r >= W || c >= H) return;  // a head that left the board leaves no trail tile there
        int cnt = n[owner];
#pragma unroll
        for (int i = 0; i < 4; ++i) cnt += (i < n_fresh && fresh_owner[i] == owner);
        const unsigned short e = pack(r, c, slide);
        ((unsigned short*)(rec + 16))[2 * cnt + owner] = e;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i == n_fresh) { fresh[i] = e; fresh_owner[i] = owner; }
        n_fresh = min(n_fresh + 1, 4);
    }
};

template <int MODE>
__global__ void __launch_bounds__(kTrailThreads, 8) step_trail_kernel(const StepParams p) {
    const int tid = threadIdx.x;
    const long long env = (long long)blockIdx.x * kTrailThreads + tid;
    if (env >= p.N) return;
    const size_t R = trail_record_bytes(p.W, p.H);
    unsigned char* rec = (unsigned char*)p.grid + (size_t)env * R;
    uint4 hdr = *(const uint4*)rec;
    EnvState e = unpack_meta(make_uint2(hdr.x, hdr.y));
    TrailCells g;
    g.rec = rec; g.W = p.W; g.H = p.H;
    g.n[0] = (int)(hdr.z & 0xFFFFu); g.n[1] = (int)(hdr.z >> 16);
    const int T = MODE == MODE_STEP ? p.T : 1;
    for (int t = 0; t < T; ++t) {
        if (MODE == MODE_STEP) {
            const uint4 a = *(const uint4*)(rec + 16), b = *(const uint4*)(rec + 32), c = *(const uint4*)(rec + 48);
            g.hot[0] = a.x; g.hot[1] = a.y; g.hot[2] = a.z; g.hot[3] = a.w; g.hot[4] = b.x; g.hot[5] = b.y; g.hot[6] = b.z; g.hot[7] = b.w;
            g.hot[8] = c.x; g.hot[9] = c.y; g.hot[10] = c.z; g.hot[11] = c.w;
        }
        g.n_fresh = 0;
        BoxRegs bx;
        const bool do_reset = env_tick<MODE, false>(g, p, e, env, t, tid, bx);
        if (do_reset) { g.n[0] = g.n[1] = 0; }
        else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (i < g.n_fresh) g.n[g.fresh_owner[i]] += 1;
        }
        if (T > 1) __threadfence_block();  // this thread re-reads its own appended entries on the next tick
    }
    const uint2 m = pack_meta(e);
    *(uint4*)rec = make_uint4(m.x, m.y, (uint32_t)g.n[0] | ((uint32_t)g.n[1] << 16), 0u);
}

// ---- fused tick + observation planes on the trail-list state (large grids) -------------------------------------------
// CTA = 128 threads, G games (G*C <= ~18 KB of shared memory).  Threads 0..G-1 tick their game on its record; then the
// whole CTA renders the G grids into shared memory (template fill, 16-byte stores), the owners scatter their trail
// entries and heads, and encode_tile() streams both players' planes out.  Compared with the int8 layout this skips the
// C-byte grid read and the C-byte write-back per game-tick (8.7 KB of 26 KB at 64x64 bf16).
template <int OD, int LP, bool CP, int CH, int MODE>
__global__ void __launch_bounds__(kTrailThreads) step_trail_obs_kernel(const StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = p.C, G = p.G, Hc = p.Hc, tid = threadIdx.x;
    int8_t* tile = (int8_t*)smem_raw;                       // [G][C]
    int8_t* tmpl = tile + ((G * C + 15) & ~15);             // [C]
    const long long env0 = (long long)blockIdx.x * G;
    const int nG = (int)min((long long)G, (long long)p.N - env0);
    for (int c = tid; c < C; c += kTrailThreads) {
        const int r = c / Hc, q = c - r * Hc;
        tmpl[c] = (r == 0 || r == p.W + 1 || q == 0 || q == p.H + 1) ? (int8_t)TRON_TILE_WALL : (int8_t)TRON_TILE_EMPTY;
    }
    const bool owner = tid < nG;
    const long long env = env0 + tid;
    const size_t R = trail_record_bytes(p.W, p.H);
    unsigned char* rec = (unsigned char*)p.grid + (size_t)(owner ? env : env0) * R;
    EnvState e = unpack_meta(make_uint2(0, 0));
    TrailCells g;
    g.rec = rec; g.W = p.W; g.H = p.H; g.n[0] = g.n[1] = 0;
    if (owner) {
        const uint4 hdr = *(const uint4*)rec;
        e = unpack_meta(make_uint2(hdr.x, hdr.y));
        g.n[0] = (int)(hdr.z & 0xFFFFu); g.n[1] = (int)(hdr.z >> 16);
    }
    __syncthreads();
    const int T = MODE == MODE_STEP ? p.T : 1;
    for (int t = 0; t < T; ++t) {
        if (MODE == MODE_STEP && owner) {
            const uint4 a = *(const uint4*)(rec + 16), b = *(const uint4*)(rec + 32), c = *(const uint4*)(rec + 48);
            g.hot[0] = a.x; g.hot[1] = a.y; g.hot[2] = a.z; g.hot[3] = a.w; g.hot[4] = b.x; g.hot[5] = b.y; g.hot[6] = b.z; g.hot[7] = b.w;
            g.hot[8] = c.x; g.hot[9] = c.y; g.hot[10] = c.z; g.hot[11] = c.w;
            g.n_fresh = 0;
            BoxRegs bx;
            if (env_tick<MODE_STEP, false>(g, p, e, env, t, tid, bx)) { g.n[0] = g.n[1] = 0; }
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (i < g.n_fresh) g.n[g.fresh_owner[i]] += 1;
            }
        }
        if (MODE == MODE_OBSERVE || p.obs_every_tick || t == T - 1) {
            // render: template for every game, then the owners scatter their trails and heads
            if ((C & 15) == 0) {
                const int per = C / 16;
                for (int it = tid; it < nG * per; it += kTrailThreads) { const int ee = it / per; ((uint4*)(tile + ee * C))[it - ee * per] = ((const uint4*)tmpl)[it - ee * per]; }
            } else if ((C & 3) == 0) {
                const int per = C / 4;
                for (int it = tid; it < nG * per; it += kTrailThreads) { const int ee = it / per; ((uint32_t*)(tile + ee * C))[it - ee * per] = ((const uint32_t*)tmpl)[it - ee * per]; }
            } else {
                for (int it = tid; it < nG * C; it += kTrailThreads) tile[it] = tmpl[it % C];
            }
            __syncthreads();
            if (owner) {
                int8_t* tl = tile + tid * C;
                const unsigned short* ent = (const unsigned short*)(rec + 16);
                const int nmax = max(g.n[0], g.n[1]);
                for (int k = 0; k < nmax; ++k) {
                    const uint32_t v = ((const uint32_t*)ent)[k];
                    if (k < g.n[0]) { const unsigned u = v & 0xFFFFu; tl[((u & 0x7F) + 1) * Hc + (u >> 8) + 1] = (u & 0x80) ? TRON_TILE_P1_SLIDE : TRON_TILE_P1_BODY; }
                    if (k < g.n[1]) { const unsigned u = v >> 16; tl[((u & 0x7F) + 1) * Hc + (u >> 8) + 1] = (u & 0x80) ? TRON_TILE_P2_SLIDE : TRON_TILE_P2_BODY; }
                }
                tl[(e.r1 + 1) * Hc + e.c1 + 1] = TRON_TILE_P1_HEAD;
                tl[(e.r2 + 1) * Hc + e.c2 + 1] = TRON_TILE_P2_HEAD;
            }
            __syncthreads();
            encode_tile<0, kTrailThreads, OD, LP, CP, CH>(tile, nG, env0, p, (MODE == MODE_STEP && p.obs_every_tick) ? t : 0);
            if (T > 1) __syncthreads();
        }
    }
    if (MODE == MODE_STEP && owner) {
        const uint2 m = pack_meta(e);
        *(uint4*)rec = make_uint4(m.x, m.y, (uint32_t)g.n[0] | ((uint32_t)g.n[1] << 16), 0u);
    }
}

template <int OD, int LP, bool CP, int MODE>
static int launch_trail_obs_one(StepParams p, cudaStream_t s) {
    // games per CTA: ~18 KB of tile like the int8 generic kernel
    int G = (int)(18432 / p.C);
    G = G < 1 ? 1 : (G > kTrailThreads ? kTrailThreads : G);
    p.G = G;
    const size_t smem = (size_t)((G * p.C + 15) & ~15) + (size_t)((p.C + 15) & ~15);
    const unsigned grid = (unsigned)(((long long)p.N + G - 1) / G);
    if ((p.C & 3) == 0) {
        auto k = step_trail_obs_kernel<OD, LP, CP, 4, MODE>;
        static size_t lim = 48 * 1024;
        if (smem > lim) { if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TRON_ERR_CUDA; lim = smem; }
        k<<<grid, kTrailThreads, smem, s>>>(p);
    } else {
        auto k = step_trail_obs_kernel<OD, LP, CP, 1, MODE>;
        static size_t lim = 48 * 1024;
        if (smem > lim) { if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return TRON_ERR_CUDA; lim = smem; }
        k<<<grid, kTrailThreads, smem, s>>>(p);
    }
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}
template <int OD, int MODE>
static int launch_trail_obs_enc(const StepParams& p, int enc_kind, cudaStream_t s) {
    switch (enc_kind) {
        case 1: return launch_trail_obs_one<OD, 1, false, MODE>(p, s);
        case 2: return launch_trail_obs_one<OD, 3, false, MODE>(p, s);
        case 3: return launch_trail_obs_one<OD, 3, true, MODE>(p, s);
        default: return TRON_ERR_INVALID;
    }
}
int launch_step_trail_obs(const StepParams& p, int mode, int od, int enc_kind, cudaStream_t s) {
    if (mode == MODE_STEP) {
        if (od == TRON_BF16) return launch_trail_obs_enc<TRON_BF16, MODE_STEP>(p, enc_kind, s);
        if (od == TRON_F32) return launch_trail_obs_enc<TRON_F32, MODE_STEP>(p, enc_kind, s);
        if (od == TRON_I8) return launch_trail_obs_enc<TRON_I8, MODE_STEP>(p, enc_kind, s);
    } else if (mode == MODE_OBSERVE) {
        if (od == TRON_BF16) return launch_trail_obs_enc<TRON_BF16, MODE_OBSERVE>(p, enc_kind, s);
        if (od == TRON_F32) return launch_trail_obs_enc<TRON_F32, MODE_OBSERVE>(p, enc_kind, s);
        if (od == TRON_I8) return launch_trail_obs_enc<TRON_I8, MODE_OBSERVE>(p, enc_kind, s);
    }
    return TRON_ERR_INVALID;
}

int launch_step_trail(const StepParams& p, int mode, cudaStream_t s) {
    const unsigned grid = (unsigned)(((long long)p.N + kTrailThreads - 1) / kTrailThreads);
    if (mode == MODE_STEP) step_trail_kernel<MODE_STEP><<<grid, kTrailThreads, 0, s>>>(p);
    else if (mode == MODE_RESET) step_trail_kernel<MODE_RESET><<<grid, kTrailThreads, 0, s>>>(p);
    else return TRON_ERR_UNSUPPORTED;
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

// ---- export: render Tile.value grids / metadata from the records; import: rebuild the lists from grids ---------------
__global__ void trail_export_kernel(const unsigned char* __restrict__ recs, int n, int W, int H, int8_t* tiles, int8_t* heads, uint8_t* alive,
                                    uint8_t* done, uint8_t* winner, int32_t* ep_len) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    const size_t R = trail_record_bytes(W, H);
    const unsigned char* rec = recs + (size_t)env * R;
    const uint4 hdr = *(const uint4*)rec;
    const EnvState e = unpack_meta(make_uint2(hdr.x, hdr.y));
    const int n1 = (int)(hdr.z & 0xFFFFu), n2 = (int)(hdr.z >> 16), Hc = H + 2, C = (W + 2) * (H + 2);
    if (tiles) {
        int8_t* t = tiles + (size_t)env * C;
        for (int c = 0; c < C; ++c) {
            const int r = c / Hc, q = c - r * Hc;
            t[c] = (r == 0 || r == W + 1 || q == 0 || q == H + 1) ? (int8_t)TRON_TILE_WALL : (int8_t)TRON_TILE_EMPTY;
        }
        const unsigned short* ent = (const unsigned short*)(rec + 16);
        for (int k = 0; k < max(n1, n2); ++k)
            for (int pl = 0; pl < 2; ++pl) {
                if (k >= (pl ? n2 : n1)) continue;
                const unsigned short v = ent[2 * k + pl];
                const int r = v & 0x7F, c = v >> 8;
                const bool slide = v & 0x80;
                t[(r + 1) * Hc + c + 1] = (int8_t)(pl ? (slide ? TRON_TILE_P2_SLIDE : TRON_TILE_P2_BODY) : (slide ? TRON_TILE_P1_SLIDE : TRON_TILE_P1_BODY));
            }
        t[(e.r1 + 1) * Hc + e.c1 + 1] = TRON_TILE_P1_HEAD;
        t[(e.r2 + 1) * Hc + e.c2 + 1] = TRON_TILE_P2_HEAD;  // written second (reference game.py:205-214)
    }
    if (heads) ((uint32_t*)heads)[env] = hdr.x;
    if (alive) { alive[2 * env] = e.flags & 1u; alive[2 * env + 1] = (e.flags >> 1) & 1u; }
    if (done) done[env] = (e.flags >> 2) & 1u;
    if (winner) winner[env] = (e.flags >> TRON_FLAG_WINNER_SHIFT) & 3u;
    if (ep_len) ep_len[env] = e.k;
}
__global__ void trail_import_kernel(unsigned char* recs, int n, int W, int H, const int8_t* __restrict__ tiles, const int8_t* heads, const uint8_t* alive,
                                    const uint8_t* done, const uint8_t* winner, const int32_t* ep_len) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    const size_t R = trail_record_bytes(W, H);
    unsigned char* rec = recs + (size_t)env * R;
    uint4 hdr = *(const uint4*)rec;
    if (tiles) {
        const int Hc = H + 2;
        const int8_t* t = tiles + (size_t)env * (W + 2) * (H + 2);
        unsigned short* ent = (unsigned short*)(rec + 16);
        int n1 = 0, n2 = 0;
        for (int r = 0; r < W; ++r)
            for (int c = 0; c < H; ++c) {
                const int v = t[(r + 1) * Hc + c + 1];
                if (v == TRON_TILE_P1_BODY || v == TRON_TILE_P1_SLIDE) ent[2 * (n1++) + 0] = TrailCells::pack(r, c, v == TRON_TILE_P1_SLIDE);
                else if (v == TRON_TILE_P2_BODY || v == TRON_TILE_P2_SLIDE) ent[2 * (n2++) + 1] = TrailCells::pack(r, c, v == TRON_TILE_P2_SLIDE);
            }
        hdr.z = (uint32_t)n1 | ((uint32_t)n2 << 16);
    }
    uint32_t f = hdr.y & 0xFFu, k = hdr.y >> 16;
    if (heads) hdr.x = ((const uint32_t*)heads)[env];
    if (alive) f = (f & ~3u) | (alive[2 * env] ? 1u : 0u) | (alive[2 * env + 1] ? 2u : 0u);
    if (done) f = (f & ~TRON_FLAG_DONE) | (done[env] ? TRON_FLAG_DONE : 0u);
    if (winner) f = (f & ~(3u << TRON_FLAG_WINNER_SHIFT)) | ((winner[env] & 3u) << TRON_FLAG_WINNER_SHIFT);
    if (ep_len) k = (uint32_t)ep_len[env] & 0xFFFFu;
    hdr.y = f | (k << 16);
    *(uint4*)rec = hdr;
}
int launch_trail_export(const void* recs, int n, int W, int H, int8_t* tiles, int8_t* heads, uint8_t* alive, uint8_t* done, uint8_t* winner,
                        int32_t* ep_len, cudaStream_t s) {
    trail_export_kernel<<<(n + 127) / 128, 128, 0, s>>>((const unsigned char*)recs, n, W, H, tiles, heads, alive, done, winner, ep_len);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}
int launch_trail_import(void* recs, int n, int W, int H, const int8_t* tiles, const int8_t* heads, const uint8_t* alive, const uint8_t* done,
                        const uint8_t* winner, const int32_t* ep_len, cudaStream_t s) {
    trail_import_kernel<<<(n + 127) / 128, 128, 0, s>>>((unsigned char*)recs, n, W, H, tiles, heads, alive, done, winner, ep_len);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

}  // namespace tron
