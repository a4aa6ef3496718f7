// step_bits10.cu -- fused tick + observation for the reference's config.py grid (10x10) on a BIT-PLANE state.
//
// TRON_LAYOUT_BITS10 keeps, per game, two 128-bit planes over the 100 interior cells (bit = p0*10 + p1):
//   occ  : a trail tile (body) lies here          own : the trail belongs to player 2
// Walls are implicit (the border ring), heads live in the 8-byte metadata and are overlaid when observations are
// encoded, so the state is 32 B instead of 144 B per game and a reset is "write zeros".  A thread keeps its game's
// planes in four 64-bit registers: the tick is bit tests/sets, no shared-memory round trip.  For the observation the
// owning thread expands its planes into an int8 Tile.value tile in shared memory (one funnel-shift + bit-spread multiply
// per 4 cells) and the CTA then runs the same PRMT encode + 16-byte streaming stores as the int8-layout kernel.
// HBM traffic per game-tick: 32 B read + 32 B written + observation planes + ~28 B metadata (668 B for bf16, 1 plane).
// Slide modes need a third plane (slide tiles 5/6) and are not supported by this layout (TRON_ERR_UNSUPPORTED).
#include "launch.h"
#include "step_kernels.cuh"

namespace tron {

constexpr int kBitsThreads = 128;  // up to one game per thread; p.G games per CTA (128, fewer when N is small so that all SMs get work)
constexpr int kW = 10, kHc = 12, kC = 144;

struct BitCells {
    unsigned long long occ_lo, occ_hi, own_lo, own_hi;
    __device__ __forceinline__ int get(int r, int c) const {
        if (r < 0 || c < 0 || r >= kW || c >= kW) return TRON_TILE_WALL;
        const int b = r * kW + c;
        const unsigned long long o = b < 64 ? occ_lo : occ_hi, w = b < 64 ? own_lo : own_hi;
        const int occ = (int)((o >> (b & 63)) & 1ull), own = (int)((w >> (b & 63)) & 1ull);
        return occ ? (own ? TRON_TILE_P2_BODY : TRON_TILE_P1_BODY) : TRON_TILE_EMPTY;
    }
    // only trail tiles are stored: heads are metadata (overlaid at encode time), border cells are implicit
    __device__ __forceinline__ void put(int r, int c, int tile) {
        if (tile != TRON_TILE_P1_BODY && tile != TRON_TILE_P2_BODY) return;
        if (r < 0 || c < 0 || r >= kW || c >= kW) return;
        const int b = r * kW + c;
        const unsigned long long m = 1ull << (b & 63);
        const bool p2 = tile == TRON_TILE_P2_BODY;
        if (b < 64) { occ_lo |= m; own_lo = p2 ? (own_lo | m) : (own_lo & ~m); }
        else { occ_hi |= m; own_hi = p2 ? (own_hi | m) : (own_hi & ~m); }
    }
    __device__ __forceinline__ void clear() { occ_lo = occ_hi = own_lo = own_hi = 0ull; }
};

// 4 bits -> 4 bytes (bit j -> byte j): the shifted copies at 0/7/14/21 never overlap, so no carries
__device__ __forceinline__ uint32_t spread4(uint32_t n) { return ((n & 0xFu) * 0x00204081u) & 0x01010101u; }
// Tile.value bytes of 4 cells from their occ / own bits: P1 body 1, P2 body 3
__device__ __forceinline__ uint32_t tiles4(uint32_t occ4, uint32_t own4) {
    const uint32_t o = spread4(occ4);
    return o + ((o & spread4(own4)) << 1);
}
__device__ __forceinline__ uint32_t row_bits(unsigned long long lo, unsigned long long hi, int r) {  // 10 bits of interior row r
    const int b = r * kW;
    if (b + kW <= 64) return (uint32_t)(lo >> b) & 0x3FFu;
    if (b >= 64) return (uint32_t)(hi >> (b - 64)) & 0x3FFu;
    return (uint32_t)((lo >> b) | (hi << (64 - b))) & 0x3FFu;
}

// expand one game's planes + heads into 144 Tile.value bytes (dst 4-byte aligned; shared or global memory)
__device__ __forceinline__ void expand_tile(const BitCells& g, const EnvState& e, int8_t* dst) {
    uint32_t* w = (uint32_t*)dst;
    w[0] = w[1] = w[2] = 0xFFFFFFFFu;  // border rows: WALL = -1
    w[33] = w[34] = w[35] = 0xFFFFFFFFu;
#pragma unroll
    for (int r = 0; r < kW; ++r) {
        const uint32_t occ = row_bits(g.occ_lo, g.occ_hi, r), own = row_bits(g.own_lo, g.own_hi, r);
        w[3 * (r + 1) + 0] = tiles4(occ << 1, own << 1) | 0x000000FFu;         // [wall, i0, i1, i2]
        w[3 * (r + 1) + 1] = tiles4(occ >> 3, own >> 3);                        // [i3 .. i6]
        w[3 * (r + 1) + 2] = tiles4((occ >> 7) & 7u, own >> 7) | 0xFF000000u;  // [i7, i8, i9, wall]
    }
    // heads overlay, P2 second so it wins a shared cell (reference game.py:205-214); positions may sit on the border
    dst[(e.r1 + 1) * kHc + e.c1 + 1] = TRON_TILE_P1_HEAD;
    dst[(e.r2 + 1) * kHc + e.c2 + 1] = TRON_TILE_P2_HEAD;
}

template <int OD, int LP, bool CP, int MODE>
__global__ void __launch_bounds__(kBitsThreads) step_bits10_kernel(const StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int8_t* tile = (int8_t*)smem_raw;  // [128][144] Tile.value bytes, only a staging area for the encode
    const int tid = threadIdx.x;
    const long long env0 = (long long)blockIdx.x * p.G;
    const int nG = (int)min((long long)p.G, (long long)p.N - env0);
    const bool owner = tid < nG;
    const long long env = env0 + tid;
    constexpr int CH = OD == TRON_F32 ? 4 : 8;

    BitCells g;
    g.clear();
    EnvState e = unpack_meta(make_uint2(0, 0));
    ulonglong2* planes = (ulonglong2*)p.grid;  // per game: {occ_lo, occ_hi}, {own_lo, own_hi}
    if (owner) {
        if (!(MODE == MODE_RESET && p.env_mask == nullptr)) {
            const ulonglong2 a = planes[2 * env], b = planes[2 * env + 1];
            g.occ_lo = a.x; g.occ_hi = a.y; g.own_lo = b.x; g.own_hi = b.y;
        }
        e = unpack_meta(p.meta[env]);
    }
    const int T = MODE == MODE_STEP ? p.T : 1;
    for (int t = 0; t < T; ++t) {
        if (MODE != MODE_OBSERVE && owner) {
            BoxRegs bx;
            if (env_tick<MODE, false>(g, p, e, env, t, tid, bx)) g.clear();
        }
        if (LP > 0 && (MODE == MODE_OBSERVE || p.obs_every_tick || t == T - 1)) {
            if (owner) expand_tile(g, e, tile + tid * kC);
            __syncthreads();
            encode_tile<kC, kBitsThreads, OD, LP, CP, CH>(tile, nG, env0, p, (MODE == MODE_STEP && p.obs_every_tick) ? t : 0);
            if (T > 1) __syncthreads();
        }
    }
    if (MODE != MODE_OBSERVE && owner) {
        planes[2 * env] = make_ulonglong2(g.occ_lo, g.occ_hi);
        planes[2 * env + 1] = make_ulonglong2(g.own_lo, g.own_hi);
        p.meta[env] = pack_meta(e);
    }
}

template <int OD, int LP, bool CP, int MODE>
static int launch_bits_one(const StepParams& p, cudaStream_t s) {
    const unsigned grid = (unsigned)(((long long)p.N + p.G - 1) / p.G);
    const size_t smem = LP > 0 ? (size_t)p.G * kC : 0;
    step_bits10_kernel<OD, LP, CP, MODE><<<grid, kBitsThreads, smem, s>>>(p);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}
template <int OD, int MODE>
static int launch_bits_enc(const StepParams& p, int enc_kind, cudaStream_t s) {
    switch (enc_kind) {
        case 1: return launch_bits_one<OD, 1, false, MODE>(p, s);
        case 2: return launch_bits_one<OD, 3, false, MODE>(p, s);
        case 3: return launch_bits_one<OD, 3, true, MODE>(p, s);
        default: return TRON_ERR_INVALID;
    }
}
int launch_step_bits10(const StepParams& p_in, int mode, int od, int enc_kind, cudaStream_t s) {
    StepParams p = p_in;
    p.G = tile_envs_small_grid(p.N);
    if (mode == MODE_RESET) return launch_bits_one<TRON_I8, 0, false, MODE_RESET>(p, s);
    if (mode == MODE_STEP && enc_kind == 0) return launch_bits_one<TRON_I8, 0, false, MODE_STEP>(p, s);
    if (mode == MODE_STEP) {
        if (od == TRON_BF16) return launch_bits_enc<TRON_BF16, MODE_STEP>(p, enc_kind, s);
        if (od == TRON_F32) return launch_bits_enc<TRON_F32, MODE_STEP>(p, enc_kind, s);
        if (od == TRON_I8) return launch_bits_enc<TRON_I8, MODE_STEP>(p, enc_kind, s);
    } else if (mode == MODE_OBSERVE) {
        if (od == TRON_BF16) return launch_bits_enc<TRON_BF16, MODE_OBSERVE>(p, enc_kind, s);
        if (od == TRON_F32) return launch_bits_enc<TRON_F32, MODE_OBSERVE>(p, enc_kind, s);
        if (od == TRON_I8) return launch_bits_enc<TRON_I8, MODE_OBSERVE>(p, enc_kind, s);
    }
    return TRON_ERR_INVALID;
}

// ---- export / import between the bit planes and Tile.value grids -------------------------------------------------
__global__ void bits10_export_kernel(const ulonglong2* __restrict__ planes, const uint2* __restrict__ meta, int n, int8_t* tiles) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    BitCells g;
    const ulonglong2 a = planes[2 * env], b = planes[2 * env + 1];
    g.occ_lo = a.x; g.occ_hi = a.y; g.own_lo = b.x; g.own_hi = b.y;
    expand_tile(g, unpack_meta(meta[env]), tiles + (size_t)env * kC);
}
// tiles -> planes: trail tiles (bodies; slide tiles are stored as bodies) set bits, everything else is implicit
__global__ void bits10_import_kernel(ulonglong2* planes, int n, const int8_t* __restrict__ tiles) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    BitCells g;
    g.clear();
    const int8_t* t = tiles + (size_t)env * kC;
    for (int r = 0; r < kW; ++r)
        for (int c = 0; c < kW; ++c) {
            const int v = t[(r + 1) * kHc + c + 1];
            if (v == TRON_TILE_P1_BODY || v == TRON_TILE_P1_SLIDE) g.put(r, c, TRON_TILE_P1_BODY);
            else if (v == TRON_TILE_P2_BODY || v == TRON_TILE_P2_SLIDE) g.put(r, c, TRON_TILE_P2_BODY);
        }
    planes[2 * env] = make_ulonglong2(g.occ_lo, g.occ_hi);
    planes[2 * env + 1] = make_ulonglong2(g.own_lo, g.own_hi);
}
int launch_bits10_export(const void* planes, const void* meta, int n, int8_t* tiles, cudaStream_t s) {
    bits10_export_kernel<<<(n + 127) / 128, 128, 0, s>>>((const ulonglong2*)planes, (const uint2*)meta, n, tiles);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}
int launch_bits10_import(void* planes, int n, const int8_t* tiles, cudaStream_t s) {
    bits10_import_kernel<<<(n + 127) / 128, 128, 0, s>>>((ulonglong2*)planes, n, tiles);
    return cudaGetLastError() == cudaSuccess ? TRON_OK : TRON_ERR_CUDA;
}

}  // namespace tron
