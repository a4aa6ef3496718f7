// step_kernels.cuh -- the fused TRON tick: move, collision, winner/draw, trail write, auto-reset,
// reward and observation encoding for a tile of G consecutive games per CTA.
//
// Data flow per CTA (one tile = G games x C cells of int8 Tile.value, contiguous in HBM):
//   1. one elected thread pulls the whole tile into shared memory with a 1-D bulk (TMA-engine) copy
//      completing on an mbarrier, while the other threads fetch per-game metadata and actions;
//   2. thread-per-game runs the reference's tick on its game's cells in shared memory
//      (reference tron/game.py:149-277), emits reward/done/winner, decides auto-reset + spawn;
//   3. all threads rebuild reset games from a per-CTA template grid;
//   4. the final tile goes back to HBM as one bulk store while all threads encode both players'
//      observation planes straight from shared memory with 16-byte streaming stores (PRMT byte-LUT).
// HBM traffic per game-tick: C read + C written (grid) + 2*P*C*sizeof(obs) written + ~30 B metadata.
#pragma once
#include "common.cuh"

namespace tron {

enum : int { MODE_STEP = 0, MODE_OBSERVE = 1, MODE_RESET = 2 };

struct StepParams {
    int8_t* grid;
    uint2* meta;  // tron_meta, 8 bytes
    const void* actions;
    void* obs;
    float* reward;
    uint8_t* done;
    uint8_t* winner;
    int32_t* eplen;
    const int8_t* spawn;
    const uint8_t* slide_tape;
    int8_t* slide_params;
    const uint8_t* env_mask;  // MODE_RESET
    unsigned long long* stats;
    unsigned long long seed, counter, env_base;
    long long ice_thr;
    int N, W, H, Hc, C, G;
    int T, obs_every_tick, auto_reset, slide_mode, action_dtype;
    int P;  // planes written per player (lut planes + optional const plane)
    float r_base, r_tick, r_win, r_lose, r_draw, const_plane;
    PlaneTab tab[2][3];
};

__device__ __forceinline__ int read_action(const void* actions, int dtype, size_t i) {
    if (dtype == TRON_U8) return ((const uint8_t*)actions)[i];
    if (dtype == TRON_I32) { const int v = ((const int32_t*)actions)[i]; return (v < 0 || v > 255) ? 255 : v; }
    const long long v = ((const long long*)actions)[i];
    return (v < 0 || v > 255) ? 255 : (int)v;
}

// ---- observation element packing --------------------------------------------------------------
// 4 cells -> 4 encoded elements of dtype OD, returned as raw 32-bit words (1 word i8, 2 bf16, 4 f32)
template <int OD>
struct Enc4;
template <>
struct Enc4<TRON_BF16> {
    static constexpr int WORDS = 2;
    __device__ __forceinline__ static void run(const PlaneTab& t, uint32_t sel, uint32_t* o) {
        const uint32_t lo = lut4(t.lo0, t.lo1, sel), hi = lut4(t.hi0, t.hi1, sel);
        o[0] = __byte_perm(lo, hi, 0x5140);
        o[1] = __byte_perm(lo, hi, 0x7362);
    }
    __device__ __forceinline__ static void fill(float v, uint32_t* o) {
        const uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
        o[0] = o[1] = b | (b << 16);
    }
};
template <>
struct Enc4<TRON_F32> {
    static constexpr int WORDS = 4;
    __device__ __forceinline__ static void run(const PlaneTab& t, uint32_t sel, uint32_t* o) {
        const uint32_t lo = lut4(t.lo0, t.lo1, sel), hi = lut4(t.hi0, t.hi1, sel);
        o[0] = __byte_perm(lo, hi, 0x4044) & 0xFFFF0000u;  // bytes: [x, x, lo0, hi0] -> keep upper half
        o[1] = __byte_perm(lo, hi, 0x5144) & 0xFFFF0000u;
        o[2] = __byte_perm(lo, hi, 0x6244) & 0xFFFF0000u;
        o[3] = __byte_perm(lo, hi, 0x7344) & 0xFFFF0000u;
    }
    __device__ __forceinline__ static void fill(float v, uint32_t* o) { o[0] = o[1] = o[2] = o[3] = __float_as_uint(v); }
};
template <>
struct Enc4<TRON_I8> {
    static constexpr int WORDS = 1;
    __device__ __forceinline__ static void run(const PlaneTab& t, uint32_t sel, uint32_t* o) { o[0] = lut4(t.lo0, t.lo1, sel); }
    __device__ __forceinline__ static void fill(float v, uint32_t* o) {
        const uint32_t b = (uint32_t)(uint8_t)(int8_t)v;
        o[0] = b * 0x01010101u;
    }
};

template <int OD>
__device__ __forceinline__ int elem_size() { return OD == TRON_F32 ? 4 : OD == TRON_BF16 ? 2 : 1; }

// ---- the kernel -----------------------------------------------------------------------------
// C_T: cells per env at compile time (0 = runtime), NT threads, OD obs dtype, LP lut planes (0 = no obs),
// CP const plane, CH cells per encode item (8, 4 or 1; C % CH == 0), MODE.
template <int C_T, int NT, int OD, int LP, bool CP, int CH, int MODE>
__global__ void __launch_bounds__(NT) step_tile_kernel(const StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = C_T ? C_T : p.C;
    const int G = p.G;
    const int tid = threadIdx.x;
    const long long env0 = (long long)blockIdx.x * G;
    const int nG = (int)min((long long)G, (long long)p.N - env0);

    // shared layout: tile[G*C] | template[C] (16B padded) | hidx[G] (ushort2) | rflag[G] | mbarrier
    int8_t* tile = (int8_t*)smem_raw;
    const int tile_bytes_full = (G * C + 15) & ~15;
    int8_t* tmpl = tile + tile_bytes_full;
    ushort2* hidx = (ushort2*)(tmpl + ((C + 15) & ~15));
    uint8_t* rflag = (uint8_t*)(hidx + G);
    uint64_t* bar = (uint64_t*)(((uintptr_t)(rflag + G) + 7) & ~(uintptr_t)7);

    const int8_t* gsrc = p.grid + env0 * C;
    const uint32_t tile_bytes = (uint32_t)(nG * C);
    const bool bulk_ok = (tile_bytes & 15u) == 0 && ((((uintptr_t)gsrc) & 15u) == 0);
    const bool need_load = !(MODE == MODE_RESET && p.env_mask == nullptr);

    if (tid == 0 && bulk_ok && need_load) mbar_init(bar, 1);
    __syncthreads();
    if (need_load) {
        if (bulk_ok) {
            if (tid == 0) { mbar_expect_tx(bar, tile_bytes); bulk_g2s(tile, gsrc, tile_bytes, bar); }
        } else {
            for (uint32_t i = tid; i < tile_bytes; i += NT) tile[i] = gsrc[i];
        }
    }
    // template grid for resets (reference tron/map.py:43-48): border WALL, interior EMPTY
    if (MODE != MODE_OBSERVE) {
        const int Hc = p.Hc, Wc = p.W + 2;
        for (int c = tid; c < C; c += NT) {
            const int r = c / Hc, q = c - r * Hc;
            tmpl[c] = (r == 0 || r == Wc - 1 || q == 0 || q == Hc - 1) ? (int8_t)TRON_TILE_WALL : (int8_t)TRON_TILE_EMPTY;
        }
    }

    const bool owner = tid < nG;
    const long long env = env0 + tid;
    uint2 mraw = make_uint2(0, 0);
    if (owner && MODE != MODE_OBSERVE) mraw = p.meta[env];

    if (need_load && bulk_ok) mbar_wait(bar, 0);
    __syncthreads();

    const int T = MODE == MODE_STEP ? p.T : 1;
    for (int t = 0; t < T; ++t) {
        if (MODE != MODE_OBSERVE) {
            // ------------------------------------------------------------ phase 1: thread-per-game tick
            bool do_reset = false;
            if (owner) {
                int r1 = (int8_t)(mraw.x & 0xFF), c1 = (int8_t)((mraw.x >> 8) & 0xFF), r2 = (int8_t)((mraw.x >> 16) & 0xFF),
                    c2 = (int8_t)(mraw.x >> 24);
                uint32_t flags = mraw.y & 0xFFu;
                int k = (int)(mraw.y >> 16);
                const unsigned long long genv = p.env_base + (unsigned long long)env;
                const unsigned long long ctr = p.counter + (unsigned long long)t;
                const size_t tn = (size_t)t * (size_t)p.N + (size_t)env;
                char4 sp = make_char4(0, 0, 0, 0);
                if (MODE == MODE_RESET) {
                    do_reset = p.env_mask ? p.env_mask[env] != 0 : true;
                } else {
                    int a1, a2;
                    if (p.actions) {
                        a1 = read_action(p.actions, p.action_dtype, 2 * tn);
                        a2 = read_action(p.actions, p.action_dtype, 2 * tn + 1);
                    } else {
                        const uint4 r = philox(p.seed, ctr, genv, TAG_ACTION, 0);
                        a1 = (int)(r.x >> 30); a2 = (int)(r.y >> 30);
                    }
                    float rw0 = 0.f, rw1 = 0.f;
                    uint32_t done = 0, winner = 0;
                    int fin = 0;
                    bool bad = false, stepped = false;
                    if (flags & TRON_FLAG_DONE) {  // finished game without auto-reset stays frozen
                        done = 1; winner = (flags >> TRON_FLAG_WINNER_SHIFT) & 3u;
                    } else if (a1 > 3 || a2 > 3) {
                        bad = true;
                    } else {
                        stepped = true;
                        int8_t* g = tile + tid * C;
                        const int Hc = p.Hc;
                        // reference game.py:155-156: both old heads become bodies before any move
                        g[(r1 + 1) * Hc + c1 + 1] = TRON_TILE_P1_BODY;
                        g[(r2 + 1) * Hc + c2 + 1] = TRON_TILE_P2_BODY;
                        // reference player.py:124-132
                        const int dr1 = (a1 == 2) - (a1 == 0), dc1 = (a1 == 1) - (a1 == 3);
                        const int dr2 = (a2 == 2) - (a2 == 0), dc2 = (a2 == 1) - (a2 == 3);
                        r1 += dr1; c1 += dc1;
                        if (p.slide_mode != TRON_SLIDE_NONE) {  // reference game.py:163-178
                            uint4 sr = make_uint4(0, 0, 0, 0);
                            if (p.slide_mode >= TRON_SLIDE_ICE) sr = philox(p.seed, ctr, genv, TAG_SLIDE, 0);
                            char4 tp = make_char4(0, 0, 0, 0);
                            if (p.slide_mode == TRON_SLIDE_TEMPER) tp = ((const char4*)p.slide_params)[env];
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                int& rr = i ? r2 : r1; int& cc = i ? c2 : c1;
                                const int dr = i ? dr2 : dr1, dc = i ? dc2 : dc1;
                                if (i) { rr += dr; cc += dc; }
                                if (rr >= 0 && cc >= 0 && rr < p.W && cc < p.H && g[(rr + 1) * Hc + cc + 1] == TRON_TILE_EMPTY) {
                                    bool slip;
                                    const long long mant = (long long)((i ? sr.y : sr.x) >> 8);
                                    if (p.slide_mode == TRON_SLIDE_TAPE) slip = p.slide_tape[2 * tn + i] != 0;
                                    else if (p.slide_mode == TRON_SLIDE_ICE) slip = mant <= p.ice_thr;
                                    else {
                                        const long long K = 6 * (30 - (long long)tp.x) - 700 + 10 * (long long)(i ? tp.z : tp.y);
                                        slip = mant * 1000 <= K * 16777216;
                                    }
                                    if (slip) {
                                        g[(rr + 1) * Hc + cc + 1] = i ? TRON_TILE_P2_SLIDE : TRON_TILE_P1_SLIDE;
                                        rr += dr; cc += dc;
                                    }
                                }
                            }
                        } else {
                            r2 += dr2; c2 += dc2;
                        }
                        // reference game.py:205-214: P1 fully resolved before P2, head written in every case
                        bool al1 = flags & TRON_FLAG_ALIVE1, al2 = flags & TRON_FLAG_ALIVE2;
                        const int i1 = (r1 + 1) * Hc + c1 + 1;
                        if (r1 < 0 || c1 < 0 || r1 >= p.W || c1 >= p.H || g[i1] != TRON_TILE_EMPTY) al1 = false;
                        g[i1] = TRON_TILE_P1_HEAD;
                        const int i2 = (r2 + 1) * Hc + c2 + 1;
                        if (r2 < 0 || c2 < 0 || r2 >= p.W || c2 >= p.H || g[i2] != TRON_TILE_EMPTY) al2 = false;
                        g[i2] = TRON_TILE_P2_HEAD;
                        // reference game.py:264-277
                        const int n_alive = (int)al1 + (int)al2;
                        if (n_alive <= 1) {
                            done = 1;
                            if (n_alive == 1 && (r1 != r2 || c1 != c2)) winner = al1 ? 1u : 2u;
                        }
                        flags = (al1 ? TRON_FLAG_ALIVE1 : 0u) | (al2 ? TRON_FLAG_ALIVE2 : 0u) | (done ? TRON_FLAG_DONE : 0u) |
                                (winner << TRON_FLAG_WINNER_SHIFT);
                        if (!done) {
                            rw0 = rw1 = p.r_base + p.r_tick * (float)k;
                        } else {
                            if (winner == 0) rw0 = rw1 = p.r_draw;
                            else { rw0 = winner == 1 ? p.r_win : p.r_lose; rw1 = winner == 2 ? p.r_win : p.r_lose; }
                            fin = k + 1;
                            do_reset = p.auto_reset != 0;
                        }
                        k += 1;
                    }
                    if (p.reward) ((float2*)p.reward)[tn] = make_float2(rw0, rw1);
                    if (p.done) p.done[tn] = (uint8_t)done;
                    if (p.winner) p.winner[tn] = (uint8_t)winner;
                    if (p.eplen) p.eplen[tn] = fin;
                    if (p.stats) {  // warp-aggregated counters, striped over TRON_STATS_SLOTS rows
                        const unsigned am = __activemask();
                        const unsigned m_fin = __ballot_sync(am, fin > 0), m_w1 = __ballot_sync(am, fin > 0 && winner == 1),
                                       m_w2 = __ballot_sync(am, fin > 0 && winner == 2), m_bad = __ballot_sync(am, bad),
                                       m_step = __ballot_sync(am, stepped);
                        const unsigned ticks = __reduce_add_sync(am, (unsigned)fin);
                        if ((tid & 31) == (__ffs(am) - 1)) {
                            unsigned long long* s = p.stats + (size_t)(blockIdx.x % TRON_STATS_SLOTS) * TRON_STATS_FIELDS;
                            if (m_fin) {
                                atomicAdd(s + TRON_STAT_EPISODES, (unsigned long long)__popc(m_fin));
                                atomicAdd(s + TRON_STAT_P1_WINS, (unsigned long long)__popc(m_w1));
                                atomicAdd(s + TRON_STAT_P2_WINS, (unsigned long long)__popc(m_w2));
                                atomicAdd(s + TRON_STAT_DRAWS, (unsigned long long)(__popc(m_fin) - __popc(m_w1) - __popc(m_w2)));
                                atomicAdd(s + TRON_STAT_EP_TICKS, (unsigned long long)ticks);
                            }
                            if (m_bad) atomicAdd(s + TRON_STAT_BAD_ACTION, (unsigned long long)__popc(m_bad));
                            atomicAdd(s + TRON_STAT_ENV_STEPS, (unsigned long long)__popc(m_step));
                        }
                    }
                }
                if (do_reset) {  // fresh game (reference game.py:70-91, util.py:70-78)
                    if (p.spawn) sp = ((const char4*)p.spawn)[tn];
                    else sp = rng_spawn(p.seed, ctr, genv, p.W, p.H);
                    r1 = sp.x; c1 = sp.y; r2 = sp.z; c2 = sp.w;
                    flags = TRON_FLAG_ALIVE1 | TRON_FLAG_ALIVE2; k = 0;
                    hidx[tid] = make_ushort2((unsigned short)((r1 + 1) * p.Hc + c1 + 1), (unsigned short)((r2 + 1) * p.Hc + c2 + 1));
                    if (MODE == MODE_STEP && p.slide_mode == TRON_SLIDE_TEMPER && p.slide_params)
                        ((char4*)p.slide_params)[env] = rng_temper(p.seed, ctr, genv);
                }
                mraw.x = (uint32_t)(uint8_t)r1 | ((uint32_t)(uint8_t)c1 << 8) | ((uint32_t)(uint8_t)r2 << 16) | ((uint32_t)(uint8_t)c2 << 24);
                mraw.y = flags | ((uint32_t)k << 16);
                rflag[tid] = do_reset ? 1 : 0;
            }
            __syncthreads();
            // ------------------------------------------------------------ phase 2: rebuild reset games
            {
                constexpr int V = (C_T != 0 && C_T % 16 == 0) ? 16 : 4;  // template copy granularity (C % 4 == 0 or scalar)
                if ((C % V) == 0) {
                    const int per = C / V;
                    for (int it = tid; it < nG * per; it += NT) {
                        const int e = it / per, ch = it - e * per;
                        if (!rflag[e]) continue;
                        if (V == 16) ((uint4*)(tile + e * C))[ch] = ((const uint4*)tmpl)[ch];
                        else ((uint32_t*)(tile + e * C))[ch] = ((const uint32_t*)tmpl)[ch];
                        const ushort2 h = hidx[e];
                        if ((int)h.x / V == ch) tile[e * C + h.x] = TRON_TILE_P1_HEAD;
                        if ((int)h.y / V == ch) tile[e * C + h.y] = TRON_TILE_P2_HEAD;
                    }
                } else {
                    for (int it = tid; it < nG * C; it += NT) {
                        const int e = it / C, c = it - e * C;
                        if (!rflag[e]) continue;
                        const ushort2 h = hidx[e];
                        tile[it] = c == h.y ? (int8_t)TRON_TILE_P2_HEAD : c == h.x ? (int8_t)TRON_TILE_P1_HEAD : tmpl[c];
                    }
                }
            }
            if (t == T - 1) fence_proxy_async();
            __syncthreads();
            // ------------------------------------------------------------ write the tile back (last tick only)
            if (t == T - 1) {
                int8_t* gdst = p.grid + env0 * C;
                if (bulk_ok) {
                    if (tid == 0) { bulk_s2g(gdst, tile, tile_bytes); bulk_commit(); }
                } else {
                    for (uint32_t i = tid; i < tile_bytes; i += NT) gdst[i] = tile[i];
                }
                if (owner) p.meta[env] = mraw;
            }
        }
        // ---------------------------------------------------------------- phase 3: observation planes
        if (LP > 0 && (MODE == MODE_OBSERVE || p.obs_every_tick || t == T - 1)) {
            constexpr int ES = OD == TRON_F32 ? 4 : OD == TRON_BF16 ? 2 : 1;
            const int P = p.P;
            const size_t tick_off = (MODE == MODE_STEP && p.obs_every_tick) ? (size_t)t * (size_t)p.N * 2 * (size_t)P * (size_t)C * ES : 0;
            char* obase = (char*)p.obs + tick_off;
            if constexpr (CH >= 4) {
                const int per = C / CH;
                for (int it = tid; it < nG * per; it += NT) {
                    const int e = it / per, ch = it - e * per;
                    const int8_t* cells = tile + e * C + ch * CH;
                    uint32_t sel[CH / 4];
                    if constexpr (CH == 8) { const uint2 w = *(const uint2*)cells; sel[0] = cell_selector(w.x); sel[1] = cell_selector(w.y); }
                    else { sel[0] = cell_selector(*(const uint32_t*)cells); }
                    char* row0 = obase + ((size_t)(env0 + e) * 2 * P * C + (size_t)ch * CH) * ES;
#pragma unroll
                    for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
                        for (int q = 0; q < LP + (CP ? 1 : 0); ++q) {
                            uint32_t o[(CH / 4) * Enc4<OD>::WORDS];
#pragma unroll
                            for (int h = 0; h < CH / 4; ++h) {
                                if (q < LP) Enc4<OD>::run(p.tab[pl][q < LP ? q : 0], sel[h], o + h * Enc4<OD>::WORDS);
                                else Enc4<OD>::fill(p.const_plane, o + h * Enc4<OD>::WORDS);
                            }
                            char* dst = row0 + (size_t)(pl * P + q) * C * ES;
                            constexpr int NW = (CH / 4) * Enc4<OD>::WORDS;
                            if (NW == 8) { st_cs((uint4*)dst, make_uint4(o[0], o[1], o[2], o[3])); st_cs((uint4*)dst + 1, make_uint4(o[4 % NW], o[5 % NW], o[6 % NW], o[7 % NW])); }
                            else if (NW == 4) st_cs((uint4*)dst, make_uint4(o[0], o[1 % NW], o[2 % NW], o[3 % NW]));
                            else if (NW == 2) st_cs((uint2*)dst, make_uint2(o[0], o[1 % NW]));
                            else *(uint32_t*)dst = o[0];
                        }
                    }
                }
            } else {  // scalar fallback for odd cell counts
                for (int it = tid; it < nG * C; it += NT) {
                    const int e = it / C, c = it - e * C;
                    const uint32_t sel = (uint32_t)tile[it] & 7u;
                    for (int pl = 0; pl < 2; ++pl)
                        for (int q = 0; q < LP + (CP ? 1 : 0); ++q) {
                            uint32_t o[Enc4<OD>::WORDS];
                            if (q < LP) Enc4<OD>::run(p.tab[pl][q < LP ? q : 0], sel, o); else Enc4<OD>::fill(p.const_plane, o);
                            char* dst = obase + ((size_t)((env0 + e) * 2 + pl) * P + q) * C * ES + (size_t)c * ES;
                            if (ES == 4) *(uint32_t*)dst = o[0];
                            else if (ES == 2) *(uint16_t*)dst = (uint16_t)o[0];
                            else *(uint8_t*)dst = (uint8_t)o[0];
                        }
                }
            }
        }
        if (T > 1) __syncthreads();  // next tick's phase 1 rewrites the tile
    }
    if (MODE != MODE_OBSERVE && bulk_ok && tid == 0) bulk_wait_read_all();  // tile must outlive the bulk store's read
}

}  // namespace tron
