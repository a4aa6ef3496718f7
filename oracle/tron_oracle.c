/*
 * tron_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, one-game-at-a-time CPU restatement of the reference's TRON stepping path, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing under
 * deep-q-learning_tron_b200/ may call into this file.
 *
 * Parity pin: the reference (ckawoalt/Deep-Q-Learning_TRON) ships no tests, so this restatement is pinned
 * against outputs of the reference's own Python `Game` generated in the authoring container by
 * tests/golden/make_golden.py (KAT table, 2x1000-game SHA-256 digests, full trajectories, slide-mode
 * trajectories, pop_up planes) and checked by tests/test_oracle_golden.py.
 *
 * Reference citations are relative to Deep-Q-learning_TRON/ :
 *   tron/map.py:9-17 Tile, :43-48 Map.__init__, :67-84 color/state_for_player, :86-92 index offset
 *   tron/player.py:107-132 ACPlayer.get_direction / next_position
 *   tron/game.py:70-91 Game.__init__, :149-252 next_frame, :254-277 step
 *   tron/util.py:11-37 pop_up, :46-84 make_game, :87-94 get_reward
 *   DQN.py:81-132 ReplayMemory, :224-241 reward;  DDQN.py:90-110 eps-greedy, :167-203 ReplayBuffer
 *
 * It shares the argument structs of include/tron_b200.h (types only) so that a test can hand the same
 * arguments to the CUDA library (device pointers) and to this file (host pointers).
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tron_b200.h"

/* ------------------------------------------------------------------ Philox4x32-10 (Salmon et al. 2011) */
enum { TAG_ACTION = 1, TAG_SPAWN = 2, TAG_SLIDE = 3, TAG_EPS = 4, TAG_SAMPLE = 5, TAG_TEMPER = 6, TAG_FAIR = 7 };

static void philox4x32_10(uint64_t seed, uint64_t counter, uint64_t stream, uint32_t tag, uint32_t sub,
                          uint32_t out[4]) {
    uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = (uint32_t)stream;
    uint32_t c3 = ((uint32_t)(stream >> 32) & 0xFFFFu) | ((tag & 0xFFu) << 16) | ((sub & 0xFFu) << 24);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void oracle_philox(uint64_t seed, uint64_t counter, uint64_t stream, uint32_t tag, uint32_t sub, uint32_t* out) {
    philox4x32_10(seed, counter, stream, tag, sub, out);
}
static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* ------------------------------------------------------------------ geometry */
static inline int cells_of(int W, int H) { return (W + 2) * (H + 2); }
static inline size_t align256(size_t x) { return (x + 255u) & ~(size_t)255u; }
static inline int cell_index(int p0, int p1, int H) { return (p0 + 1) * (H + 2) + (p1 + 1); } /* map.py:86-92 */
static int8_t* grid_of(void* state) { return (int8_t*)state; }
static tron_meta* meta_of(void* state, int N, int W, int H) {
    return (tron_meta*)((char*)state + align256((size_t)N * cells_of(W, H)));
}
size_t oracle_state_bytes(int N, int W, int H) { return align256((size_t)N * cells_of(W, H)) + (size_t)N * 8u; }

/* box bounds of make_game(mode="fair") around point (px,py) (util.py:53-62): b = {lo1x,hi1x,lo1y,hi1y,lo2x,hi2x,lo2y,hi2y} */
void oracle_fair_bounds(int W, int H, int px, int py, int b[8]) {
    b[0] = px - 1 > 0 ? px - 1 : 0;          b[1] = px + 1 < W - 1 ? px + 1 : W - 1;
    b[2] = py - 1 > 0 ? py - 1 : 0;          b[3] = py + 1 < H - 1 ? py + 1 : H - 1;
    b[4] = W - 1 - b[1];                     b[5] = W - 1 - b[0];
    b[6] = H - 1 - b[3];                     b[7] = H - 1 - b[2];
}
/* spawn rule of make_game (util.py:46-84): draws inside the boxes, re-draw only (x1,y1) while the heads coincide */
static void rng_spawn(uint64_t seed, uint64_t counter, uint64_t env, int W, int H, int fair, int8_t s[4]) {
    int b[8] = {0, W - 1, 0, H - 1, 0, W - 1, 0, H - 1};
    uint32_t r[4];
    if (fair) {
        philox4x32_10(seed, counter, env, TAG_FAIR, 0, r);
        oracle_fair_bounds(W, H, (int)mulhi32(r[1], (uint32_t)W), (int)mulhi32(r[0], (uint32_t)H), b);
    }
    philox4x32_10(seed, counter, env, TAG_SPAWN, 0, r);
    int x1 = b[0] + (int)mulhi32(r[0], (uint32_t)(b[1] - b[0] + 1)), y1 = b[2] + (int)mulhi32(r[1], (uint32_t)(b[3] - b[2] + 1));
    int x2 = b[4] + (int)mulhi32(r[2], (uint32_t)(b[5] - b[4] + 1)), y2 = b[6] + (int)mulhi32(r[3], (uint32_t)(b[7] - b[6] + 1));
    uint32_t attempt = 0;
    while (x1 == x2 && y1 == y2) {
        if (++attempt >= 64) { x1 = x1 == b[0] ? b[1] : b[0]; if (x1 == x2 && y1 == y2) y1 = y1 == b[2] ? b[3] : b[2]; break; }
        philox4x32_10(seed, counter, env, TAG_SPAWN, attempt, r);
        x1 = b[0] + (int)mulhi32(r[0], (uint32_t)(b[1] - b[0] + 1)); y1 = b[2] + (int)mulhi32(r[1], (uint32_t)(b[3] - b[2] + 1));
    }
    s[0] = (int8_t)x1; s[1] = (int8_t)y1; s[2] = (int8_t)x2; s[3] = (int8_t)y2;
}
/* Game.__init__ draws weight x2 in [40,101] and degree in [-30,30] (game.py:83,87) */
static void rng_temper(uint64_t seed, uint64_t counter, uint64_t env, int8_t p[4]) {
    uint32_t r[4];
    philox4x32_10(seed, counter, env, TAG_TEMPER, 0, r);
    p[0] = (int8_t)(-30 + (int)mulhi32(r[2], 61u));
    p[1] = (int8_t)(40 + (int)mulhi32(r[0], 62u));
    p[2] = (int8_t)(40 + (int)mulhi32(r[1], 62u));
    p[3] = 0;
}

/* Game.__init__ (game.py:70-91) on one env */
static void fresh_game(int8_t* g, tron_meta* m, int W, int H, const int8_t s[4]) {
    const int R = W + 2, Cc = H + 2;
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < Cc; ++j)
            g[i * Cc + j] = (i == 0 || i == R - 1 || j == 0 || j == Cc - 1) ? TRON_TILE_WALL : TRON_TILE_EMPTY;
    g[cell_index(s[0], s[1], H)] = TRON_TILE_P1_HEAD;
    g[cell_index(s[2], s[3], H)] = TRON_TILE_P2_HEAD; /* P2 written second (game.py:90-91) */
    m->r1 = s[0]; m->c1 = s[1]; m->r2 = s[2]; m->c2 = s[3];
    m->flags = TRON_FLAG_ALIVE1 | TRON_FLAG_ALIVE2; m->reserved = 0; m->ep_len = 0;
}

int oracle_reset(void* state, int N, int W, int H, const int8_t* spawn, int spawn_mode, const uint8_t* mask, uint64_t seed,
                 uint64_t counter, uint64_t base) {
    int8_t* grid = grid_of(state); tron_meta* meta = meta_of(state, N, W, H); const int C = cells_of(W, H);
    for (int e = 0; e < N; ++e) {
        if (mask && !mask[e]) continue;
        int8_t s[4];
        if (spawn) memcpy(s, spawn + 4 * (size_t)e, 4); else rng_spawn(seed, counter, base + (uint64_t)e, W, H, spawn_mode, s);
        fresh_game(grid + (size_t)e * C, meta + e, W, H, s);
    }
    return 0;
}

/* [degree, weight_p] of the game an observation shows (Game.get_multy, game.py:137-139): extra[N][2][2] */
static void write_extra(const tron_step_args* a, int e) {
    if (!a->extra || !a->slide_params) return;
    const int8_t* sp = a->slide_params + 4 * (size_t)e;
    float* x = a->extra + 4 * (size_t)e;
    x[0] = (float)sp[0]; x[1] = (float)sp[1]; x[2] = (float)sp[0]; x[3] = (float)sp[2];
}
/* tron_reset_ex: like oracle_reset, plus Game.__init__'s three draws (game.py:83,87) for every game that is reset */
int oracle_reset_ex(const tron_step_args* a, const uint8_t* mask) {
    const int N = a->n_envs, W = a->width, H = a->height, C = cells_of(W, H);
    int8_t* grid = grid_of(a->state); tron_meta* meta = meta_of(a->state, N, W, H);
    for (int e = 0; e < N; ++e) {
        if (!mask || mask[e]) {
            int8_t s[4];
            const uint64_t env = a->env_id_base + (uint64_t)e;
            if (a->spawn) memcpy(s, a->spawn + 4 * (size_t)e, 4); else rng_spawn(a->seed, a->counter, env, W, H, a->spawn_mode, s);
            fresh_game(grid + (size_t)e * C, meta + e, W, H, s);
            if (a->slide_params) rng_temper(a->seed, a->counter, env, a->slide_params + 4 * (size_t)e);
        }
        write_extra(a, e);
    }
    return 0;
}

/* ------------------------------------------------------------------ observation encoding */
/* Map.color (map.py:67-81) as a table over Tile.value+1; lut6 = {empty, wall, own body, enemy body, own head, enemy head} */
static void color_table(const int8_t lut6[6], int player /*0|1*/, int8_t t[8]) {
    t[TRON_TILE_WALL + 1] = lut6[1];
    t[TRON_TILE_EMPTY + 1] = lut6[0];
    t[TRON_TILE_P1_BODY + 1] = t[TRON_TILE_P1_SLIDE + 1] = player == 0 ? lut6[2] : lut6[3];
    t[TRON_TILE_P2_BODY + 1] = t[TRON_TILE_P2_SLIDE + 1] = player == 0 ? lut6[3] : lut6[2];
    t[TRON_TILE_P1_HEAD + 1] = player == 0 ? lut6[4] : lut6[5];
    t[TRON_TILE_P2_HEAD + 1] = player == 0 ? lut6[5] : lut6[4];
}
static int planes_of(int enc) { return enc == TRON_ENC_LUT1 ? 1 : enc == TRON_ENC_POPUP3 ? 3 : enc == TRON_ENC_POPUP3_CONST ? 4 : 0; }
/* tab[2][lut_planes][8]; pop_up (util.py:11-37) is applied to the colour value */
int oracle_build_plane_tables(const int8_t lut6_in[6], int enc, int8_t* tab) {
    static const int8_t dflt[6] = {1, -1, -2, -3, 10, -10};
    int8_t lut6[6]; int allzero = 1;
    for (int i = 0; i < 6; ++i) allzero &= lut6_in[i] == 0;
    memcpy(lut6, allzero ? dflt : lut6_in, 6);
    const int LP = enc == TRON_ENC_LUT1 ? 1 : 3;
    for (int p = 0; p < 2; ++p) {
        int8_t col[8]; color_table(lut6, p, col);
        for (int t = 0; t < 8; ++t) {
            if (enc == TRON_ENC_LUT1) { tab[(p * LP + 0) * 8 + t] = col[t]; continue; }
            int o = col[t];
            tab[(p * LP + 0) * 8 + t] = (int8_t)(o == -1);
            tab[(p * LP + 1) * 8 + t] = (int8_t)(o == -2 ? 1 : o == 10 ? 10 : 0);
            tab[(p * LP + 2) * 8 + t] = (int8_t)(o == -3 ? 1 : o == -10 ? 10 : 0);
        }
    }
    return LP;
}
static uint16_t f32_to_bf16_rne(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40u);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static void store_elem(void* obs, int dtype, size_t i, float v) {
    if (dtype == TRON_F32) ((float*)obs)[i] = v;
    else if (dtype == TRON_BF16) ((uint16_t*)obs)[i] = f32_to_bf16_rne(v);
    else ((int8_t*)obs)[i] = (int8_t)v;
}
static void encode_env(const int8_t* g, int C, const int8_t* tab, int LP, int P, float cplane, void* obs, int dtype,
                       size_t env) {
    for (int p = 0; p < 2; ++p)
        for (int pl = 0; pl < P; ++pl) {
            size_t o = ((env * 2 + (size_t)p) * (size_t)P + (size_t)pl) * (size_t)C;
            for (int c = 0; c < C; ++c)
                store_elem(obs, dtype, o + (size_t)c, pl < LP ? (float)tab[(p * LP + pl) * 8 + ((g[c] + 1) & 7)] : cplane);
        }
}
int oracle_observe(const tron_step_args* a) {
    const int N = a->n_envs, W = a->width, H = a->height, C = cells_of(W, H), P = planes_of(a->obs_enc);
    if (!P || !a->obs) return TRON_ERR_INVALID;
    int8_t tab[2 * 3 * 8]; const int LP = oracle_build_plane_tables(a->lut, a->obs_enc, tab);
    const int8_t* grid = grid_of(a->state);
    for (int e = 0; e < N; ++e) { encode_env(grid + (size_t)e * C, C, tab, LP, P, a->const_plane, a->obs, a->obs_dtype, (size_t)e); write_extra(a, e); }
    return 0;
}

/* ------------------------------------------------------------------ one tick */
static int read_action(const void* actions, int dtype, size_t i) {
    if (dtype == TRON_U8) return ((const uint8_t*)actions)[i];
    if (dtype == TRON_I32) { int32_t v = ((const int32_t*)actions)[i]; return (v < 0 || v > 255) ? 255 : v; }
    int64_t v = ((const int64_t*)actions)[i]; return (v < 0 || v > 255) ? 255 : (int)v;
}
static const int DR[4] = {-1, 0, 1, 0}, DC[4] = {0, 1, 0, -1}; /* player.py:124-132 */

typedef struct { uint64_t f[TRON_STATS_FIELDS]; } stat_acc;

/* tick `t` of a (possibly multi-tick) call; per-tick arrays are offset by the caller */
static void step_tick(const tron_step_args* a, uint64_t counter, const void* actions, const int8_t* spawn,
                      const uint8_t* slide_tape, void* obs, float* reward, uint8_t* done_out, uint8_t* winner_out,
                      int32_t* eplen_out, stat_acc* st, int last_tick) {
    const int N = a->n_envs, W = a->width, H = a->height, C = cells_of(W, H), P = planes_of(a->obs_enc);
    int8_t* grid = grid_of(a->state); tron_meta* meta = meta_of(a->state, N, W, H);
    int8_t tab[2 * 3 * 8]; int LP = 0;
    if (P) LP = oracle_build_plane_tables(a->lut, a->obs_enc, tab);
    const int64_t ice_thr = (int64_t)((double)a->slide_rate * 16777216.0);

    for (int e = 0; e < N; ++e) {
        int8_t* g = grid + (size_t)e * C; tron_meta* m = meta + e;
        const uint64_t env = a->env_id_base + (uint64_t)e;
        float rw[2] = {0.f, 0.f}; uint8_t done = 0, winner = 0; int32_t fin = 0;
        int a1, a2;
        if (actions) { a1 = read_action(actions, a->action_dtype, 2 * (size_t)e); a2 = read_action(actions, a->action_dtype, 2 * (size_t)e + 1); }
        else {
            uint32_t r[4]; philox4x32_10(a->seed, counter, env, TAG_ACTION, 0, r); a1 = (int)(r[0] >> 30); a2 = (int)(r[1] >> 30);
            if (a->policy == TRON_POLICY_FREE_EPS && !(m->flags & TRON_FLAG_DONE)) { /* epsilon-greedy proxy of SURVEY 8d */
                const int64_t eps_thr = (int64_t)((double)a->policy_epsilon * 16777216.0);
                for (int i = 0; i < 2; ++i) {
                    int* act = i ? &a2 : &a1;
                    *act = (int)(r[i] & 3u);
                    if ((int64_t)(r[i] >> 8) <= eps_thr) continue;
                    const int hr = i ? m->r2 : m->r1, hc = i ? m->c2 : m->c1;
                    int fl[4], nf = 0;
                    for (int k = 0; k < 4; ++k) {
                        const int rr = hr + DR[k], cc = hc + DC[k];
                        if (rr >= 0 && cc >= 0 && rr < W && cc < H && g[cell_index(rr, cc, H)] == TRON_TILE_EMPTY) fl[nf++] = k;
                    }
                    if (nf) *act = fl[mulhi32(r[2 + i], (uint32_t)nf)];
                }
            }
        }

        if (m->flags & TRON_FLAG_DONE) { /* frozen: finished game without auto-reset */
            done = 1; winner = (uint8_t)((m->flags >> TRON_FLAG_WINNER_SHIFT) & 3u);
        } else if (a1 > 3 || a2 > 3) {
            st->f[TRON_STAT_BAD_ACTION]++;
        } else {
            int pos[2][2] = {{m->r1, m->c1}, {m->r2, m->c2}};
            int act[2] = {a1, a2};
            int alive[2] = {(m->flags & TRON_FLAG_ALIVE1) != 0, (m->flags & TRON_FLAG_ALIVE2) != 0};
            const int k = m->ep_len;
            /* game.py:155-156 both old heads become bodies before any move */
            g[cell_index(pos[0][0], pos[0][1], H)] = TRON_TILE_P1_BODY;
            g[cell_index(pos[1][0], pos[1][1], H)] = TRON_TILE_P2_BODY;
            uint32_t sr[4] = {0, 0, 0, 0};
            if (a->slide_mode == TRON_SLIDE_ICE || a->slide_mode == TRON_SLIDE_TEMPER) philox4x32_10(a->seed, counter, env, TAG_SLIDE, 0, sr);
            for (int i = 0; i < 2; ++i) { /* game.py:158-178 */
                pos[i][0] += DR[act[i]]; pos[i][1] += DC[act[i]];
                if (a->slide_mode != TRON_SLIDE_NONE && pos[i][0] >= 0 && pos[i][1] >= 0 && pos[i][0] < W && pos[i][1] < H &&
                    g[cell_index(pos[i][0], pos[i][1], H)] == TRON_TILE_EMPTY) {
                    int slip;
                    const int64_t mant = (int64_t)(sr[i] >> 8);
                    if (a->slide_mode == TRON_SLIDE_TAPE) slip = slide_tape[2 * (size_t)e + i] != 0;
                    else if (a->slide_mode == TRON_SLIDE_ICE) slip = mant <= ice_thr;
                    else { /* game.py:96-102: rate = -((degree-30)*0.6)/100 - (70-weight[i])/100, as exact integers */
                        const int8_t* sp = a->slide_params + 4 * (size_t)e;
                        const int64_t K = 6 * (30 - (int64_t)sp[0]) - 700 + 10 * (int64_t)sp[1 + i];
                        slip = mant * 1000 <= K * 16777216;
                    }
                    if (slip) {
                        g[cell_index(pos[i][0], pos[i][1], H)] = i == 0 ? TRON_TILE_P1_SLIDE : TRON_TILE_P2_SLIDE;
                        pos[i][0] += DR[act[i]]; pos[i][1] += DC[act[i]];
                    }
                }
            }
            for (int i = 0; i < 2; ++i) { /* game.py:205-214, P1 fully resolved before P2; head written in all cases */
                const int oob = pos[i][0] < 0 || pos[i][1] < 0 || pos[i][0] >= W || pos[i][1] >= H;
                const int ci = cell_index(pos[i][0], pos[i][1], H);
                if (oob || g[ci] != TRON_TILE_EMPTY) alive[i] = 0;
                g[ci] = i == 0 ? TRON_TILE_P1_HEAD : TRON_TILE_P2_HEAD;
            }
            const int n_alive = alive[0] + alive[1]; /* game.py:264-277 */
            if (n_alive <= 1) {
                done = 1;
                if (n_alive == 1 && (pos[0][0] != pos[1][0] || pos[0][1] != pos[1][1])) winner = alive[0] ? 1 : 2;
            }
            m->r1 = (int8_t)pos[0][0]; m->c1 = (int8_t)pos[0][1]; m->r2 = (int8_t)pos[1][0]; m->c2 = (int8_t)pos[1][1];
            m->ep_len = (uint16_t)(k + 1);
            m->flags = (uint8_t)((alive[0] ? TRON_FLAG_ALIVE1 : 0) | (alive[1] ? TRON_FLAG_ALIVE2 : 0) | (done ? TRON_FLAG_DONE : 0) |
                                 (winner << TRON_FLAG_WINNER_SHIFT));
            st->f[TRON_STAT_ENV_STEPS]++;
            if (!done) {
                rw[0] = rw[1] = a->reward_table.step_base + a->reward_table.step_per_tick * (float)k;
            } else {
                if (winner == 0) rw[0] = rw[1] = a->reward_table.draw;
                else { rw[winner - 1] = a->reward_table.win; rw[2 - winner] = a->reward_table.lose; }
                fin = k + 1;
                st->f[TRON_STAT_EPISODES]++; st->f[TRON_STAT_EP_TICKS] += (uint64_t)fin;
                st->f[winner == 0 ? TRON_STAT_DRAWS : winner == 1 ? TRON_STAT_P1_WINS : TRON_STAT_P2_WINS]++;
                if (a->auto_reset) { /* ACKTR.py:296-310: replaced by a fresh game; returned obs is the new game's */
                    /* DDQN.py:270-308 stores the finished game's last frame as next_state of the terminal transition */
                    if (P && a->obs_terminal) encode_env(g, C, tab, LP, P, a->const_plane, a->obs_terminal, a->obs_dtype, (size_t)e);
                    int8_t s[4];
                    if (spawn) memcpy(s, spawn + 4 * (size_t)e, 4); else rng_spawn(a->seed, counter, env, W, H, a->spawn_mode, s);
                    fresh_game(g, m, W, H, s);
                    if (a->slide_params) rng_temper(a->seed, counter, env, (int8_t*)a->slide_params + 4 * (size_t)e);
                }
            }
        }
        if (reward) { reward[2 * (size_t)e] = rw[0]; reward[2 * (size_t)e + 1] = rw[1]; }
        if (done_out) done_out[e] = done;
        if (winner_out) winner_out[e] = winner;
        if (eplen_out) eplen_out[e] = fin;
        if (P && obs) encode_env(g, C, tab, LP, P, a->const_plane, obs, a->obs_dtype, (size_t)e);
        if (last_tick) write_extra(a, e);
    }
}

static size_t dsize(int dt) { return dt == TRON_U8 || dt == TRON_I8 ? 1 : dt == TRON_BF16 ? 2 : dt == TRON_I64 ? 8 : 4; }

int oracle_step_many(const tron_step_args* a) {
    const int T = a->n_ticks > 0 ? a->n_ticks : 1;
    const size_t N = (size_t)a->n_envs, C = (size_t)cells_of(a->width, a->height), P = (size_t)planes_of(a->obs_enc);
    stat_acc st; memset(&st, 0, sizeof st);
    for (int t = 0; t < T; ++t) {
        const size_t tt = (size_t)t;
        const void* act = a->actions ? (const char*)a->actions + tt * N * 2 * dsize(a->action_dtype) : NULL;
        const int8_t* sp = a->spawn ? a->spawn + tt * N * 4 : NULL;
        const uint8_t* sl = a->slide_tape ? a->slide_tape + tt * N * 2 : NULL;
        void* obs = NULL;
        if (a->obs && P) obs = (a->obs_every_tick) ? (char*)a->obs + tt * N * 2 * P * C * dsize(a->obs_dtype) : (t == T - 1 ? a->obs : NULL);
        step_tick(a, a->counter + (a->counter_dev ? *a->counter_dev : 0) + (uint64_t)t, act, sp, sl, obs, a->reward ? a->reward + tt * N * 2 : NULL,
                  a->done ? a->done + tt * N : NULL, a->winner ? a->winner + tt * N : NULL,
                  a->ep_len_out ? a->ep_len_out + tt * N : NULL, &st, t == T - 1);
    }
    if (a->stats) for (int i = 0; i < TRON_STATS_FIELDS; ++i) a->stats[i] += st.f[i];
    return 0;
}
int oracle_step(const tron_step_args* a) {
    tron_step_args b = *a; b.n_ticks = 1; b.obs_every_tick = 1;
    return oracle_step_many(&b);
}

int oracle_export_grid(const void* state, int N, int W, int H, int8_t* tiles, int8_t* heads, uint8_t* alive,
                       uint8_t* done, uint8_t* winner, int32_t* ep_len) {
    const int C = cells_of(W, H); const tron_meta* meta = meta_of((void*)state, N, W, H);
    if (tiles) memcpy(tiles, state, (size_t)N * C);
    for (int e = 0; e < N; ++e) {
        const tron_meta* m = meta + e;
        if (heads) { heads[4 * e] = m->r1; heads[4 * e + 1] = m->c1; heads[4 * e + 2] = m->r2; heads[4 * e + 3] = m->c2; }
        if (alive) { alive[2 * e] = m->flags & 1u; alive[2 * e + 1] = (m->flags >> 1) & 1u; }
        if (done) done[e] = (m->flags >> 2) & 1u;
        if (winner) winner[e] = (m->flags >> TRON_FLAG_WINNER_SHIFT) & 3u;
        if (ep_len) ep_len[e] = m->ep_len;
    }
    return 0;
}
int oracle_import_grid(void* state, int N, int W, int H, const int8_t* tiles, const int8_t* heads,
                       const uint8_t* alive, const uint8_t* done, const uint8_t* winner, const int32_t* ep_len) {
    const int C = cells_of(W, H); tron_meta* meta = meta_of(state, N, W, H);
    if (tiles) memcpy(state, tiles, (size_t)N * C);
    for (int e = 0; e < N; ++e) {
        tron_meta* m = meta + e;
        if (heads) { m->r1 = heads[4 * e]; m->c1 = heads[4 * e + 1]; m->r2 = heads[4 * e + 2]; m->c2 = heads[4 * e + 3]; }
        uint8_t f = m->flags;
        if (alive) f = (uint8_t)((f & ~3u) | (alive[2 * e] ? 1u : 0u) | (alive[2 * e + 1] ? 2u : 0u));
        if (done) f = (uint8_t)((f & ~TRON_FLAG_DONE) | (done[e] ? TRON_FLAG_DONE : 0u));
        if (winner) f = (uint8_t)((f & ~(3u << TRON_FLAG_WINNER_SHIFT)) | ((winner[e] & 3u) << TRON_FLAG_WINNER_SHIFT));
        m->flags = f;
        if (ep_len) m->ep_len = (uint16_t)ep_len[e];
    }
    return 0;
}

/* ------------------------------------------------------------------ policies */
int oracle_random_actions(uint8_t* actions, int N, uint64_t seed, uint64_t counter, uint64_t base) {
    for (int e = 0; e < N; ++e) {
        uint32_t r[4]; philox4x32_10(seed, counter, base + (uint64_t)e, TAG_ACTION, 0, r);
        actions[2 * e] = (uint8_t)(r[0] >> 30); actions[2 * e + 1] = (uint8_t)(r[1] >> 30);
    }
    return 0;
}
/* DDQN.py:90-110: explore iff random() <= eps, else argmax (first maximum, like np.argmax / torch.argmax) */
int oracle_select_actions(const float* q, int n_rows, float eps, uint8_t* actions, uint64_t seed, uint64_t counter,
                          uint64_t base) {
    for (int i = 0; i < n_rows; ++i) {
        uint32_t r[4]; philox4x32_10(seed, counter, base + (uint64_t)i, TAG_EPS, 0, r);
        const float u = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
        int best = 0;
        for (int j = 1; j < 4; ++j) if (q[4 * (size_t)i + j] > q[4 * (size_t)i + best]) best = j;
        actions[i] = (uint8_t)(u <= eps ? (r[1] >> 30) : (uint32_t)best);
    }
    return 0;
}

/* ------------------------------------------------------------------ replay ring */
int oracle_replay_push(const replay_ring* ring, uint64_t cursor, const void* s, const void* s2, const uint8_t* action,
                       const float* reward, const uint8_t* done, int done_stride, int64_t n) {
    const size_t fb = (size_t)ring->frame_elems * dsize(ring->frame_dtype);
    for (int64_t i = 0; i < n; ++i) {
        const size_t slot = (size_t)((cursor + (uint64_t)i) % (uint64_t)ring->capacity);
        memcpy((char*)ring->state + slot * fb, (const char*)s + (size_t)i * fb, fb);
        memcpy((char*)ring->next_state + slot * fb, (const char*)s2 + (size_t)i * fb, fb);
        ring->action[slot] = action[i]; ring->reward[slot] = reward[i];
        ring->done[slot] = done[done_stride == 2 ? i / 2 : i];
    }
    return 0;
}
static float load_as_f32(const void* p, int dtype, size_t i) {
    if (dtype == TRON_F32) return ((const float*)p)[i];
    if (dtype == TRON_BF16) { uint32_t u = (uint32_t)((const uint16_t*)p)[i] << 16; float f; memcpy(&f, &u, 4); return f; }
    return (float)((const int8_t*)p)[i];
}
int oracle_replay_gather(const replay_ring* ring, const int64_t* idx, int64_t k, void* out_s, void* out_s2, int out_dtype,
                         int64_t* out_a, float* out_r, float* out_d) {
    const size_t F = (size_t)ring->frame_elems;
    for (int64_t i = 0; i < k; ++i) {
        const size_t slot = (size_t)idx[i];
        for (size_t j = 0; j < F; ++j) {
            store_elem(out_s, out_dtype, (size_t)i * F + j, load_as_f32(ring->state, ring->frame_dtype, slot * F + j));
            store_elem(out_s2, out_dtype, (size_t)i * F + j, load_as_f32(ring->next_state, ring->frame_dtype, slot * F + j));
        }
        out_a[i] = ring->action[slot]; out_r[i] = ring->reward[slot]; out_d[i] = (float)ring->done[slot];
    }
    return 0;
}
/* Uniform sampling WITHOUT replacement (random.sample, DDQN.py:193 / DQN.py:111-112) as a keyed pseudo-random permutation of
 * [0,size): idx[i] = pi(i), pi = 6-round balanced Feistel network on 2*hb bits (2^(2hb) >= size) with round keys from
 * Philox(seed; counter) and cycle-walking back into [0,size).  Same construction as the product's sampler (include/tron_b200.h). */
typedef struct { uint32_t key[6]; uint32_t mask; int hb; uint64_t size; } feistel_perm;
static feistel_perm feistel_make(uint64_t size, uint64_t seed, uint64_t counter) {
    feistel_perm f; uint32_t a[4], b[4];
    philox4x32_10(seed, counter, 0, TAG_SAMPLE, 0, a); philox4x32_10(seed, counter, 0, TAG_SAMPLE, 1, b);
    f.key[0] = a[0]; f.key[1] = a[1]; f.key[2] = a[2]; f.key[3] = a[3]; f.key[4] = b[0]; f.key[5] = b[1];
    int bits = 1;
    while (bits < 62 && (1ull << bits) < size) ++bits;
    f.hb = (bits + 1) >> 1;
    f.mask = f.hb >= 32 ? 0xFFFFFFFFu : ((1u << f.hb) - 1u);
    f.size = size;
    return f;
}
static uint32_t feistel_mix(uint32_t v) { v *= 0x85EBCA6Bu; v ^= v >> 13; v *= 0xC2B2AE35u; v ^= v >> 16; return v; }
static uint64_t feistel_apply(const feistel_perm* f, uint64_t i) {
    uint64_t x = i;
    do {
        uint32_t L = (uint32_t)(x >> f->hb) & f->mask, R = (uint32_t)x & f->mask;
        for (int r = 0; r < 6; ++r) { const uint32_t t = L ^ (feistel_mix(R ^ f->key[r]) & f->mask); L = R; R = t; }
        x = ((uint64_t)L << f->hb) | R;
    } while (x >= f->size);
    return x;
}
int oracle_replay_sample_indices(int64_t size, int64_t k, uint64_t seed, uint64_t counter, int64_t* idx) {
    if (k > size || k <= 0) return TRON_ERR_INVALID;
    const feistel_perm f = feistel_make((uint64_t)size, seed, counter);
    for (int64_t i = 0; i < k; ++i) idx[i] = (int64_t)feistel_apply(&f, (uint64_t)i);
    return 0;
}
/* frame-sharing ring (include/tron_b200.h replay_frames): transition u = (tick, row); state = frames[tick % S][row], next_state =
 * terminal[tick % S][row] if the env finished at that tick and terminal frames are kept, else frames[(tick+1) % S][row] */
int oracle_replay_frames_sample_gather(const replay_frames* fr, int64_t first_tick, int64_t n_ticks, int64_t k, uint64_t seed,
                                       uint64_t counter, void* out_s, void* out_s2, int out_dtype, int64_t* out_a, float* out_r,
                                       float* out_d, int64_t* out_idx) {
    const uint64_t total = (uint64_t)n_ticks * (uint64_t)fr->rows;
    if (k <= 0 || (uint64_t)k > total || n_ticks > fr->n_slots - 1) return TRON_ERR_INVALID;
    const feistel_perm f = feistel_make(total, seed, counter);
    const size_t F = (size_t)fr->frame_elems;
    for (int64_t i = 0; i < k; ++i) {
        const uint64_t u = feistel_apply(&f, (uint64_t)i);
        const int64_t tick = first_tick + (int64_t)(u / (uint64_t)fr->rows), r = (int64_t)(u % (uint64_t)fr->rows);
        const size_t slot = (size_t)(tick % fr->n_slots), nslot = (size_t)((tick + 1) % fr->n_slots);
        const uint8_t dn = fr->done[slot * (size_t)(fr->rows / 2) + (size_t)(r / 2)];
        const size_t a0 = (slot * (size_t)fr->rows + (size_t)r) * F, b0 = (nslot * (size_t)fr->rows + (size_t)r) * F;
        for (size_t j = 0; j < F; ++j) {
            store_elem(out_s, out_dtype, (size_t)i * F + j, load_as_f32(fr->frames, fr->frame_dtype, a0 + j));
            store_elem(out_s2, out_dtype, (size_t)i * F + j,
                       (dn && fr->terminal) ? load_as_f32(fr->terminal, fr->frame_dtype, a0 + j) : load_as_f32(fr->frames, fr->frame_dtype, b0 + j));
        }
        out_a[i] = fr->action[slot * (size_t)fr->rows + (size_t)r];
        out_r[i] = fr->reward[slot * (size_t)fr->rows + (size_t)r];
        out_d[i] = (float)dn;
        if (out_idx) out_idx[i] = tick * fr->rows + r;
    }
    return 0;
}

/* ------------------------------------------------------------------ cpu_baseline helper (bench.py only) */
/* Plays `ticks` ticks of n_envs auto-reset games with the on-device-equivalent random policy; returns env-steps. */
#include <time.h>
double oracle_bench_random(int n_envs, int W, int H, int ticks, int obs_dtype, int obs_enc, uint64_t seed, double* seconds) {
    tron_step_args a; memset(&a, 0, sizeof a);
    a.struct_size = sizeof a; a.n_envs = n_envs; a.width = W; a.height = H; a.obs_dtype = obs_dtype; a.obs_enc = obs_enc;
    a.auto_reset = 1; a.seed = seed; a.reward_table.step_base = -1.f; a.reward_table.win = 100.f; a.reward_table.lose = -100.f;
    const size_t C = (size_t)cells_of(W, H), P = (size_t)planes_of(obs_enc);
    a.state = malloc(oracle_state_bytes(n_envs, W, H));
    a.obs = P ? malloc((size_t)n_envs * 2 * P * C * dsize(obs_dtype)) : NULL;
    a.reward = (float*)malloc((size_t)n_envs * 8); a.done = (uint8_t*)malloc((size_t)n_envs); a.winner = (uint8_t*)malloc((size_t)n_envs);
    oracle_reset(a.state, n_envs, W, H, NULL, 0, NULL, seed, 0, 0);
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < ticks; ++t) { a.counter = (uint64_t)t + 1; oracle_step(&a); }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(a.state); free(a.obs); free(a.reward); free(a.done); free(a.winner);
    return (double)n_envs * (double)ticks;
}

/* ------------------------------------------------------------------ scripted opponent: depth-2 minimax + Voronoi heuristic
 * Restatement of MinimaxPlayer(2, voronoi).action (tron/minimax.py:58-310), quirks included:
 *  - the search runs on the TRANSPOSED observation of the moving player (minimax.py:298): gm[i0][i1] = obs[i1][i0];
 *  - heads are located with argmax / argmin over the flattened map (minimax.py:151-153,219-220);
 *  - get_shortest_path (minimax.py:64-86) is a FIFO over an ordered SET of (x, y, l) tuples: a cell is marked only when it is
 *    popped, so a cell can be queued again by a same-level neighbour with a larger l and its distance is then overwritten;
 *  - get_voronoi_value (minimax.py:88-123) counts cells literally as coded (e.g. enemy-body cells, value -3, count for player 1);
 *  - a node whose mover has no free neighbour keeps value 0 (minimax.py:233-234, TreeNode value initialised to 0);
 *  - ties at the root are broken by random.choice, a fully blocked root by random.randint(1,4) (minimax.py:234,267).
 * tie_mode 0: first best action / action 1 (what the golden fixtures use, with random.choice and randint patched accordingly);
 * tie_mode 1: Philox-uniform among the best actions. */
enum { TAG_MINIMAX = 8 };
#define MM_MAXC 256
typedef struct { int rows, cols; int16_t v[MM_MAXC]; } mm_map;  /* rows = H+2 (i0), cols = W+2 (i1) */

static int mm_arg(const mm_map* m, int want_max) {
    int best = 0;
    for (int i = 1; i < m->rows * m->cols; ++i)
        if (want_max ? m->v[i] > m->v[best] : m->v[i] < m->v[best]) best = i;
    return best;
}
static inline int mm_wrap(int i, int n) { return i < 0 ? i + n : (i >= n ? n - 1 : i); } /* numpy negative-index wrap; clamp above */
static inline int mm_at(const mm_map* m, int i0, int i1) { return mm_wrap(i0, m->rows) * m->cols + mm_wrap(i1, m->cols); }

static void mm_shortest_path(const mm_map* gm, int ind, int pl_mi, mm_map* dist) {
    static const int D0[4] = {0, 1, 0, -1}, D1[4] = {-1, 0, 1, 0}; /* (x,y-1) (x+1,y) (x,y+1) (x-1,y) */
    int16_t qx[4 * MM_MAXC + 8], qy[4 * MM_MAXC + 8], ql[4 * MM_MAXC + 8];
    int head = 0, tail = 0;
    *dist = *gm;
    qx[tail] = (int16_t)(ind / gm->cols); qy[tail] = (int16_t)(ind % gm->cols); ql[tail] = (int16_t)pl_mi; ++tail;
    while (head < tail) {
        const int x = qx[head], y = qy[head], l = ql[head]; ++head;
        dist->v[mm_at(dist, x, y)] = (int16_t)(l + pl_mi);
        for (int k = 0; k < 4; ++k) {
            const int nx = x + D0[k], ny = y + D1[k];
            if (dist->v[mm_at(dist, nx, ny)] != 1) continue;
            int dup = 0;
            for (int q = head; q < tail; ++q) dup |= (qx[q] == nx && qy[q] == ny && ql[q] == l + pl_mi);

            if (!dup && tail < 4 * MM_MAXC + 8) { qx[tail] = (int16_t)nx; qy[tail] = (int16_t)ny; ql[tail] = (int16_t)(l + pl_mi); ++tail; }
        }
    }
}
static int mm_voronoi(const mm_map* gm, int ind1, int ind2) {
    mm_map p1, p2;
    mm_shortest_path(gm, ind1, 1, &p1);
    mm_shortest_path(gm, ind2, -1, &p2);
    int a1 = 0, a2 = 0;
    for (int i = 0; i < gm->rows * gm->cols; ++i) {
        const int u = p1.v[i], w = p2.v[i];
        if (u == -1 || u == 2 || w == -2) continue;
        if (u != 1 && w == 1) a1++;
        else if (u == 1 && w != 1) a2++;
        else if (u + w < 0) a1++;
        else if (u + w > 0) a2++;
    }
    return a1 - a2;
}
static int mm_blocked(const mm_map* gm, int deo, int blocked[4]) {
    static const int D0[4] = {0, 1, 0, -1}, D1[4] = {-1, 0, 1, 0};
    const int ind = mm_arg(gm, deo == 1), x = ind / gm->cols, y = ind % gm->cols;
    int all = 1;
    for (int k = 0; k < 4; ++k) {
        const int v = gm->v[mm_at(gm, x + D0[k], y + D1[k])];
        blocked[k] = v != 1 ? (v == 10 ? 2 : 1) : 0;
        if (blocked[k] == 0) all = 0;
    }
    return all;
}
static void mm_next_map(const mm_map* gm, int action /*0..3*/, int deo, mm_map* out) {
    static const int D0[4] = {0, 1, 0, -1}, D1[4] = {-1, 0, 1, 0};
    const int ind = mm_arg(gm, deo == 1), x = ind / gm->cols, y = ind % gm->cols;
    *out = *gm;
    out->v[mm_at(out, x + D0[action], y + D1[action])] = (int16_t)(10 * deo);
    out->v[ind] = -1;
}
/* values of the root's children (INT32_MIN for unexpanded moves); returns 1 if the root is fully blocked */
/* ties[a]: what the depth-1 node of root move a draws from the global RNG when it finishes: -1 not expanded, 0 one randint (enemy
 * boxed in, minimax.py:233-234), L = one random.choice over the L minimising enemy moves (minimax.py:266-267) */
static int mm_root_values(const mm_map* gm, int value[4], int ties[4]) {
    int b0[4];
    for (int k = 0; k < 4; ++k) { value[k] = INT32_MIN; ties[k] = -1; }
    if (mm_blocked(gm, 1, b0)) return 1;
    for (int a = 0; a < 4; ++a) {
        if (b0[a] == 1) continue;
        mm_map m1; mm_next_map(gm, a, 1, &m1);
        int b1[4];
        if (mm_blocked(&m1, -1, b1)) { value[a] = 0; ties[a] = 0; continue; }
        int best = INT32_MAX, n_best = 0;
        for (int b = 0; b < 4; ++b) {
            if (b1[b] == 1) continue;
            mm_map m2; mm_next_map(&m1, b, -1, &m2);
            const int v = mm_voronoi(&m2, mm_arg(&m2, 1), mm_arg(&m2, 0));
            if (v < best) { best = v; n_best = 1; } else if (v == best) n_best++;
        }
        value[a] = best; ties[a] = n_best;
    }
    return 0;
}
/* action (0..3) of MinimaxPlayer(2) for `player` (1|2) in every env; values_out [N,4] optional */
int oracle_minimax_actions(const void* state, int N, int W, int H, int player, int tie_mode, uint64_t seed, uint64_t counter,
                           uint64_t base, uint8_t* actions, int32_t* values_out, int32_t* ties_out) {
    const int C = cells_of(W, H);
    if (C > MM_MAXC) return TRON_ERR_UNSUPPORTED;
    int8_t lut6[6] = {0, 0, 0, 0, 0, 0}, tab[2 * 3 * 8];
    oracle_build_plane_tables(lut6, TRON_ENC_LUT1, tab);
    const int8_t* grid = (const int8_t*)state;
    for (int e = 0; e < N; ++e) {
        mm_map gm; gm.rows = H + 2; gm.cols = W + 2;
        for (int r = 0; r < W + 2; ++r)       /* obs[r][c] -> gm[c][r] */
            for (int c = 0; c < H + 2; ++c)
                gm.v[c * (W + 2) + r] = tab[(player - 1) * 8 + ((grid[(size_t)e * C + r * (H + 2) + c] + 1) & 7)];
        int value[4], ties[4];
        uint32_t rnd[4];
        philox4x32_10(seed, counter, base + (uint64_t)e, TAG_MINIMAX, (uint32_t)player, rnd);
        int act;
        if (mm_root_values(&gm, value, ties)) {
            act = tie_mode ? (int)(rnd[0] >> 30) : 0;
        } else {
            int best = INT32_MIN, n = 0, list[4];
            for (int a = 0; a < 4; ++a) if (value[a] != INT32_MIN && value[a] > best) best = value[a];
            for (int a = 0; a < 4; ++a) if (value[a] == best) list[n++] = a;
            act = list[tie_mode ? (int)mulhi32(rnd[0], (uint32_t)n) : 0];
        }
        actions[e] = (uint8_t)act;
        if (values_out) for (int a = 0; a < 4; ++a) values_out[4 * e + a] = value[a];
        if (ties_out) for (int a = 0; a < 4; ++a) ties_out[4 * e + a] = ties[a];
    }
    return 0;
}
