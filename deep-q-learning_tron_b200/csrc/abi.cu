// abi.cu -- the extern "C" boundary declared in include/tron_b200.h: argument validation, kernel
// dispatch and the host-buffer front end.  No torch types, no exceptions, no CPU fallback.
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "launch.h"
#include "step_kernels.cuh"

using namespace tron;

namespace {

inline size_t align256(size_t x) { return (x + 255u) & ~(size_t)255u; }
inline bool geometry_ok(int n, int w, int h) { return n > 0 && w >= 2 && h >= 2 && w <= 126 && h <= 126; }
inline bool layout_known(int layout) { return layout == TRON_LAYOUT_TILE8 || layout == TRON_LAYOUT_BITS10 || layout == TRON_LAYOUT_TRAIL; }
inline bool layout_ok(int layout, int w, int h) { return layout == TRON_LAYOUT_TILE8 || layout == TRON_LAYOUT_TRAIL || (layout == TRON_LAYOUT_BITS10 && w == 10 && h == 10); }
// bytes of grid state per game
inline size_t grid_stride(int layout, int w, int h) {
    return layout == TRON_LAYOUT_BITS10 ? 32u : layout == TRON_LAYOUT_TRAIL ? trail_record_bytes_host(w, h) : (size_t)(w + 2) * (size_t)(h + 2);
}
inline int planes_of(int enc) { return enc == TRON_ENC_LUT1 ? 1 : enc == TRON_ENC_POPUP3 ? 3 : enc == TRON_ENC_POPUP3_CONST ? 4 : 0; }
inline int enc_kind_of(int enc) { return enc == TRON_ENC_LUT1 ? 1 : enc == TRON_ENC_POPUP3 ? 2 : enc == TRON_ENC_POPUP3_CONST ? 3 : 0; }

uint16_t bf16_bits_of_int8(int v) {  // every int8 is exactly representable in bf16
    float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    return (uint16_t)(u >> 16);
}

// device lookup tables indexed by (tile & 7): WALL(-1) -> 7, EMPTY 0, bodies/heads/slides 1..6
void build_device_tables(const int8_t lut6[6], int enc, int obs_dtype, PlaneTab out[2][3]) {
    int8_t tab[2 * 3 * 8];
    const int LP = tron_build_plane_tables(lut6, enc, tab);
    for (int p = 0; p < 2; ++p)
        for (int q = 0; q < 3; ++q) {
            uint8_t lo[8] = {0}, hi[8] = {0};
            if (q < LP)
                for (int tile = -1; tile <= 6; ++tile) {
                    const int v = tab[(p * LP + q) * 8 + (tile + 1)];
                    const int di = tile & 7;
                    if (obs_dtype == TRON_I8) { lo[di] = (uint8_t)(int8_t)v; }
                    else { const uint16_t b = bf16_bits_of_int8(v); lo[di] = (uint8_t)(b & 0xFF); hi[di] = (uint8_t)(b >> 8); }
                }
            memcpy(&out[p][q].lo0, lo, 4); memcpy(&out[p][q].lo1, lo + 4, 4);
            memcpy(&out[p][q].hi0, hi, 4); memcpy(&out[p][q].hi1, hi + 4, 4);
        }
}

// Fill kernel parameters from the public argument block.  grid/meta may be overridden (chunked host path).
int fill_params(const tron_step_args* a, int mode, StepParams& p) {
    if (!a || a->struct_size != sizeof(tron_step_args)) return TRON_ERR_INVALID;
    if (!geometry_ok(a->n_envs, a->width, a->height) || !a->state) return TRON_ERR_INVALID;
    if (!layout_known(a->layout)) return TRON_ERR_INVALID;
    if (!layout_ok(a->layout, a->width, a->height)) return TRON_ERR_UNSUPPORTED;
    if (a->layout == TRON_LAYOUT_BITS10 && mode == MODE_STEP && a->slide_mode != TRON_SLIDE_NONE) return TRON_ERR_UNSUPPORTED;
    if (((uintptr_t)a->state & 15u) != 0) return TRON_ERR_ALIGN;
    memset(&p, 0, sizeof p);
    p.N = a->n_envs; p.W = a->width; p.H = a->height; p.Hc = a->height + 2; p.C = (a->width + 2) * (a->height + 2);
    p.layout = a->layout;
    p.grid = (int8_t*)a->state;
    p.meta = (uint2*)((char*)a->state + align256((size_t)p.N * grid_stride(a->layout, a->width, a->height)));
    p.boxes = (uint2*)((char*)p.meta + align256((size_t)p.N * sizeof(tron_meta)));
    p.T = 1; p.obs_every_tick = 1;
    const int planes = planes_of(a->obs_enc);
    if (a->obs_enc != TRON_ENC_NONE) {
        if (!planes || !a->obs) return TRON_ERR_INVALID;
        if (a->obs_dtype != TRON_BF16 && a->obs_dtype != TRON_F32 && a->obs_dtype != TRON_I8) return TRON_ERR_INVALID;
        if (((uintptr_t)a->obs & 15u) != 0) return TRON_ERR_ALIGN;
        p.obs = a->obs; p.P = planes; p.const_plane = a->const_plane;
        build_device_tables(a->lut, a->obs_enc, a->obs_dtype, p.tab);
    } else if (mode == MODE_OBSERVE) {
        return TRON_ERR_INVALID;
    }
    if (mode == MODE_STEP) {
        if (a->actions && a->action_dtype != TRON_U8 && a->action_dtype != TRON_I32 && a->action_dtype != TRON_I64) return TRON_ERR_INVALID;
        if (a->slide_mode < TRON_SLIDE_NONE || a->slide_mode > TRON_SLIDE_TEMPER) return TRON_ERR_INVALID;
        if (a->slide_mode == TRON_SLIDE_TAPE && !a->slide_tape) return TRON_ERR_INVALID;
        if (a->slide_mode == TRON_SLIDE_TEMPER && !a->slide_params) return TRON_ERR_INVALID;
        p.actions = a->actions; p.action_dtype = a->action_dtype;
        p.reward = a->reward; p.done = a->done; p.winner = a->winner; p.eplen = a->ep_len_out;
        p.spawn = a->spawn; p.slide_tape = a->slide_tape; p.slide_params = a->slide_params; p.stats = (unsigned long long*)a->stats;
        p.auto_reset = a->auto_reset; p.slide_mode = a->slide_mode;
        if (a->spawn_mode != TRON_SPAWN_UNIFORM && a->spawn_mode != TRON_SPAWN_FAIR) return TRON_ERR_INVALID;
        p.ice_thr = (long long)((double)a->slide_rate * 16777216.0);
        if (a->policy != TRON_POLICY_UNIFORM && a->policy != TRON_POLICY_FREE_EPS) return TRON_ERR_INVALID;
        p.eps_thr = a->policy == TRON_POLICY_FREE_EPS ? (long long)((double)a->policy_epsilon * 16777216.0) : -1;
        p.r_base = a->reward_table.step_base; p.r_tick = a->reward_table.step_per_tick;
        p.r_win = a->reward_table.win; p.r_lose = a->reward_table.lose; p.r_draw = a->reward_table.draw;
    }
    p.seed = a->seed; p.counter = a->counter; p.env_base = a->env_id_base; p.spawn_mode = a->spawn_mode;
    p.counter_dev = (const unsigned long long*)a->counter_dev;
    return TRON_OK;
}

long long g_sparse_min_cells = 1024;  // TRON_OPT_SPARSE_MIN_CELLS
}  // namespace
namespace tron { extern long long g_tile_bytes; }
namespace {

int dispatch(StepParams& p, int mode, int obs_dtype, int obs_enc, cudaStream_t s) {
    const int kind = enc_kind_of(obs_enc);
    if (p.layout == TRON_LAYOUT_BITS10) return launch_step_bits10(p, mode, obs_dtype, kind, s);
    if (p.layout == TRON_LAYOUT_TRAIL) return (kind == 0 || mode == MODE_RESET) ? launch_step_trail(p, mode, s) : launch_step_trail_obs(p, mode, obs_dtype, kind, s);
    if (mode == MODE_STEP && kind == 0 && p.C >= g_sparse_min_cells) return launch_step_sparse(p, s);
    if (p.C == 144 && p.Hc == 12) { p.G = tile_envs_c144(p.N); return launch_step_c144(p, mode, obs_dtype, kind, s); }
    p.G = tile_envs_generic(p.C);
    return launch_step_generic(p, mode, obs_dtype, kind, s);
}

}  // namespace

extern "C" {

int tron_abi_version(void) { return TRON_B200_ABI_VERSION; }

const char* tron_status_string(int status) {
    switch (status) {
        case TRON_OK: return "ok";
        case TRON_ERR_INVALID: return "invalid argument";
        case TRON_ERR_UNSUPPORTED: return "unsupported";
        case TRON_ERR_CUDA: return "CUDA error (no usable device or launch failure); there is no CPU fallback";
        case TRON_ERR_ALIGN: return "misaligned pointer";
        default: return "unknown status";
    }
}

int tron_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return TRON_ERR_CUDA; }
    return n;
}

int tron_cells_per_env(int width, int height) { return (width + 2) * (height + 2); }
int tron_enc_planes(int obs_enc) { return planes_of(obs_enc); }
int tron_dtype_size(int dtype) { return tron_elem(dtype); }

int tron_state_offsets(int n_envs, int width, int height, int layout, size_t* grid_off, size_t* meta_off, size_t* boxes_off) {
    if (!geometry_ok(n_envs, width, height) || !layout_known(layout)) return TRON_ERR_INVALID;
    if (!layout_ok(layout, width, height)) return TRON_ERR_UNSUPPORTED;
    const size_t mo = align256((size_t)n_envs * grid_stride(layout, width, height));
    if (grid_off) *grid_off = 0;
    if (meta_off) *meta_off = mo;
    if (boxes_off) *boxes_off = mo + align256((size_t)n_envs * sizeof(tron_meta));
    return TRON_OK;
}
int tron_state_bytes(int n_envs, int width, int height, int layout, size_t* total_bytes) {
    size_t bo = 0;
    const int rc = tron_state_offsets(n_envs, width, height, layout, nullptr, nullptr, &bo);
    if (rc != TRON_OK || !total_bytes) return rc != TRON_OK ? rc : TRON_ERR_INVALID;
    *total_bytes = bo + (size_t)n_envs * 8u;
    return TRON_OK;
}
int tron_set_option(int option, int64_t value) {
    if (option == TRON_OPT_SPARSE_MIN_CELLS && value >= 0) { g_sparse_min_cells = value; return TRON_OK; }
    if (option == TRON_OPT_TILE_BYTES && value >= 1024 && value <= 200 * 1024) { tron::g_tile_bytes = value; return TRON_OK; }
    return TRON_ERR_INVALID;
}

// reference tron/map.py:67-81 (colour table) and tron/util.py:11-37 (pop_up applied to the colour value)
int tron_build_plane_tables(const int8_t lut6_in[6], int obs_enc, int8_t* tab) {
    static const int8_t dflt[6] = {1, -1, -2, -3, 10, -10};
    if (!lut6_in || !tab || !planes_of(obs_enc)) return TRON_ERR_INVALID;
    int8_t lut6[6];
    bool allzero = true;
    for (int i = 0; i < 6; ++i) allzero = allzero && lut6_in[i] == 0;
    memcpy(lut6, allzero ? dflt : lut6_in, 6);
    const int LP = obs_enc == TRON_ENC_LUT1 ? 1 : 3;
    for (int p = 0; p < 2; ++p) {
        int8_t col[8];
        col[TRON_TILE_WALL + 1] = lut6[1];
        col[TRON_TILE_EMPTY + 1] = lut6[0];
        col[TRON_TILE_P1_BODY + 1] = col[TRON_TILE_P1_SLIDE + 1] = p == 0 ? lut6[2] : lut6[3];
        col[TRON_TILE_P2_BODY + 1] = col[TRON_TILE_P2_SLIDE + 1] = p == 0 ? lut6[3] : lut6[2];
        col[TRON_TILE_P1_HEAD + 1] = p == 0 ? lut6[4] : lut6[5];
        col[TRON_TILE_P2_HEAD + 1] = p == 0 ? lut6[5] : lut6[4];
        for (int t = 0; t < 8; ++t) {
            const int o = col[t];
            if (LP == 1) { tab[p * 8 + t] = (int8_t)o; continue; }
            tab[(p * 3 + 0) * 8 + t] = (int8_t)(o == -1);
            tab[(p * 3 + 1) * 8 + t] = (int8_t)(o == -2 ? 1 : o == 10 ? 10 : 0);
            tab[(p * 3 + 2) * 8 + t] = (int8_t)(o == -3 ? 1 : o == -10 ? 10 : 0);
        }
    }
    return LP;
}

int tron_reset(void* state, int n_envs, int width, int height, int layout, const int8_t* spawn, int spawn_mode, const uint8_t* env_mask,
               uint64_t seed, uint64_t counter, uint64_t env_id_base, tron_stream_t stream) {
    if (spawn_mode != TRON_SPAWN_UNIFORM && spawn_mode != TRON_SPAWN_FAIR) return TRON_ERR_INVALID;
    tron_step_args a;
    memset(&a, 0, sizeof a);
    a.struct_size = sizeof a; a.n_envs = n_envs; a.width = width; a.height = height; a.layout = layout; a.state = state;
    a.seed = seed; a.counter = counter; a.env_id_base = env_id_base; a.spawn_mode = spawn_mode;
    StepParams p;
    const int rc = fill_params(&a, MODE_RESET, p);
    if (rc != TRON_OK) return rc;
    p.spawn = spawn; p.env_mask = env_mask;
    return dispatch(p, MODE_RESET, TRON_I8, TRON_ENC_NONE, (cudaStream_t)stream);
}

int tron_step(const tron_step_args* args, tron_stream_t stream) {
    StepParams p;
    const int rc = fill_params(args, MODE_STEP, p);
    if (rc != TRON_OK) return rc;
    return dispatch(p, MODE_STEP, args->obs_dtype, args->obs_enc, (cudaStream_t)stream);
}

int tron_step_many(const tron_step_args* args, tron_stream_t stream) {
    StepParams p;
    const int rc = fill_params(args, MODE_STEP, p);
    if (rc != TRON_OK) return rc;
    if (args->n_ticks < 1) return TRON_ERR_INVALID;
    p.T = args->n_ticks; p.obs_every_tick = args->obs_every_tick;
    return dispatch(p, MODE_STEP, args->obs_dtype, args->obs_enc, (cudaStream_t)stream);
}

int tron_observe(const tron_step_args* args, tron_stream_t stream) {
    StepParams p;
    const int rc = fill_params(args, MODE_OBSERVE, p);
    if (rc != TRON_OK) return rc;
    return dispatch(p, MODE_OBSERVE, args->obs_dtype, args->obs_enc, (cudaStream_t)stream);
}

int tron_export_grid(const void* state, int n_envs, int width, int height, int layout, int8_t* tiles, int8_t* heads,
                     uint8_t* alive, uint8_t* done, uint8_t* winner, int32_t* ep_len, tron_stream_t stream) {
    size_t mo = 0;
    const int rc = tron_state_offsets(n_envs, width, height, layout, nullptr, &mo, nullptr);
    if (rc != TRON_OK || !state) return rc != TRON_OK ? rc : TRON_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (layout == TRON_LAYOUT_TRAIL) return launch_trail_export(state, n_envs, width, height, tiles, heads, alive, done, winner, ep_len, s);
    if (tiles && layout == TRON_LAYOUT_BITS10) {
        if (launch_bits10_export(state, (const char*)state + mo, n_envs, tiles, s) != TRON_OK) return TRON_ERR_CUDA;
    } else if (tiles && cudaMemcpyAsync(tiles, state, (size_t)n_envs * tron_cells_per_env(width, height), cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
        return TRON_ERR_CUDA;
    }
    return launch_export_meta((const char*)state + mo, n_envs, heads, alive, done, winner, ep_len, s);
}

int tron_import_grid(void* state, int n_envs, int width, int height, int layout, const int8_t* tiles, const int8_t* heads,
                     const uint8_t* alive, const uint8_t* done, const uint8_t* winner, const int32_t* ep_len, tron_stream_t stream) {
    size_t mo = 0;
    const int rc = tron_state_offsets(n_envs, width, height, layout, nullptr, &mo, nullptr);
    if (rc != TRON_OK || !state) return rc != TRON_OK ? rc : TRON_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    if (layout == TRON_LAYOUT_TRAIL) return launch_trail_import(state, n_envs, width, height, tiles, heads, alive, done, winner, ep_len, s);
    if (tiles && layout == TRON_LAYOUT_BITS10) {
        if (launch_bits10_import(state, n_envs, tiles, s) != TRON_OK) return TRON_ERR_CUDA;
    } else if (tiles && cudaMemcpyAsync(state, tiles, (size_t)n_envs * tron_cells_per_env(width, height), cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
        return TRON_ERR_CUDA;
    }
    return launch_import_meta((char*)state + mo, n_envs, heads, alive, done, winner, ep_len, s);
}

int tron_random_actions(uint8_t* actions, int n_envs, uint64_t seed, uint64_t counter, const uint64_t* counter_dev, uint64_t env_id_base,
                        tron_stream_t stream) {
    if (!actions || n_envs <= 0) return TRON_ERR_INVALID;
    return launch_random_actions(actions, n_envs, seed, counter, counter_dev, env_id_base, (cudaStream_t)stream);
}
int tron_advance_counter(uint64_t* counter_dev, uint64_t delta, tron_stream_t stream) {
    if (!counter_dev) return TRON_ERR_INVALID;
    return launch_advance_counter(counter_dev, delta, (cudaStream_t)stream);
}

int tron_select_actions(const void* q, int q_dtype, int n_rows, float epsilon, uint8_t* actions, uint64_t seed, uint64_t counter,
                        const uint64_t* counter_dev, uint64_t row_id_base, tron_stream_t stream) {
    if (!q || !actions || n_rows <= 0) return TRON_ERR_INVALID;
    if (((uintptr_t)q & (q_dtype == TRON_F32 ? 15u : 7u)) != 0) return TRON_ERR_ALIGN;
    return launch_select_actions(q, q_dtype, n_rows, epsilon, actions, seed, counter, counter_dev, row_id_base, (cudaStream_t)stream);
}

int tron_minimax_actions(const int8_t* tiles, int n_envs, int width, int height, int player, int tie_mode, uint64_t seed, uint64_t counter,
                         const uint64_t* counter_dev, uint64_t env_id_base, uint8_t* actions, int32_t* values, tron_stream_t stream) {
    if (!tiles || !actions || !geometry_ok(n_envs, width, height) || (player != 1 && player != 2) || (tie_mode != 0 && tie_mode != 1)) return TRON_ERR_INVALID;
    if ((width + 2) * (height + 2) > 256) return TRON_ERR_UNSUPPORTED;
    return launch_minimax(tiles, n_envs, width, height, player, tie_mode, seed, counter, counter_dev, env_id_base, actions, values, (cudaStream_t)stream);
}

int tron_pop_up(const void* obs, int obs_dtype, int64_t n_maps, int cells, void* planes, int out_dtype, tron_stream_t stream) {
    if (!obs || !planes || n_maps <= 0 || cells <= 0) return TRON_ERR_INVALID;
    if (obs_dtype != TRON_F32 && obs_dtype != TRON_BF16 && obs_dtype != TRON_I8 && obs_dtype != TRON_I32 && obs_dtype != TRON_I64) return TRON_ERR_INVALID;
    if (out_dtype != TRON_F32 && out_dtype != TRON_BF16 && out_dtype != TRON_I8) return TRON_ERR_INVALID;
    return launch_pop_up(obs, obs_dtype, n_maps, cells, planes, out_dtype, (cudaStream_t)stream);
}

static int ring_ok(const replay_ring* r) {
    if (!r || r->struct_size != sizeof(replay_ring) || r->capacity <= 0 || r->frame_elems <= 0) return 0;
    if (r->frame_dtype != TRON_BF16 && r->frame_dtype != TRON_F32 && r->frame_dtype != TRON_I8) return 0;
    return r->state && r->next_state && r->action && r->reward && r->done;
}

int replay_push(const replay_ring* ring, uint64_t cursor, const void* state, const void* next_state, const uint8_t* action,
                const float* reward, const uint8_t* done, int done_stride, int64_t n, tron_stream_t stream) {
    if (!ring_ok(ring) || !state || !next_state || !action || !reward || !done) return TRON_ERR_INVALID;
    if (n <= 0 || n > ring->capacity || (done_stride != 1 && done_stride != 2)) return TRON_ERR_INVALID;
    return launch_replay_push(ring, cursor, state, next_state, action, reward, done, done_stride, n, (cudaStream_t)stream);
}

int replay_gather(const replay_ring* ring, const int64_t* idx, int64_t k, void* out_state, void* out_next, int out_dtype,
                  int64_t* out_action, float* out_reward, float* out_done, tron_stream_t stream) {
    if (!ring_ok(ring) || !idx || k <= 0 || !out_state || !out_next || !out_action || !out_reward || !out_done) return TRON_ERR_INVALID;
    if (out_dtype != TRON_F32 && out_dtype != TRON_BF16) return TRON_ERR_INVALID;
    if ((((uintptr_t)out_state | (uintptr_t)out_next | (uintptr_t)ring->state | (uintptr_t)ring->next_state) & 15u) != 0) return TRON_ERR_ALIGN;
    return launch_replay_gather(ring, idx, k, out_state, out_next, out_dtype, out_action, out_reward, out_done, (cudaStream_t)stream);
}

int replay_sample_indices(int64_t size, int k, uint64_t seed, uint64_t counter, int64_t* idx, tron_stream_t stream) {
    if (!idx || k <= 0 || k > 4096 || (int64_t)k > size) return TRON_ERR_INVALID;
    return launch_replay_sample(size, k, seed, counter, idx, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------
// host-buffer front end: what a caller holding numpy arrays binds to (Game.step with host in/out).
// Envs are cut into chunks; chunk i's H2D copy, kernel and D2H copies run on stream i % kStreams so
// PCIe transfers in both directions overlap the kernels of neighbouring chunks.
// ---------------------------------------------------------------------------------------------------
struct tron_host_env {
    tron_step_args proto;
    int n_chunks;
    int planes, cells, esize;
    void* d_state;
    uint8_t* d_actions;
    int8_t* d_spawn;
    void* d_obs;
    float* d_reward;
    uint8_t* d_done;
    uint8_t* d_winner;
    uint64_t counter;
    std::vector<cudaStream_t> streams;
};

static void host_env_free(tron_host_env* e) {
    if (!e) return;
    for (cudaStream_t s : e->streams) cudaStreamDestroy(s);
    cudaFree(e->d_state); cudaFree(e->d_actions); cudaFree(e->d_spawn); cudaFree(e->d_obs);
    cudaFree(e->d_reward); cudaFree(e->d_done); cudaFree(e->d_winner);
    delete e;
}

int tron_host_env_create(tron_host_env** out, const tron_step_args* proto, int n_chunks) {
    if (!out || !proto || proto->struct_size != sizeof(tron_step_args)) return TRON_ERR_INVALID;
    if (!geometry_ok(proto->n_envs, proto->width, proto->height)) return TRON_ERR_INVALID;
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > proto->n_envs) n_chunks = proto->n_envs;
    tron_host_env* e = new (std::nothrow) tron_host_env();
    if (!e) return TRON_ERR_INVALID;
    e->proto = *proto;
    e->n_chunks = n_chunks;
    e->planes = planes_of(proto->obs_enc);
    e->cells = tron_cells_per_env(proto->width, proto->height);
    e->esize = tron_elem(proto->obs_dtype);
    e->counter = 0;
    const size_t N = (size_t)proto->n_envs;
    size_t sb = 0;
    if (tron_state_bytes(proto->n_envs, proto->width, proto->height, proto->layout, &sb) != TRON_OK) { delete e; return TRON_ERR_UNSUPPORTED; }
    bool ok = cudaMalloc(&e->d_state, sb) == cudaSuccess && cudaMalloc((void**)&e->d_actions, N * 2) == cudaSuccess &&
              cudaMalloc((void**)&e->d_spawn, N * 4) == cudaSuccess && cudaMalloc((void**)&e->d_reward, N * 8) == cudaSuccess &&
              cudaMalloc((void**)&e->d_done, N) == cudaSuccess && cudaMalloc((void**)&e->d_winner, N) == cudaSuccess;
    if (ok && e->planes) ok = cudaMalloc(&e->d_obs, N * 2 * e->planes * e->cells * e->esize) == cudaSuccess;
    const int ns = n_chunks < 4 ? n_chunks : 4;
    for (int i = 0; ok && i < ns; ++i) {
        cudaStream_t s;
        ok = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
        if (ok) e->streams.push_back(s);
    }
    if (!ok) { cudaGetLastError(); host_env_free(e); return TRON_ERR_CUDA; }
    *out = e;
    return TRON_OK;
}

int tron_host_env_destroy(tron_host_env* env) { host_env_free(env); return TRON_OK; }
void* tron_host_env_state(tron_host_env* env) { return env ? env->d_state : nullptr; }

// one pass over the chunks; mode MODE_RESET (reset + observe) or MODE_STEP
static int host_env_run(tron_host_env* e, int mode, const uint8_t* actions_h, const int8_t* spawn_h, void* obs_h, float* reward_h,
                        uint8_t* done_h, uint8_t* winner_h) {
    const int N = e->proto.n_envs, C = e->cells;
    const size_t frame = (size_t)2 * e->planes * C * e->esize;  // obs bytes per env
    StepParams base;
    tron_step_args a = e->proto;
    a.state = e->d_state; a.obs = e->d_obs; a.actions = e->d_actions; a.action_dtype = TRON_U8;
    a.reward = e->d_reward; a.done = e->d_done; a.winner = e->d_winner; a.ep_len_out = nullptr; a.stats = nullptr;
    a.spawn = spawn_h ? e->d_spawn : nullptr; a.counter = e->counter; a.n_ticks = 1;
    int rc = fill_params(&a, mode == MODE_RESET ? MODE_RESET : MODE_STEP, base);
    if (rc != TRON_OK) return rc;
    if (mode == MODE_RESET) base.spawn = a.spawn;
    const int per = (N + e->n_chunks - 1) / e->n_chunks;
    bool ok = true;
    for (int c = 0; c < e->n_chunks && ok; ++c) {
        const int lo = c * per, n = (lo + per <= N ? per : N - lo);
        if (n <= 0) break;
        cudaStream_t s = e->streams[c % e->streams.size()];
        if (mode == MODE_STEP) ok = ok && cudaMemcpyAsync(e->d_actions + 2 * (size_t)lo, actions_h + 2 * (size_t)lo, 2 * (size_t)n, cudaMemcpyHostToDevice, s) == cudaSuccess;
        if (spawn_h) ok = ok && cudaMemcpyAsync(e->d_spawn + 4 * (size_t)lo, spawn_h + 4 * (size_t)lo, 4 * (size_t)n, cudaMemcpyHostToDevice, s) == cudaSuccess;
        StepParams p = base;
        p.N = n; p.env_base = base.env_base + (unsigned long long)lo;
        p.grid = base.grid + (size_t)lo * grid_stride(base.layout, base.W, base.H); p.meta = base.meta + lo; p.boxes = base.boxes + lo;
        if (p.actions) p.actions = (const uint8_t*)base.actions + 2 * (size_t)lo;
        if (p.spawn) p.spawn = base.spawn + 4 * (size_t)lo;
        if (p.obs) p.obs = (char*)base.obs + (size_t)lo * frame;
        if (p.reward) p.reward = base.reward + 2 * (size_t)lo;
        if (p.done) p.done = base.done + lo;
        if (p.winner) p.winner = base.winner + lo;
        if (mode == MODE_RESET) {
            rc = dispatch(p, MODE_RESET, TRON_I8, TRON_ENC_NONE, s);
            if (rc == TRON_OK && obs_h && e->planes) {
                StepParams q = p;
                rc = dispatch(q, MODE_OBSERVE, e->proto.obs_dtype, e->proto.obs_enc, s);
            }
        } else {
            rc = dispatch(p, MODE_STEP, e->proto.obs_dtype, e->proto.obs_enc, s);
        }
        if (rc != TRON_OK) return rc;
        if (obs_h && e->planes) ok = ok && cudaMemcpyAsync((char*)obs_h + (size_t)lo * frame, (char*)e->d_obs + (size_t)lo * frame, (size_t)n * frame, cudaMemcpyDeviceToHost, s) == cudaSuccess;
        if (mode == MODE_STEP) {
            if (reward_h) ok = ok && cudaMemcpyAsync(reward_h + 2 * (size_t)lo, e->d_reward + 2 * (size_t)lo, 8 * (size_t)n, cudaMemcpyDeviceToHost, s) == cudaSuccess;
            if (done_h) ok = ok && cudaMemcpyAsync(done_h + lo, e->d_done + lo, (size_t)n, cudaMemcpyDeviceToHost, s) == cudaSuccess;
            if (winner_h) ok = ok && cudaMemcpyAsync(winner_h + lo, e->d_winner + lo, (size_t)n, cudaMemcpyDeviceToHost, s) == cudaSuccess;
        }
    }
    for (cudaStream_t s : e->streams) ok = (cudaStreamSynchronize(s) == cudaSuccess) && ok;
    e->counter += 1;
    if (!ok) { cudaGetLastError(); return TRON_ERR_CUDA; }
    return TRON_OK;
}

int tron_host_env_reset(tron_host_env* env, const int8_t* spawn_host, void* obs_host) {
    if (!env) return TRON_ERR_INVALID;
    return host_env_run(env, MODE_RESET, nullptr, spawn_host, obs_host, nullptr, nullptr, nullptr);
}

int tron_host_env_step(tron_host_env* env, const uint8_t* actions_host, const int8_t* spawn_host, void* obs_host, float* reward_host,
                       uint8_t* done_host, uint8_t* winner_host) {
    if (!env || !actions_host) return TRON_ERR_INVALID;
    return host_env_run(env, MODE_STEP, actions_host, spawn_host, obs_host, reward_host, done_host, winner_host);
}

int tron_host_alloc(void** ptr, size_t bytes) {
    if (!ptr || !bytes) return TRON_ERR_INVALID;
    if (cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return TRON_ERR_CUDA; }
    return TRON_OK;
}
int tron_host_free(void* ptr) {
    if (!ptr) return TRON_OK;
    if (cudaFreeHost(ptr) != cudaSuccess) { cudaGetLastError(); return TRON_ERR_CUDA; }
    return TRON_OK;
}

}  // extern "C"
