"""ctypes front end of oracle/_ref/libtron_oracle.so (the plain-C restatement).  TEST INFRASTRUCTURE.

Works on numpy arrays (host memory).  Shares the argument structs with the product ABI
(tron_b200.abi) so tests can build one argument set for both sides.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import tron_b200  # noqa: E402
from tron_b200 import abi  # noqa: E402

_SO = os.path.join(_HERE, "_ref", "libtron_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "tron_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_state_bytes.restype = C.c_size_t
        L.oracle_bench_random.restype = C.c_double
        L.oracle_bench_random.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64,
                                          C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


NP_OF = {abi.U8: np.uint8, abi.I8: np.int8, abi.I32: np.int32, abi.I64: np.int64, abi.F32: np.float32,
         abi.BF16: np.uint16}


def dtype_code(arr):
    return {np.dtype(np.uint8): abi.U8, np.dtype(np.int32): abi.I32, np.dtype(np.int64): abi.I64}[arr.dtype]


class OracleEnv:
    """N independent games on the CPU, same call surface as tron_b200.BatchedTron (numpy in/out)."""

    def __init__(self, n_envs, width=10, height=10, obs_dtype=abi.BF16, obs_enc=abi.ENC_LUT1, lut=None,
                 const_plane=0.0, reward="ddqn", auto_reset=True, seed=0, env_id_base=0,
                 slide_mode=abi.SLIDE_NONE, slide_rate=0.0, spawn_mode=0, policy=0, policy_epsilon=0.0):
        self.N, self.W, self.H = n_envs, width, height
        self.C = abi.cells_per_env(width, height)
        self.P = abi.enc_planes(obs_enc)
        self.obs_dtype, self.obs_enc = obs_dtype, obs_enc
        self.lut = tuple(lut) if lut is not None else (0,) * 6
        self.const_plane = const_plane
        self.reward_table = abi.Reward(*(abi.REWARD_POLICIES[reward] if isinstance(reward, str) else reward))
        self.auto_reset, self.seed, self.env_id_base = int(auto_reset), seed, env_id_base
        self.slide_mode, self.slide_rate = slide_mode, slide_rate
        self.spawn_mode = spawn_mode
        self.policy, self.policy_epsilon = policy, policy_epsilon
        self.state = np.zeros(lib().oracle_state_bytes(n_envs, width, height), np.uint8)
        self.slide_params = np.zeros((n_envs, 4), np.int8)
        self.stats = np.zeros(abi.STATS_FIELDS, np.uint64)
        self.counter = 0

    # -- helpers
    def _args(self, **kw):
        a = abi.new_step_args(n_envs=self.N, width=self.W, height=self.H, state=_p(self.state),
                              obs_dtype=self.obs_dtype, obs_enc=self.obs_enc, lut=self.lut,
                              const_plane=self.const_plane, reward_table=self.reward_table,
                              auto_reset=self.auto_reset, seed=self.seed, env_id_base=self.env_id_base,
                              slide_mode=self.slide_mode, slide_rate=self.slide_rate, spawn_mode=self.spawn_mode, policy=self.policy, policy_epsilon=self.policy_epsilon,
                              slide_params=_p(self.slide_params), stats=_p(self.stats))
        for k, v in kw.items():
            setattr(a, k, v)
        return a

    def new_obs(self, ticks=None):
        shape = (self.N, 2, self.P, self.W + 2, self.H + 2)
        if ticks is not None:
            shape = (ticks,) + shape
        return np.zeros(shape, NP_OF[self.obs_dtype])

    def reset(self, spawn=None, mask=None, counter=None):
        if counter is None:
            counter = self.counter
            self.counter += 1
        sp = None if spawn is None else np.ascontiguousarray(spawn, np.int8)
        mk = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        a = self._args(spawn=_p(sp), counter=counter, obs_enc=abi.ENC_NONE)
        rc = lib().oracle_reset_ex(C.byref(a), _p(mk))
        assert rc == 0, rc
        return self.observe() if self.P else None

    def extra(self):
        """[N,2,2] f32 {degree, weight_p} of the current games (Game.get_multy)"""
        x = np.zeros((self.N, 2, 2), np.float32)
        x[:, 0, 0] = x[:, 1, 0] = self.slide_params[:, 0]
        x[:, 0, 1] = self.slide_params[:, 1]
        x[:, 1, 1] = self.slide_params[:, 2]
        return x

    def observe(self):
        obs = self.new_obs()
        a = self._args(obs=_p(obs))
        rc = lib().oracle_observe(C.byref(a))
        assert rc == 0, rc
        return obs

    def step(self, actions=None, spawn=None, slide_tape=None, counter=None, obs_terminal=None):
        """-> obs, reward[N,2], done[N], winner[N], ep_len[N].  obs_terminal: array like obs; rows of games that finished
        and were auto-reset receive the finished game's last frame."""
        if counter is None:
            counter = self.counter
            self.counter += 1
        obs = self.new_obs() if self.P else None
        reward = np.zeros((self.N, 2), np.float32)
        done = np.zeros(self.N, np.uint8)
        winner = np.zeros(self.N, np.uint8)
        eplen = np.zeros(self.N, np.int32)
        act = None if actions is None else np.ascontiguousarray(actions)
        sp = None if spawn is None else np.ascontiguousarray(spawn, np.int8)
        sl = None if slide_tape is None else np.ascontiguousarray(slide_tape, np.uint8)
        a = self._args(actions=_p(act), action_dtype=0 if act is None else dtype_code(act), obs=_p(obs),
                       reward=_p(reward), done=_p(done), winner=_p(winner), ep_len_out=_p(eplen), spawn=_p(sp),
                       slide_tape=_p(sl), counter=counter, obs_terminal=_p(obs_terminal))
        rc = lib().oracle_step(C.byref(a))
        assert rc == 0, rc
        return obs, reward, done, winner, eplen

    def step_many(self, n_ticks, actions=None, spawn=None, obs_every_tick=True, counter=None):
        if counter is None:
            counter = self.counter
            self.counter += n_ticks
        T = n_ticks
        obs = (self.new_obs(T) if obs_every_tick else self.new_obs()) if self.P else None
        reward = np.zeros((T, self.N, 2), np.float32)
        done = np.zeros((T, self.N), np.uint8)
        winner = np.zeros((T, self.N), np.uint8)
        eplen = np.zeros((T, self.N), np.int32)
        act = None if actions is None else np.ascontiguousarray(actions)
        sp = None if spawn is None else np.ascontiguousarray(spawn, np.int8)
        a = self._args(actions=_p(act), action_dtype=0 if act is None else dtype_code(act), obs=_p(obs),
                       reward=_p(reward), done=_p(done), winner=_p(winner), ep_len_out=_p(eplen), spawn=_p(sp),
                       counter=counter, n_ticks=T, obs_every_tick=int(obs_every_tick))
        rc = lib().oracle_step_many(C.byref(a))
        assert rc == 0, rc
        return obs, reward, done, winner, eplen

    def export(self):
        tiles = np.zeros((self.N, self.W + 2, self.H + 2), np.int8)
        heads = np.zeros((self.N, 4), np.int8)
        alive = np.zeros((self.N, 2), np.uint8)
        done = np.zeros(self.N, np.uint8)
        winner = np.zeros(self.N, np.uint8)
        eplen = np.zeros(self.N, np.int32)
        lib().oracle_export_grid(_p(self.state), self.N, self.W, self.H, _p(tiles), _p(heads), _p(alive), _p(done),
                                 _p(winner), _p(eplen))
        return dict(tiles=tiles, heads=heads, alive=alive, done=done, winner=winner, ep_len=eplen)

    def import_(self, tiles=None, heads=None, alive=None, done=None, winner=None, ep_len=None):
        c = lambda x, dt: None if x is None else np.ascontiguousarray(x, dt)
        lib().oracle_import_grid(_p(self.state), self.N, self.W, self.H, _p(c(tiles, np.int8)), _p(c(heads, np.int8)),
                                 _p(c(alive, np.uint8)), _p(c(done, np.uint8)), _p(c(winner, np.uint8)),
                                 _p(c(ep_len, np.int32)))


def bf16_to_f32(a):
    return (a.astype(np.uint32) << 16).view(np.float32)


def obs_to_float(obs, dtype):
    if dtype == abi.BF16:
        return bf16_to_f32(obs)
    return obs.astype(np.float32)


def philox(seed, counter, stream, tag, sub=0):
    out = (C.c_uint32 * 4)()
    lib().oracle_philox(C.c_uint64(seed), C.c_uint64(counter), C.c_uint64(stream), C.c_uint32(tag), C.c_uint32(sub), out)
    return list(out)


def random_actions(n, seed, counter, base=0):
    a = np.zeros((n, 2), np.uint8)
    lib().oracle_random_actions(_p(a), n, C.c_uint64(seed), C.c_uint64(counter), C.c_uint64(base))
    return a


def select_actions(q, eps, seed, counter, base=0):
    q = np.ascontiguousarray(q, np.float32)
    a = np.zeros(q.shape[0], np.uint8)
    lib().oracle_select_actions(_p(q), q.shape[0], C.c_float(eps), _p(a), C.c_uint64(seed), C.c_uint64(counter),
                                C.c_uint64(base))
    return a


def sample_indices(size, k, seed, counter):
    idx = np.zeros(k, np.int64)
    rc = lib().oracle_replay_sample_indices(C.c_int64(size), C.c_int64(k), C.c_uint64(seed), C.c_uint64(counter), _p(idx))
    assert rc == 0, rc
    return idx


class OracleRing:
    def __init__(self, capacity, frame_elems, frame_dtype=abi.BF16):
        self.capacity, self.F, self.dt = capacity, frame_elems, frame_dtype
        self.state = np.zeros((capacity, frame_elems), NP_OF[frame_dtype])
        self.next_state = np.zeros_like(self.state)
        self.action = np.zeros(capacity, np.uint8)
        self.reward = np.zeros(capacity, np.float32)
        self.done = np.zeros(capacity, np.uint8)
        self.cursor = 0
        self.ring = abi.ReplayRing(struct_size=C.sizeof(abi.ReplayRing), frame_elems=frame_elems, frame_dtype=frame_dtype,
                                   capacity=capacity, state=_p(self.state).value, next_state=_p(self.next_state).value,
                                   action=_p(self.action).value, reward=_p(self.reward).value, done=_p(self.done).value)

    def __len__(self):
        return min(self.cursor, self.capacity)

    def push(self, s, s2, action, reward, done, done_stride=1):
        n = action.shape[0]
        s = np.ascontiguousarray(s); s2 = np.ascontiguousarray(s2)
        action = np.ascontiguousarray(action, np.uint8); reward = np.ascontiguousarray(reward, np.float32)
        done = np.ascontiguousarray(done, np.uint8)
        lib().oracle_replay_push(C.byref(self.ring), C.c_uint64(self.cursor), _p(s), _p(s2), _p(action), _p(reward),
                                 _p(done), done_stride, C.c_int64(n))
        self.cursor += n

    def gather(self, idx, out_dtype=abi.F32):
        idx = np.ascontiguousarray(idx, np.int64)
        k = idx.shape[0]
        s = np.zeros((k, self.F), NP_OF[out_dtype]); s2 = np.zeros_like(s)
        a = np.zeros(k, np.int64); r = np.zeros(k, np.float32); d = np.zeros(k, np.float32)
        lib().oracle_replay_gather(C.byref(self.ring), _p(idx), C.c_int64(k), _p(s), _p(s2), out_dtype, _p(a), _p(r), _p(d))
        return s, a, r, s2, d


def bench_random(n_envs, W, H, ticks, obs_dtype=abi.BF16, obs_enc=abi.ENC_LUT1, seed=0):
    sec = C.c_double()
    steps = lib().oracle_bench_random(n_envs, W, H, ticks, obs_dtype, obs_enc, C.c_uint64(seed), C.byref(sec))
    return steps, sec.value


def fair_bounds(W, H, px, py):
    b = (C.c_int * 8)()
    lib().oracle_fair_bounds(W, H, px, py, b)
    return list(b)


def minimax_actions(env, player, tie_mode=0, counter=0, want_values=False, want_ties=False):
    """MinimaxPlayer(2, voronoi) decisions for `player` in every game of an OracleEnv (tron/minimax.py)."""
    act = np.zeros(env.N, np.uint8)
    vals = np.zeros((env.N, 4), np.int32)
    ties = np.zeros((env.N, 4), np.int32)
    rc = lib().oracle_minimax_actions(_p(env.state), env.N, env.W, env.H, player, tie_mode, C.c_uint64(env.seed), C.c_uint64(counter),
                                      C.c_uint64(env.env_id_base), _p(act), _p(vals), _p(ties))
    assert rc == 0, rc
    if want_values or want_ties:
        return (act,) + ((vals,) if want_values else ()) + ((ties,) if want_ties else ())
    return act


def frames_sample_gather(frames, terminal, action, reward, done, frame_dtype, first_tick, n_ticks, k, seed, counter, out_dtype=abi.F32):
    """oracle of replay_frames_sample_gather: frames/terminal [S, rows, F] (terminal may be None), action u8 [S, rows],
    reward f32 [S, rows], done u8 [S, rows/2] -> (s, a, r, s2, d, idx)"""
    S, rows, F = frames.shape
    frames = np.ascontiguousarray(frames); action = np.ascontiguousarray(action, np.uint8)
    reward = np.ascontiguousarray(reward, np.float32); done = np.ascontiguousarray(done, np.uint8)
    terminal = None if terminal is None else np.ascontiguousarray(terminal)
    fr = abi.ReplayFrames(struct_size=C.sizeof(abi.ReplayFrames), frame_elems=F, frame_dtype=frame_dtype, n_slots=S, rows=rows,
                          frames=_p(frames).value, terminal=None if terminal is None else _p(terminal).value,
                          action=_p(action).value, reward=_p(reward).value, done=_p(done).value)
    s = np.zeros((k, F), NP_OF[out_dtype]); s2 = np.zeros_like(s)
    a = np.zeros(k, np.int64); r = np.zeros(k, np.float32); d = np.zeros(k, np.float32); idx = np.zeros(k, np.int64)
    rc = lib().oracle_replay_frames_sample_gather(C.byref(fr), C.c_int64(first_tick), C.c_int64(n_ticks), C.c_int64(k), C.c_uint64(seed),
                                                  C.c_uint64(counter), _p(s), _p(s2), out_dtype, _p(a), _p(r), _p(d), _p(idx))
    assert rc == 0, rc
    return s, a, r, s2, d, idx
