"""One-env GPU contexts shared by the drop-in classes (kept per grid size; no CPU compute path)."""
import numpy as np
import torch

import tron_b200
from tron_b200 import _lib
from tron_b200 import abi
from tron_b200.batch_env import BatchedTron

_ctx = {}


def env_for(width, height, slide_mode=None):
    """A cached 1-env BatchedTron with int8 observations and no auto-reset (Game semantics)."""
    key = (width, height, slide_mode)
    if key not in _ctx:
        # int8 layout: a Map may hold ANY arrangement of tiles (Map.__setitem__), which only the Tile.value grid can represent
        _ctx[key] = BatchedTron(1, width, height, obs_dtype=torch.int8, obs_enc="lut1", auto_reset=False, slide_mode=slide_mode,
                                collect_stats=False, layout="tile8")
    return _ctx[key]


class OneGame:
    """A 1-env BatchedTron plus one packed device buffer, so a tick costs one kernel, one export and ONE device->host copy.

    pack layout (bytes): [0:288) both observations int8 | [288:432) tiles | [432:436) heads | [436:438) alive | [438] done | [439] winner
    """
    _pool = {}

    def __init__(self, width, height, slide_mode):
        import ctypes as C
        self.key = (width, height, slide_mode)
        self.env = BatchedTron(1, width, height, obs_dtype=torch.int8, obs_enc="lut1", auto_reset=False, slide_mode=slide_mode, collect_stats=False,
                               layout="tile8")
        c = self.env.C
        self.c = c
        o_obs, o_tiles = 0, (2 * c + 15) & ~15
        o_heads = (o_tiles + c + 15) & ~15
        self.off = (o_obs, o_tiles, o_heads, o_heads + 4, o_heads + 6, o_heads + 7)
        self.pack = torch.zeros(o_heads + 16, dtype=torch.uint8, device=self.env.device)
        self.host = torch.zeros(o_heads + 16, dtype=torch.uint8).pin_memory()
        self.obs = self.pack[:2 * c].view(torch.int8).view(1, 2, 1, width + 2, height + 2)
        self.actions = torch.zeros((1, 2), dtype=torch.uint8).pin_memory()
        self.actions_dev = torch.zeros((1, 2), dtype=torch.uint8, device=self.env.device)
        self.tape = torch.zeros((1, 2), dtype=torch.uint8).pin_memory()
        self.tape_dev = torch.zeros((1, 2), dtype=torch.uint8, device=self.env.device)
        self.reward = torch.zeros((1, 2), dtype=torch.float32, device=self.env.device)
        self.scratch = torch.zeros(8, dtype=torch.uint8, device=self.env.device)

    @classmethod
    def acquire(cls, width, height, slide_mode=None):
        free = cls._pool.setdefault((width, height, slide_mode), [])
        return free.pop() if free else cls(width, height, slide_mode)

    def release(self):
        self._pool.setdefault(self.key, []).append(self)

    def _fetch(self):
        """export the state next to the observations and bring everything to the host in one copy"""
        e, base, (o_obs, o_tiles, o_heads, o_alive, o_done, o_win) = self.env, self.pack.data_ptr(), self.off
        _lib.check(e.lib.tron_export_grid(e.state.data_ptr(), 1, e.W, e.H, e.layout, base + o_tiles, base + o_heads, base + o_alive,
                                          base + o_done, base + o_win, None, e._stream()), "tron_export_grid")
        self.host.copy_(self.pack, non_blocking=True)
        torch.cuda.current_stream(e.device).synchronize()
        h = self.host.numpy()
        c = self.c
        obs = h[:2 * c].view(np.int8).astype(np.int64).reshape(2, e.W + 2, e.H + 2)
        return dict(obs1=obs[0], obs2=obs[1], tiles=h[o_tiles:o_tiles + c].view(np.int8).reshape(e.W + 2, e.H + 2).copy(),
                    heads=h[o_heads:o_heads + 4].view(np.int8).copy(), alive=h[o_alive:o_alive + 2].copy(), done=int(h[o_done]), winner=int(h[o_win]))

    def reset(self, spawn):
        self.env.reset(spawn=spawn, obs=self.obs)
        return self._fetch()

    def step(self, actions, slide_tape=None):
        self.actions[0, 0], self.actions[0, 1] = int(actions[0]), int(actions[1])
        self.actions_dev.copy_(self.actions, non_blocking=True)
        tp = None
        if slide_tape is not None:
            self.tape[0, 0], self.tape[0, 1] = int(slide_tape[0]), int(slide_tape[1])
            self.tape_dev.copy_(self.tape, non_blocking=True)
            tp = self.tape_dev
        self.env.step(self.actions_dev, slide_tape=tp, obs=self.obs, reward=self.reward, done=self.scratch[0:1], winner=self.scratch[1:2],
                      want_ep_len=False)
        return self._fetch()


def new_env(width, height, slide_mode=None):
    return OneGame.acquire(width, height, slide_mode)


def observe_codes(codes, width, height):
    """Tile.value grid (W+2,H+2) int8 -> (obs_p1, obs_p2) int64, computed by tron_observe on the GPU."""
    env = env_for(width, height)
    env.import_(tiles=torch.as_tensor(np.ascontiguousarray(codes, np.int8)).view(1, width + 2, height + 2))
    o = env.observe().cpu().numpy().astype(np.int64)
    return o[0, 0, 0], o[0, 1, 0]


def pop_up_gpu(obs):
    """pop_up (reference tron/util.py:11-37) through tron_pop_up: (R,C) any numeric -> (3,R,C) float64"""
    import ctypes as C
    lib = _lib.load()
    _lib.require_cuda()
    a = np.ascontiguousarray(obs)
    t = torch.as_tensor(a.astype(np.int64)).cuda()
    out = torch.empty((3,) + tuple(a.shape), dtype=torch.float32, device="cuda")
    _lib.check(lib.tron_pop_up(t.data_ptr(), abi.I64, 1, int(a.size), out.data_ptr(), abi.F32, torch.cuda.current_stream().cuda_stream), "tron_pop_up")
    return out.cpu().numpy().astype(np.float64)
