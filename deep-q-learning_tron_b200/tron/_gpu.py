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
        _ctx[key] = BatchedTron(1, width, height, obs_dtype=torch.int8, obs_enc="lut1", auto_reset=False, slide_mode=slide_mode,
                                collect_stats=False)
    return _ctx[key]


def new_env(width, height, slide_mode=None):
    return BatchedTron(1, width, height, obs_dtype=torch.int8, obs_enc="lut1", auto_reset=False, slide_mode=slide_mode, collect_stats=False)


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
