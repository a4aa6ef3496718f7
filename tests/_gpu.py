"""Helpers for the CUDA parity tests: run the same calls on tron_b200.BatchedTron (GPU, through the C ABI)
and on oracle.c_oracle.OracleEnv (CPU), compare bit for bit."""
import numpy as np
import torch

from oracle import c_oracle as oc
from tron_b200 import abi
from tron_b200.batch_env import BatchedTron

TORCH_DT = {abi.BF16: torch.bfloat16, abi.F32: torch.float32, abi.I8: torch.int8}
ENC_NAME = {abi.ENC_NONE: "none", abi.ENC_LUT1: "lut1", abi.ENC_POPUP3: "popup3", abi.ENC_POPUP3_CONST: "popup3_const"}


def to_np(t):
    """torch tensor -> numpy with the oracle's dtype conventions (bf16 as raw uint16)"""
    if t is None:
        return None
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


class GpuEnvNumpy:
    """BatchedTron with the OracleEnv call surface (numpy in / numpy out) so tests can drive both identically."""

    def __init__(self, n_envs, width=10, height=10, obs_dtype=abi.BF16, obs_enc=abi.ENC_LUT1, **kw):
        slide = kw.pop("slide_mode", abi.SLIDE_NONE)
        self.env = BatchedTron(n_envs, width, height, obs_dtype=obs_dtype, obs_enc=obs_enc, slide_mode=slide, **kw)

    def reset(self, spawn=None, mask=None, counter=None):
        return to_np(self.env.reset(spawn=spawn, mask=mask, counter=counter))

    def observe(self):
        return to_np(self.env.observe())

    def step(self, actions=None, spawn=None, slide_tape=None, counter=None):
        r = self.env.step(actions=None if actions is None else torch.as_tensor(actions), spawn=spawn, slide_tape=slide_tape, counter=counter)
        return tuple(to_np(x) for x in r)

    def step_many(self, n_ticks, actions=None, spawn=None, obs_every_tick=True, counter=None):
        r = self.env.step_many(n_ticks, actions=None if actions is None else torch.as_tensor(actions), spawn=spawn,
                               obs_every_tick=obs_every_tick, counter=counter)
        return tuple(to_np(x) for x in r)

    def export(self):
        return {k: to_np(v) for k, v in self.env.export().items()}

    def import_(self, **kw):
        self.env.import_(**kw)

    @property
    def stats(self):
        s = self.env.stats.view(abi.STATS_SLOTS, abi.STATS_FIELDS).sum(0).cpu().numpy().astype(np.uint64)
        return s


def make_pair(n_envs, width=10, height=10, layout="tile8", **kw):
    return GpuEnvNumpy(n_envs, width, height, layout=layout, **kw), oc.OracleEnv(n_envs, width, height, **kw)


def assert_same_step(got, want, what=""):
    names = ("obs", "reward", "done", "winner", "ep_len")
    for n, g, w in zip(names, got, want):
        if w is None:
            assert g is None
            continue
        assert g.shape == w.shape and g.dtype == w.dtype, (what, n, g.shape, w.shape, g.dtype, w.dtype)
        if not np.array_equal(g, w):
            bad = np.argwhere(g != w)
            raise AssertionError("%s: %s differs at %d places, first %s: got %s want %s" % (what, n, len(bad), bad[0], g[tuple(bad[0])], w[tuple(bad[0])]))


def assert_same_state(genv, oenv, what=""):
    a, b = genv.export(), oenv.export()
    for k in b:
        assert np.array_equal(a[k], b[k]), (what, k)
