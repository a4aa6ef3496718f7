"""Helpers shared by the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def ragged_to_tapes(length, actions, extra=None):
    """Flat per-tick records (one leading 'initial' row per game) -> padded [T,G,2] tapes + row index [T+1,G]."""
    G = len(length)
    T = int(length.max())
    starts = np.concatenate([[0], np.cumsum(length + 1)[:-1]])
    tape = np.zeros((T, G, 2), np.uint8)
    ext = None if extra is None else np.zeros((T, G, 2), np.uint8)
    row = -np.ones((T + 1, G), np.int64)
    for g in range(G):
        row[0, g] = starts[g]
        for t in range(int(length[g])):
            tape[t, g] = actions[starts[g] + 1 + t]
            if ext is not None:
                ext[t, g] = extra[starts[g] + 1 + t]
            row[t + 1, g] = starts[g] + 1 + t
    return tape, ext, row


def digest_with(env_factory, seed, ngames, W=10):
    """SURVEY 8c digest, driven through any env exposing reset(spawn)->obs and step(actions)->obs,rew,done,winner,..

    env_factory(W) must build a 1-env, int8-observation, non-auto-reset environment returning numpy arrays."""
    rng = np.random.default_rng(seed)
    h = hashlib.sha256()
    steps = 0
    wins = [0, 0, 0]
    env = env_factory(W)
    for _ in range(ngames):
        while True:
            x1, y1, x2, y2 = (int(v) for v in rng.integers(0, [W, W, W, W]))
            if (x1, y1) != (x2, y2):
                break
        obs = env.reset(spawn=np.array([[x1, y1, x2, y2]], np.int8))
        h.update(np.asarray(obs)[0, 0, 0].astype(np.int8).tobytes())
        done = False
        while not done:
            a1, a2 = (int(v) for v in rng.integers(0, 4, size=2))
            obs, rew, dn, wn, _ = env.step(np.array([[a1, a2]], np.uint8))
            obs = np.asarray(obs)
            steps += 1
            h.update(obs[0, 0, 0].astype(np.int8).tobytes())
            h.update(obs[0, 1, 0].astype(np.int8).tobytes())
            done = bool(dn[0])
            h.update(bytes([int(done), int(wn[0])]))
        wins[int(wn[0])] += 1
    return h.hexdigest(), steps, wins


# ---- learn-step fixture (tests/golden/learn.npz, made by make_learn_golden.py from the reference's own learn steps) ----
def fill_params(net, seed):
    """Deterministic, framework-independent initial weights: every parameter of `net` (in named_parameters order) is drawn from
    numpy's PCG64 as uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the scale of PyTorch's default Conv2d / Linear init).  Used by the
    fixture generator on the reference's nets and by the tests on ours, so the fixture does not have to store 2 MB per net."""
    import torch
    rng = np.random.default_rng(seed)
    with torch.no_grad():
        for name, p in net.named_parameters():
            w = dict(net.named_parameters())[name.rsplit(".", 1)[0] + ".weight"]
            fan_in = int(np.prod(w.shape[1:]))
            b = 1.0 / np.sqrt(fan_in)
            p.copy_(torch.from_numpy(rng.uniform(-b, b, size=tuple(p.shape)).astype(np.float32)))


def summarize(t):
    """compact signature of a tensor: [sum, sum of |x|] in float64 and up to 256 evenly strided elements"""
    a = np.asarray(t, dtype=np.float32).reshape(-1)
    stride = max(1, a.size // 256)
    return np.array([a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()]), a[::stride][:256].copy()
