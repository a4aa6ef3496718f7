#!/usr/bin/env python
"""Randomised differential test: random (layout, size, encoding, dtype, slide mode, policy, spawn mode, tapes / RNG, step /
step_many / masked reset) configurations, CUDA (through the C ABI) vs the CPU oracle, bit for bit.  usage: fuzz_gpu.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import c_oracle as oc  # noqa: E402
from tron_b200 import abi  # noqa: E402
from _gpu import assert_same_state, assert_same_step, make_pair  # noqa: E402


def one_case(rng, case):
    layout = rng.choice(["tile8", "bits10", "trail"])
    W = 10 if layout == "bits10" else int(rng.choice([2, 3, 5, 8, 10, 12, 15, 21, 32, 47, 64]))
    H = W if (layout == "bits10" or rng.random() < 0.8) else int(rng.integers(2, 20))
    N = int(rng.choice([1, 7, 128, 129, 1000, 3000])) if W <= 32 else int(rng.choice([3, 40, 300]))
    enc = int(rng.choice([abi.ENC_NONE, abi.ENC_LUT1, abi.ENC_POPUP3, abi.ENC_POPUP3_CONST]))
    dt = int(rng.choice([abi.BF16, abi.F32, abi.I8]))
    slide = abi.SLIDE_NONE if layout == "bits10" else int(rng.choice([abi.SLIDE_NONE, abi.SLIDE_NONE, abi.SLIDE_ICE, abi.SLIDE_TEMPER, abi.SLIDE_TAPE]))
    policy = int(rng.choice([abi.POLICY_UNIFORM, abi.POLICY_FREE_EPS]))
    kw = dict(layout=layout, obs_dtype=dt, obs_enc=enc, const_plane=float(rng.integers(-3, 9)), seed=int(rng.integers(0, 1 << 30)),
              env_id_base=int(rng.integers(0, 1 << 20)), slide_mode=slide, slide_rate=float(rng.choice([0.0, 0.15, 0.7])),
              spawn_mode=int(rng.integers(0, 2)), policy=policy, policy_epsilon=float(rng.choice([0.0, 0.05, 0.5])),
              auto_reset=bool(rng.random() < 0.8), reward=str(rng.choice(["ddqn", "survivor", "acktr2"])))
    desc = "case %d: %s %dx%d N=%d enc=%d dt=%d slide=%d policy=%d auto=%s" % (case, layout, W, H, N, enc, dt, slide, policy, kw["auto_reset"])
    g, o = make_pair(N, W, H, **kw)
    if slide == abi.SLIDE_TEMPER:
        prm = np.stack([rng.integers(-30, 31, N), rng.integers(40, 102, N), rng.integers(40, 102, N), np.zeros(N, np.int64)], 1).astype(np.int8)
        g.env.slide_params.copy_(torch.as_tensor(prm)); o.slide_params[...] = prm
    a, b = g.reset(), o.reset()
    assert (a is None and b is None) or np.array_equal(a, b), desc
    use_tape = rng.random() < 0.5
    for t in range(int(rng.integers(5, 40))):
        act = rng.integers(0, 4, size=(N, 2)).astype(rng.choice([np.uint8, np.int32, np.int64])) if use_tape else None
        st = rng.integers(0, 2, size=(N, 2)).astype(np.uint8) if slide == abi.SLIDE_TAPE else None
        r = rng.random()
        if r < 0.15 and slide != abi.SLIDE_TAPE:
            T = int(rng.integers(2, 6))
            acts = rng.integers(0, 4, size=(T, N, 2)).astype(np.uint8) if use_tape else None
            every = bool(rng.integers(0, 2))
            assert_same_step(g.step_many(T, actions=acts, obs_every_tick=every), o.step_many(T, actions=acts, obs_every_tick=every), desc + " step_many")
        elif r < 0.2:
            mask = rng.integers(0, 2, size=N).astype(np.uint8)
            x, y = g.reset(mask=mask), o.reset(mask=mask)
            assert (x is None and y is None) or np.array_equal(x, y), desc + " masked reset"
        else:
            assert_same_step(g.step(act, slide_tape=st), o.step(act, slide_tape=st), desc + " tick %d" % t)
    assert_same_state(g, o, desc)
    assert np.array_equal(g.stats, o.stats), desc
    return desc


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    t0, n = time.time(), 0
    while time.time() - t0 < budget:
        last = one_case(rng, n)
        n += 1
        if n % 25 == 0:
            print(last, flush=True)
    print("fuzz ok: %d random configurations bit-exact vs oracle in %.0f s (seed %d)" % (n, time.time() - t0, seed))
