#!/usr/bin/env python
"""Config #5 (64x64 pure tick on trail lists): uniform policy, action tape and epsilon-greedy streams (long episodes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import run  # noqa: E402

M = 1 << 20
N = int(os.environ.get("TRAIL_N", 2 * M))
W = int(os.environ.get("TRAIL_W", 64))
run("warm-up (ignore)", N, W, "bf16", "none", steps=400, layout="trail", actions="rng")
run("%dx%d pure step trail, tape" % (W, W), N, W, "bf16", "none", steps=200, layout="trail")
run("%dx%d pure step trail, in-kernel policy" % (W, W), N, W, "bf16", "none", steps=200, layout="trail", actions="rng")
for eps in (0.5, 0.1, 0.003):
    run("%dx%d pure step trail, eps-greedy %g" % (W, W, eps), N, W, "bf16", "none", steps=100, warmup=300, layout="trail", actions="rng", policy="free_eps", eps=eps)
