#!/usr/bin/env python
"""A/B helper: run sweep cases given on the command line as N,W,dtype,enc,layout,actions[,steps] (repeat each 3x)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import run  # noqa: E402

run("warm-up", 1 << 21, 10, "bf16", "lut1", steps=100, layout="bits10")
for spec in sys.argv[1:]:
    f = spec.split(",")
    N, W, dt, enc, layout, actions = int(f[0]), int(f[1]), f[2], f[3], f[4], f[5]
    steps = int(f[6]) if len(f) > 6 else 20
    for rep in range(3):
        run(spec, N, W, dt, enc, steps=steps, layout=layout, actions=actions)
