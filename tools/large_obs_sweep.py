#!/usr/bin/env python
"""Fused observations on larger boards: int8 tile kernel vs trail lists (which one should layout="auto" pick?)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import run  # noqa: E402

for W, n in ((16, 1 << 20), (20, 1 << 20), (24, 1 << 19), (32, 1 << 19), (48, 1 << 18), (64, 1 << 17)):
    for enc, dt in (("lut1", "bf16"), ("popup3", "bf16"), ("lut1", "i8")):
        for layout in ("tile8", "trail"):
            run("%dx%d %s %s" % (W, W, dt, enc), n, W, dt, enc, steps=10, layout=layout, actions="rng")
