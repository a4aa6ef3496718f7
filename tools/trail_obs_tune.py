#!/usr/bin/env python
"""the bulk-store trail observation kernel against the element-store kernel it replaces (variant bit 32) and the int8 tile kernel"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import run  # noqa: E402

run("warm-up", 1 << 21, 10, "bf16", "lut1", steps=100, layout="bits10")
for W, n in ((20, 1 << 20), (24, 1 << 19), (32, 1 << 19), (48, 1 << 18), (64, 1 << 17)):
    for enc, dt in (("lut1", "bf16"), ("popup3", "bf16"), ("lut1", "f32")):
        run("%dx%d %s %s trail bulk-store" % (W, W, dt, enc), n, W, dt, enc, steps=30, layout="trail", actions="rng")
        run("%dx%d %s %s trail element stores" % (W, W, dt, enc), n, W, dt, enc, steps=30, layout="trail", actions="rng", variant=32)
        run("%dx%d %s %s tile8" % (W, W, dt, enc), n, W, dt, enc, steps=30, layout="tile8", actions="rng")
for W, n in ((16, 1 << 20), (21, 1 << 20), (32, 1 << 19), (64, 1 << 17)):
    run("%dx%d i8 lut1 trail bulk-store" % (W, W), n, W, "i8", "lut1", steps=30, layout="trail", actions="rng")
    run("%dx%d i8 lut1 trail element stores" % (W, W), n, W, "i8", "lut1", steps=30, layout="trail", actions="rng", variant=32)
    run("%dx%d i8 lut1 tile8" % (W, W), n, W, "i8", "lut1", steps=30, layout="tile8", actions="rng")
