#!/usr/bin/env python
"""small boards: bulk-store trail kernel vs the int8 tile kernel vs bit planes"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import run  # noqa: E402

run("warm-up", 1 << 21, 10, "bf16", "lut1", steps=100, layout="bits10")
M = 1 << 20
for W, n in ((6, 4 * M), (8, 4 * M), (10, 4 * M), (12, 2 * M), (16, 2 * M)):
    for enc, dt in (("lut1", "bf16"), ("popup3", "bf16"), ("lut1", "i8")):
        for layout in ("trail", "tile8") + (("bits",) if W * W <= 128 else ()):
            run("%dx%d %s %s" % (W, W, dt, enc), n // (3 if enc == "popup3" else 1), W, dt, enc, steps=30, layout=layout, actions="rng")
