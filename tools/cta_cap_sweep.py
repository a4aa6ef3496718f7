#!/usr/bin/env python
"""Resident CTAs per SM of the fused kernels (TRON_OPT_BITS_CTAS_PER_SM / TRON_OPT_TILE_CTAS_PER_SM) vs throughput.
usage: cta_cap_sweep.py bits|tile"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import run  # noqa: E402  (puts the repository root on sys.path)
import tron_b200  # noqa: E402

M = 1 << 20
which = sys.argv[1] if len(sys.argv) > 1 else "bits"
if which == "i8":
    for cap in (32, 7, 6, 5, 4):
        tron_b200.lib.check(tron_b200.lib.load().tron_set_option(tron_b200.abi.OPT_BITS_CTAS_PER_SM, cap))
        run("cap %d: 10x10 i8 1-plane bits10" % cap, 4 * M, 10, "i8", "lut1", layout="bits10", steps=40)
        run("cap %d: 10x10 i8 pop_up3 bits10" % cap, 4 * M, 10, "i8", "popup3", layout="bits10", steps=40)
        run("cap %d: 10x10 i8 pop_up3+const bits10" % cap, 4 * M, 10, "i8", "popup3_const", layout="bits10", steps=40)
elif which == "bits":
    for cap in (32, 7, 6, 5, 4, 3, 2):
        tron_b200.lib.check(tron_b200.lib.load().tron_set_option(tron_b200.abi.OPT_BITS_CTAS_PER_SM, cap))
        run("cap %d: 10x10 bf16 1-plane bits10" % cap, 4 * M, 10, "bf16", "lut1", layout="bits10", steps=40)
        run("cap %d: 10x10 bf16 pop_up3 bits10" % cap, 2 * M, 10, "bf16", "popup3", layout="bits10", steps=40)
        run("cap %d: 10x10 bf16 pop_up3+const bits10" % cap, 2 * M, 10, "bf16", "popup3_const", layout="bits10", steps=40)
        run("cap %d: 10x10 f32 1-plane bits10" % cap, 2 * M, 10, "f32", "lut1", layout="bits10", steps=40)
        run("cap %d: 10x10 temper bf16 1-plane bits" % cap, 4 * M, 10, "bf16", "lut1", layout="bits", slide_mode="temper", actions="rng", steps=40)
else:
    for cap in (32, 6, 5, 4, 3, 2):
        tron_b200.lib.check(tron_b200.lib.load().tron_set_option(tron_b200.abi.OPT_TILE_CTAS_PER_SM, cap))
        run("cap %d: 10x10 bf16 1-plane tile8" % cap, 4 * M, 10, "bf16", "lut1", layout="tile8", steps=40)
        run("cap %d: 10x10 bf16 pop_up3 tile8" % cap, 2 * M, 10, "bf16", "popup3", layout="tile8", steps=40)
        run("cap %d: 10x10 temper bf16 1-plane tile8" % cap, 4 * M, 10, "bf16", "lut1", layout="tile8", slide_mode="temper", actions="rng", steps=40)
        run("cap %d: 8x8 bf16 1-plane tile8" % cap, 4 * M, 8, "bf16", "lut1", layout="tile8", actions="rng", steps=40)
        run("cap %d: 32x32 bf16 1-plane tile8" % cap, 512 * 1024, 32, "bf16", "lut1", layout="tile8", steps=10)
        run("cap %d: 64x64 bf16 1-plane tile8" % cap, 128 * 1024, 64, "bf16", "lut1", layout="tile8", steps=10)
