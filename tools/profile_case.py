#!/usr/bin/env python
"""Run one sweep case for a few steps (to be wrapped by ncu).  usage: profile_case.py <envs> <grid> <dtype> <enc> [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import run  # noqa: E402

if __name__ == "__main__":
    n, w, dt, enc = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    layout = sys.argv[6] if len(sys.argv) > 6 else "tile8"
    if len(sys.argv) > 7:
        from tron_b200 import _lib, abi
        _lib.check(_lib.load().tron_set_option(abi.OPT_TILE_BYTES, int(sys.argv[7])))
    run("profile", n, w, dt, enc, steps=steps, warmup=2, layout=layout)
