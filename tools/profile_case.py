#!/usr/bin/env python
"""Run one sweep case for a few steps (to be wrapped by ncu).
usage: profile_case.py <envs> <grid> <dtype> <enc> [steps] [layout] [--tile-bytes B] [--slide MODE] [--actions tape|rng] [--variant V] [--policy P --eps E]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sweep import run  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("envs", type=int); ap.add_argument("grid", type=int); ap.add_argument("dtype"); ap.add_argument("enc")
    ap.add_argument("steps", type=int, nargs="?", default=4); ap.add_argument("layout", nargs="?", default="tile8")
    ap.add_argument("--tile-bytes", type=int, default=0); ap.add_argument("--slide", default=None); ap.add_argument("--actions", default="tape")
    ap.add_argument("--variant", type=int, default=0); ap.add_argument("--policy", default="uniform"); ap.add_argument("--eps", type=float, default=0.0)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    if a.tile_bytes:
        from tron_b200 import _lib, abi
        _lib.check(_lib.load().tron_set_option(abi.OPT_TILE_BYTES, a.tile_bytes))
    run("profile", a.envs, a.grid, a.dtype, a.enc, steps=a.steps, warmup=a.warmup, layout=a.layout, slide_mode=a.slide, actions=a.actions, variant=a.variant,
        policy=a.policy, eps=a.eps)
