#!/usr/bin/env python
"""Throughput sweep over the BASELINE configurations on one GPU (CUDA events, inputs >> L2).  Prints one JSON line per case.

  python tools/sweep.py [--quick]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tron_b200  # noqa: E402
from tron_b200.batch_env import BatchedTron  # noqa: E402

PEAK = 6452.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def run(name, N, W, dtype, enc, steps=20, warmup=3, actions="tape", auto_reset=True, layout="tile8"):
    tdt = {"bf16": torch.bfloat16, "f32": torch.float32, "i8": torch.int8}[dtype]
    env = BatchedTron(N, W, W, obs_dtype=tdt, obs_enc=enc, seed=0, layout=layout)
    obs = env.reset()
    tape = [env.random_actions(100 + i) for i in range(4)] if actions == "tape" else [None] * 4
    reward = torch.empty((N, 2), dtype=torch.float32, device="cuda"); done = torch.empty(N, dtype=torch.uint8, device="cuda")
    winner = torch.empty(N, dtype=torch.uint8, device="cuda")
    for i in range(warmup):
        env.step(tape[i & 3], obs=obs, reward=reward, done=done, winner=winner, want_ep_len=False)
    torch.cuda.synchronize()
    s0 = env.stats_dict()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        env.step(tape[i & 3], obs=obs, reward=reward, done=done, winner=winner, want_ep_len=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    s1 = env.stats_dict()
    f = (s1["episodes"] - s0["episodes"]) / max(1, s1["env_steps"] - s0["env_steps"])
    C, P = env.C, env.P
    b_o = {"bf16": 2, "f32": 4, "i8": 1}[dtype]
    if P:
        B = (32 if layout == "bits10" else C) * (1 + f) + 2 * P * C * b_o + 48
        if layout == "trail":
            B = 128 + 2 * P * C * b_o + 48
    else:
        B = 320 + f * C  # SURVEY 8d pure-step sector model
        if layout == "trail":
            B = 64 + 64 + 16  # one 64-byte record head read + written back, actions + reward/done/winner
    rate = N / (ms * 1e-3)
    out = dict(case=name, layout=layout, envs=N, grid=W, obs=dtype, enc=enc, ms_per_step=ms, env_steps_per_s=rate, reset_fraction=f, bytes_per_env_step=B,
               achieved_GBps=rate * B / 1e9, frac_of_measured_peak=rate * B / 1e9 / PEAK)
    print(json.dumps(out), flush=True)
    del env, obs
    torch.cuda.empty_cache()
    return out


def main():
    quick = "--quick" in sys.argv
    M = 1 << 20
    for lay in ("bits10", "tile8"):
        run("10x10 bf16 1-plane (headline)", 4 * M, 10, "bf16", "lut1", layout=lay)
    run("10x10 f32 1-plane", 2 * M, 10, "f32", "lut1", layout="bits10")
    run("10x10 i8 1-plane", 4 * M, 10, "i8", "lut1", layout="bits10")
    run("10x10 bf16 pop_up 3-plane (DDQN path)", 2 * M, 10, "bf16", "popup3", layout="bits10")
    run("10x10 pure step (no obs)", 8 * M, 10, "bf16", "none", layout="bits10")
    run("10x10 bf16 1-plane, in-kernel Philox policy", 4 * M, 10, "bf16", "lut1", actions="rng")
    run("10x10 f32 1-plane", 2 * M, 10, "f32", "lut1")
    run("10x10 i8 1-plane", 4 * M, 10, "i8", "lut1")
    run("10x10 bf16 pop_up 3-plane (DDQN path)", 2 * M, 10, "bf16", "popup3")
    run("10x10 bf16 pop_up + const plane", 2 * M, 10, "bf16", "popup3_const")
    run("10x10 pure step (no obs)", 4 * M, 10, "bf16", "none")
    run("64x64 bf16 1-plane", 128 * 1024, 64, "bf16", "lut1", steps=10)
    run("64x64 bf16 1-plane, trail-list state", 128 * 1024, 64, "bf16", "lut1", steps=10, layout="trail")
    run("32x32 bf16 1-plane, trail-list state", 512 * 1024, 32, "bf16", "lut1", steps=10, layout="trail")
    run("64x64 pure step (config #5, 2M envs/GPU)", 2 * M, 64, "bf16", "none", steps=10)
    run("64x64 pure step, trail-list state", 2 * M, 64, "bf16", "none", steps=20, layout="trail")
    run("64x64 pure step, trail-list state, in-kernel policy", 2 * M, 64, "bf16", "none", steps=20, layout="trail", actions="rng")
    if not quick:
        run("10x10 bf16, 65,536 envs (config #3 size)", 65536, 10, "bf16", "popup3", steps=200)
        run("10x10 bf16, 4,096 envs (config #2 size, launch-bound)", 4096, 10, "bf16", "lut1", steps=500)
        run("32x32 bf16 1-plane", 512 * 1024, 32, "bf16", "lut1", steps=10)


if __name__ == "__main__":
    main()
