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


def state_bytes(layout, C):
    return {"bits10": 32, "bits": 48}.get(layout, C)


def run(name, N, W, dtype, enc, steps=20, warmup=3, actions="tape", auto_reset=True, layout="tile8", slide_mode=None, variant=0, policy="uniform", eps=0.0):
    tdt = {"bf16": torch.bfloat16, "f32": torch.float32, "i8": torch.int8}[dtype]
    tron_b200.lib.check(tron_b200.lib.load().tron_set_option(tron_b200.abi.OPT_ENCODE_VARIANT, variant), "tron_set_option")
    env = BatchedTron(N, W, W, obs_dtype=tdt, obs_enc=enc, seed=0, layout=layout, slide_mode=slide_mode, policy=policy, policy_epsilon=eps)
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
    M = 28 + (2 if actions == "tape" else 0)  # meta 8 r + 8 w, reward 8, done 1, winner 1 (+ actions 2), as the kernels really move them
    if P:
        B = state_bytes(layout, C) * (1 + f) + 2 * P * C * b_o + M
        if layout in ("bits10", "bits"):
            B = 2 * state_bytes(layout, C) + 2 * P * C * b_o + M  # planes are read and written every tick (a reset writes zeros)
        if layout == "trail":
            B = 64 + 32 + 2 * P * C * b_o + M - 16
    else:
        B = 320 + f * C  # SURVEY 8d pure-step sector model
        if layout in ("bits10", "bits"):
            B = 2 * state_bytes(layout, C) + M
        if layout == "trail":
            B = 64 + 32 + M - 16  # 64 hot bytes read; header + the one list uint4 that changed written; reward/done/winner (+ actions)
    rate = N / (ms * 1e-3)
    out = dict(case=name, layout=layout, envs=N, grid=W, obs=dtype, enc=enc, slide_mode=slide_mode, variant=variant, policy=policy, ms_per_step=ms, env_steps_per_s=rate, reset_fraction=f, bytes_per_env_step=B,
               achieved_GBps=rate * B / 1e9, frac_of_measured_peak=rate * B / 1e9 / PEAK)
    print(json.dumps(out), flush=True)
    del env, obs
    torch.cuda.empty_cache()
    return out


def main():
    quick = "--quick" in sys.argv
    M = 1 << 20
    if "--r2" in sys.argv:  # round-2 focus: training encodings, temper on bit planes, config #5, encode-schedule variants
        for v in (0, 8):
            run("10x10 bf16 1-plane bits10 variant %d" % v, 4 * M, 10, "bf16", "lut1", layout="bits10", variant=v)
            run("10x10 bf16 pop_up3 bits10 variant %d" % v, 2 * M, 10, "bf16", "popup3", layout="bits10", variant=v)
            run("10x10 bf16 pop_up3+const bits10 variant %d" % v, 2 * M, 10, "bf16", "popup3_const", layout="bits10", variant=v)
        run("10x10 bf16 pop_up3 tile8", 2 * M, 10, "bf16", "popup3")
        for v in (0, 8):
            run("10x10 temper bf16 1-plane bits variant %d" % v, 4 * M, 10, "bf16", "lut1", layout="bits", slide_mode="temper", actions="rng", variant=v)
        run("10x10 temper bf16 1-plane tile8", 4 * M, 10, "bf16", "lut1", layout="tile8", slide_mode="temper", actions="rng")
        run("10x10 temper bf16 pop_up3 bits", 2 * M, 10, "bf16", "popup3", layout="bits", slide_mode="temper", actions="rng")
        run("10x10 ice bf16 1-plane bits", 4 * M, 10, "bf16", "lut1", layout="bits", slide_mode="ice", actions="rng")
        run("10x10 f32 1-plane bits10", 2 * M, 10, "f32", "lut1", layout="bits10")
        run("10x10 i8 1-plane bits10", 4 * M, 10, "i8", "lut1", layout="bits10")
        run("10x10 pure step bits10", 8 * M, 10, "bf16", "none", layout="bits10")
        run("8x8 bf16 1-plane bits", 4 * M, 8, "bf16", "lut1", layout="bits", actions="rng")
        run("8x8 bf16 1-plane tile8", 4 * M, 8, "bf16", "lut1", layout="tile8", actions="rng")
        for v in (0,):
            run("64x64 pure step trail, tape, variant %d" % v, 2 * M, 64, "bf16", "none", steps=20, layout="trail", variant=v)
            run("64x64 pure step trail, in-kernel policy, variant %d" % v, 2 * M, 64, "bf16", "none", steps=20, layout="trail", actions="rng", variant=v)
            run("64x64 pure step trail, eps-greedy 0.1 (long episodes), variant %d" % v, 2 * M, 64, "bf16", "none", steps=20, warmup=60, layout="trail", actions="rng",
                policy="free_eps", eps=0.1, variant=v)
        run("64x64 bf16 1-plane trail", 128 * 1024, 64, "bf16", "lut1", steps=10, layout="trail")
        run("64x64 bf16 1-plane tile8", 128 * 1024, 64, "bf16", "lut1", steps=10)
        run("10x10 bf16 pop_up3, 65,536 envs (config #3/#4 size)", 65536, 10, "bf16", "popup3", steps=200, layout="bits10")
        run("10x10 bf16 pop_up3, 131,072 envs (config #4 size)", 131072, 10, "bf16", "popup3", steps=200, layout="bits10")
        return
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
