#!/usr/bin/env python
"""Throughput of the GPU replay ring kernels (replay_push / replay_gather / replay_sample_indices), CUDA events, buffers >> L2."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tron_b200  # noqa: E402
from tron_b200.replay import ReplayRing  # noqa: E402

PEAK = 6452.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    for planes, dt, name in ((3, torch.bfloat16, "pop_up bf16 (DDQN)"), (1, torch.float32, "1-plane f32 (DQN)")):
        F = planes * 144
        es = 2 if dt == torch.bfloat16 else 4
        n = 1 << 20                      # transitions per push (= 2 x 524,288 envs)
        ring = ReplayRing(4 * n, (planes, 12, 12), dt)
        s = torch.randint(-3, 3, (n, F), device="cuda").to(dt); s2 = torch.randint(-3, 3, (n, F), device="cuda").to(dt)
        a = torch.randint(0, 4, (n,), device="cuda", dtype=torch.uint8); r = torch.randn(n, device="cuda"); d = torch.zeros(n // 2, device="cuda", dtype=torch.uint8)
        t = timed(lambda: ring.push(s, s2, a, r, d, done_stride=2), 10)
        B = n * (4 * F * es + 2 * (1 + 4) + 1.5)  # read s,s' + write s,s' + scalars
        print(json.dumps(dict(op="replay_push", frame=name, transitions=n, seconds=t, transitions_per_s=n / t, GBps=B / t / 1e9, frac_of_measured_peak=B / t / 1e9 / PEAK)))
        for k in (64, 4096, 262144):
            idx = torch.randint(0, len(ring), (k,), device="cuda")
            t = timed(lambda: ring.gather(idx, torch.float32), 20)
            Bg = k * (2 * F * es + 2 * F * 4 + 30)
            print(json.dumps(dict(op="replay_gather->f32", frame=name, rows=k, seconds=t, rows_per_s=k / t, GBps=Bg / t / 1e9, frac_of_measured_peak=Bg / t / 1e9 / PEAK)))
        for k in (64, 4096, 262144):
            t = timed(lambda: ring.sample_indices(k), 50)
            print(json.dumps(dict(op="replay_sample_indices", k=k, seconds=t)))
            t = timed(lambda: ring.sample(k), 50)
            Bg = k * (2 * F * es + 2 * F * 4 + 30)
            print(json.dumps(dict(op="replay_sample_gather->f32 (one launch)", frame=name, k=k, seconds=t, GBps=Bg / t / 1e9, frac_of_measured_peak=Bg / t / 1e9 / PEAK)))
        del ring, s, s2
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
