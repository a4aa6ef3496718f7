#!/usr/bin/env python
"""Small-batch latencies: wall time per Python call (launch-bound regime) and device time per launch (CUDA-graph replay of 50 calls)
for tron_step at 4096 envs and for the one-launch replay sampling (transition ring and frame ring) at k = 64 / 4096."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tron_b200  # noqa: E402
from tron_b200.batch_env import BatchedTron  # noqa: E402
from tron_b200.replay import FrameRing, ReplayRing  # noqa: E402


def wall_us(fn, iters=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6


def device_us(fn, calls=50, replays=20):
    """device time per call: `calls` calls captured into one CUDA graph, replayed back to back"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (calls * replays) * 1e3


def main():
    out = []
    for layout, enc in (("bits10", "lut1"), ("bits10", "popup3"), ("tile8", "lut1")):
        env = BatchedTron(4096, 10, 10, obs_dtype=torch.bfloat16, obs_enc=enc, layout=layout)
        env.use_device_counter()
        obs = env.reset()
        rw = torch.empty((4096, 2), device="cuda"); dn = torch.empty(4096, dtype=torch.uint8, device="cuda"); wn = torch.empty(4096, dtype=torch.uint8, device="cuda")
        fn = lambda: env.step(obs=obs, reward=rw, done=dn, winner=wn, want_ep_len=False)
        w, d = wall_us(fn), device_us(fn)
        out.append(dict(op="tron_step 4096 envs", layout=layout, enc=enc, wall_us_per_call=w, device_us_per_launch=d, env_steps_per_s_wall=4096 / (w * 1e-6),
                        env_steps_per_s_device=4096 / (d * 1e-6)))
        env2 = BatchedTron(4096, 10, 10, obs_dtype=torch.bfloat16, obs_enc=enc, layout=layout)  # host-side counter: one launch per tick
        env2.reset()
        bound = env2.bind_step(obs=obs, reward=rw, done=dn, winner=wn)
        w = wall_us(bound)
        out.append(dict(op="bind_step 4096 envs (pre-bound arguments)", layout=layout, enc=enc, wall_us_per_call=w, env_steps_per_s_wall=4096 / (w * 1e-6)))
    for planes, dt in ((3, torch.bfloat16), (1, torch.float32)):
        ring = ReplayRing(1 << 20, (planes, 12, 12), dt)
        ring.cursor = 1 << 20
        ring.state.random_(0, 2) if dt != torch.bfloat16 else ring.state.copy_(torch.randint(0, 2, ring.state.shape, device="cuda").to(dt))
        env = BatchedTron(65536, 10, 10, obs_dtype=dt, obs_enc="popup3" if planes == 3 else "lut1", layout="auto")
        fr = FrameRing(env, 6)
        fr.begin()
        for _ in range(5):
            fr.step(actions=env.random_actions())
        for k in (64, 4096):
            for name, obj in (("ReplayRing.sample", ring), ("FrameRing.sample", fr)):
                fn = lambda: obj.sample(k)
                bufs = obj.sample(k)
                fn2 = lambda: obj.sample(k, out=bufs)
                out.append(dict(op=name, k=k, planes=planes, dtype=str(dt), wall_us_per_call=wall_us(fn), wall_us_per_call_reusing_outputs=wall_us(fn2),
                                device_us_per_launch=device_us(fn)))
        del ring, fr, env
        torch.cuda.empty_cache()
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
