#!/usr/bin/env python
"""config #5 with several ticks per launch (tron_step_many on the trail layout)"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tron_b200  # noqa
from tron_b200.batch_env import BatchedTron

N = 1 << 21
for T in (1, 4, 16):
    env = BatchedTron(N, 64, 64, obs_enc="none", seed=0, layout="trail")
    env.reset()
    for _ in range(3):
        env.step_many(T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(4, 256 // T)
    e0.record()
    for _ in range(reps):
        env.step_many(T)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"case": "64x64 pure tick trail step_many", "T": T, "env_steps_per_s": N * T * reps / (ms * 1e-3)}), flush=True)
    del env
    torch.cuda.empty_cache()
