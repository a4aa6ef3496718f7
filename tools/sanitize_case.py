#!/usr/bin/env python
"""Small run through every kernel family, meant to be wrapped by `compute-sanitizer --tool memcheck`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tron_b200  # noqa: E402
from tron_b200 import _lib, abi  # noqa: E402
from tron_b200.batch_env import BatchedTron, HostTron  # noqa: E402
from tron_b200.replay import ReplayRing  # noqa: E402

for layout in ("tile8", "bits10"):
    for dt in (torch.bfloat16, torch.float32, torch.int8):
        for enc in ("lut1", "popup3", "popup3_const", "none"):
            env = BatchedTron(333, 10, 10, obs_dtype=dt, obs_enc=enc, layout=layout, seed=1)
            env.reset()
            for _ in range(3):
                env.step()
            env.step_many(3)
            env.export()
for W, n in ((3, 77), (7, 130), (31, 50), (64, 20), (126, 3)):
    for enc in ("lut1", "none"):
        env = BatchedTron(n, W, W, obs_dtype=torch.bfloat16, obs_enc=enc, seed=2, slide_mode="ice")
        env.reset()
        for _ in range(4):
            env.step()
_lib.load().tron_set_option(abi.OPT_SPARSE_MIN_CELLS, 0)
env = BatchedTron(999, 12, 12, obs_enc="none", seed=3)
env.reset()
for _ in range(20):
    env.step()
env.step_many(5)
_lib.load().tron_set_option(abi.OPT_SPARSE_MIN_CELLS, 1024)
q = torch.randn(999, 2, 4, device="cuda")
env.select_actions(q, 0.3, counter=1); env.random_actions(2)
ring = ReplayRing(1000, (3, 12, 12), torch.bfloat16)
s = torch.zeros((600, 3, 12, 12), dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    ring.push(s, s, torch.zeros(600, dtype=torch.uint8), torch.zeros(600), torch.zeros(300, dtype=torch.uint8), done_stride=2)
ring.sample(64); ring.sample(64, torch.bfloat16)
h = HostTron(1000, 10, 10, n_chunks=3)
h.reset(); h.step(); h.close()
torch.cuda.synchronize()
print("sanitize case done")
