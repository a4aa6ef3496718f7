#!/usr/bin/env python
"""BASELINE configs #3 / #4: batched DDQN self-play loop with the GPU replay ring.

  python tools/train_ddqn.py --envs 65536 --ticks 40                                   # config #3 (1 GPU)
  torchrun --nproc-per-node 8 tools/train_ddqn.py --envs 131072 --ticks 40            # config #4 (1M envs, NCCL grad all-reduce)
Prints per-phase device time so "bottlenecked on the network, not the environment" can be read off directly.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tron_b200  # noqa: E402
from tron_b200 import dropin  # noqa: E402

dropin.install()
import DDQN  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--ticks", type=int, default=40)
    ap.add_argument("--learn-every", type=int, default=4)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast for the acting forward")
    ap.add_argument("--algo", default="ddqn", choices=["ddqn", "dqn"], help="ddqn: DDQN.py loop (pop_up obs); dqn: DQN.py survivor loop (1-plane obs)")
    a = ap.parse_args()
    if a.algo == "dqn":  # BASELINE config #3 as worded: DQN.py survivor loop, GPU replay, batched self-play
        import DQN
        tm, cyc = {}, []
        model, losses = DQN.train(n_envs=a.envs, iterations=max(1, a.ticks // DQN.GAME_CYCLE), timings=tm, on_cycle=lambda c, st: cyc.append(st))
        print(json.dumps({"config": "DQN survivor loop (DQN.py:135-309 restated), %d envs, 1-plane f32 obs, frame-sharing GPU replay" % a.envs,
                          "ms_per_tick": {"q_forward(2N obs)": tm["q_forward_ms"] / tm["ticks"], "select+env_step (writes the replay ring)": tm["env_replay_ms"] / tm["ticks"]},
                          "learn_ms_per_step": tm["learn_ms"] / max(1, tm["learn_steps"]),
                          "env_fraction_of_tick": tm["env_replay_ms"] / (tm["env_replay_ms"] + tm["q_forward_ms"]),
                          "loss": losses[-1], "last_cycle": cyc[-1]}))
        sys.exit(0)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    tm, cyc = {}, []
    agent, env = DDQN.train(n_envs=a.envs, env_steps=a.ticks, learn_every=a.learn_every, amp=a.amp, timings=tm, warmup_steps=a.ticks // 4,
                            on_cycle=lambda c, st: cyc.append(st))
    st = env.stats_dict()
    if rank == 0:
        n = tm["ticks"]
        print(json.dumps({"config": "DDQN self-play, %d envs/GPU x %d GPU(s), pop_up bf16 obs, frame-sharing GPU replay%s" % (a.envs, world, ", bf16 autocast forward" if a.amp else ""),
                          "ms_per_tick": {"q_forward(2N obs)": tm["q_forward_ms"] / n, "select+env_step (writes the replay ring)": tm["env_replay_ms"] / n,
                                          "sample+learn": tm["learn_ms"] / n},
                          "learn_ms_per_step": tm["learn_ms"] / max(1, tm["learn_steps"]),
                          "allreduce_us_per_call": 1e3 * tm["allreduce_ms"] / max(1, tm["allreduce_calls"]),
                          "env_fraction_of_tick": tm["env_replay_ms"] / (tm["q_forward_ms"] + tm["env_replay_ms"] + tm["learn_ms"]),
                          "env_steps_per_s_per_gpu": a.envs * n / ((tm["q_forward_ms"] + tm["env_replay_ms"] + tm["learn_ms"]) * 1e-3),
                          "episodes": st["episodes"], "mean_episode_ticks": st["ep_ticks"] / max(1, st["episodes"]), "last_cycle": cyc[-1] if cyc else None}))
    if world > 1:
        torch.distributed.destroy_process_group()
