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
        out = []
        model, losses = DQN.train(n_envs=a.envs, iterations=max(1, a.ticks // DQN.GAME_CYCLE), log=lambda it, loss, st: out.append((loss, st)))
        loss, st = out[-1]
        print(json.dumps({"config": "DQN survivor loop (DQN.py:135-309 restated), %d envs, 1-plane f32 obs, GPU replay ring" % a.envs,
                          "ms_per_tick": {"q_forward(2N obs)": st["q_forward_ms_per_tick"], "select+env_step+replay_push": st["env_replay_ms_per_tick"]},
                          "env_fraction_of_tick": st["env_replay_ms_per_tick"] / (st["env_replay_ms_per_tick"] + st["q_forward_ms_per_tick"]),
                          "loss": loss, "episodes": st["episodes"], "mean_episode_ticks": st["ep_ticks"] / max(1, st["episodes"])}))
        sys.exit(0)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    logs = []
    agent, env = DDQN.train(n_envs=a.envs, env_steps=a.ticks, learn_every=a.learn_every, log=lambda t, d: logs.append(d), amp=a.amp)
    warm = logs[len(logs) // 4:]
    fw = sum(d["forward_ms"] for d in warm) / len(warm); er = sum(d["env_replay_ms"] for d in warm) / len(warm)
    lr = sum(d["learn_ms"] for d in warm) / len(warm)
    st = env.stats_dict()
    if rank == 0:
        print(json.dumps({"config": "DDQN self-play, %d envs/GPU x %d GPU(s), pop_up bf16 obs, GPU replay ring%s" % (a.envs, world, ", bf16 autocast forward" if a.amp else ""),
                          "ms_per_tick": {"q_forward(2N obs)": fw, "select+env_step+replay_push": er, "sample+learn(+allreduce)": lr},
                          "env_fraction_of_tick": er / (fw + er + lr), "env_steps_per_s_per_gpu": a.envs / ((fw + er + lr) * 1e-3),
                          "loss": logs[-1]["loss"], "episodes": st["episodes"], "mean_episode_ticks": st["ep_ticks"] / max(1, st["episodes"]),
                          "replay_len": len(agent.memory)}))
    if world > 1:
        torch.distributed.destroy_process_group()
