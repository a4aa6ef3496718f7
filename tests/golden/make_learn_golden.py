#!/usr/bin/env python
"""Generate tests/golden/learn.npz by RUNNING THE REFERENCE'S OWN LEARN STEPS (authoring container only; needs /root/reference).

  DDQN:  Agent.learn (DDQN.py:115-151) + soft_update (DDQN.py:154-165) on fixed weights and a fixed batch of 64 transitions
         -> loss, gradients of the local net, Adam-updated local parameters, soft-updated target parameters.
         Twice: with dropout disabled (p = 0, device independent) and with the reference's dropout p = 0.2 under
         torch.manual_seed(123) on the CPU generator (compared by the CPU test, same generator stream).
  DQN:   the learn block inside train() (DQN.py:263-292: smooth-L1, target r or r + gamma * max Q(s')) is not a function in the
         reference; the generator reads exactly those source lines from the reference file at run time and executes them
         (nothing is copied into this repository) on a fixed model / batch of 128 transitions -> loss, gradients, updated parameters.

The only patch applied to the reference is the missing activation: Net.mish is assigned in Net.__init__ (Net/DQNNet.py:31) but
never defined, so Net() cannot be constructed; the generator defines it as x * tanh(softplus(x)) (Net/ACNet.py:56-57).

Usage:  python tests/golden/make_learn_golden.py
"""
import os
import random
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from make_golden import REF, import_reference  # noqa: E402
from _golden import fill_params, summarize  # noqa: E402


def put(out, key, tensor):
    """store the compact signature of a tensor (tests/_golden.py summarize): the fixture stays a few hundred KB"""
    sums, sample = summarize(tensor.detach().cpu().numpy())
    out[key + "/sums"] = sums
    out[key + "/sample"] = sample


def main():
    import_reference()
    import torch
    import torch.nn.functional as F
    from Net.DQNNet import Net
    Net.mish = staticmethod(lambda x: x * torch.tanh(F.softplus(x)))  # the one repair (see module docstring)
    import DDQN as RD
    import DQN as RQ

    out = {}
    rng = np.random.default_rng(0)
    obs_values = np.array([0, 0, 0, 0, 1, 1, 10, 5], np.float32)  # pop_up planes hold 0 / 1 / 10, the const plane 5

    # ------------------------------------------------------------------ DDQN Agent.learn
    B = 64
    states = rng.choice(obs_values, size=(B, 4, 12, 12)).astype(np.float32)
    next_states = rng.choice(obs_values, size=(B, 4, 12, 12)).astype(np.float32)
    actions = rng.integers(0, 4, size=(B, 1)).astype(np.int64)
    rewards = rng.choice(np.array([-1, -1, -1, 100, -100, 0], np.float32), size=(B, 1)).astype(np.float32)
    dones = (rewards != -1).astype(np.float32)
    out.update(ddqn_states=states.astype(np.int8), ddqn_next_states=next_states.astype(np.int8), ddqn_actions=actions, ddqn_rewards=rewards, ddqn_dones=dones)
    for tag, p_drop in (("nodrop", 0.0), ("drop", 0.2)):
        agent = RD.Agent()
        fill_params(agent.qnetwork_local, 7)    # initial weights: numpy PCG64, reproduced by the tests (not stored)
        fill_params(agent.qnetwork_target, 8)
        agent.qnetwork_local.dropout.p = p_drop
        agent.qnetwork_target.dropout.p = p_drop
        exp = tuple(torch.from_numpy(x) for x in (states, actions, rewards, next_states, dones))
        torch.manual_seed(123)
        agent.learn(exp, RD.GAMMA)
        out["ddqn_%s_loss" % tag] = np.float32(float(agent.totalloss))
        for k, p in agent.qnetwork_local.named_parameters():
            put(out, "ddqn_%s_grad_%s" % (tag, k), p.grad)
            put(out, "ddqn_%s_local1_%s" % (tag, k), p)
        for k, p in agent.qnetwork_target.named_parameters():
            put(out, "ddqn_%s_target1_%s" % (tag, k), p)
    out["ddqn_gamma"] = np.float32(RD.GAMMA); out["ddqn_tau"] = np.float32(RD.TAU)

    # ------------------------------------------------------------------ DQN learn block (DQN.py:263-292), executed from the reference file
    lines = open(os.path.join(REF, "DQN.py")).read().split("\n")
    start = next(i for i, ln in enumerate(lines) if "transitions = memory.sample(" in ln)
    stop = next(i for i, ln in enumerate(lines) if i > start and ln.strip() == "optimizer.step()")
    block = textwrap.dedent("\n".join(lines[start:stop + 1]))
    assert (start, stop) == (262, 291), (start, stop)  # 0-based line indices: DQN.py:263-292
    B = 128
    model = Net()
    fill_params(model, 11)
    model.dropout.p = 0.0
    model.batch_size, model.gamma = B, RQ.GAMMA  # the reference's train() reads these off the model (DQN.py:263,278)
    s = rng.choice(np.array([1, 1, 1, -1, -2, -3, 10, -10], np.float32), size=(B, 4, 12, 12)).astype(np.float32)
    s2 = rng.choice(np.array([1, 1, 1, -1, -2, -3, 10, -10], np.float32), size=(B, 4, 12, 12)).astype(np.float32)
    a = rng.integers(0, 4, size=(B, 1)).astype(np.float32)       # DQN.py:216 stores actions as (1,1) float tensors
    r = rng.choice(np.array([1, 2, 3, 5, 100, -25, 0], np.float32), size=(B, 1)).astype(np.float32)
    term = (r[:, 0] >= 100) | (r[:, 0] == -25) | (r[:, 0] == 0)
    out.update(dqn_states=s.astype(np.int8), dqn_next_states=s2.astype(np.int8), dqn_actions=a, dqn_rewards=r, dqn_terminal=term.astype(np.uint8))
    transitions = [RQ.Transition(torch.from_numpy(s[i:i + 1]), torch.from_numpy(a[i:i + 1]), torch.from_numpy(s2[i:i + 1]),
                                 torch.from_numpy(r[i:i + 1]), bool(term[i])) for i in range(B)]

    class FixedMemory(list):
        def sample(self, k):
            assert k == len(self)
            return list(self)
    ns = dict(memory=FixedMemory(transitions), model=model, optimizer=torch.optim.Adam(model.parameters()), Transition=RQ.Transition,
              torch=torch, F=F, device="cpu")
    exec(compile(block, os.path.join(REF, "DQN.py"), "exec"), ns)
    out["dqn_loss"] = np.float32(float(ns["loss"]))
    for k, p in model.named_parameters():
        put(out, "dqn_grad_" + k, p.grad)
        put(out, "dqn_model1_" + k, p)
    out["dqn_gamma"] = np.float32(RQ.GAMMA)

    np.savez_compressed(os.path.join(HERE, "learn.npz"), **out)
    print("wrote learn.npz: %d arrays, ddqn loss %.6f / %.6f (dropout), dqn loss %.6f" %
          (len(out), out["ddqn_nodrop_loss"], out["ddqn_drop_loss"], out["dqn_loss"]))


if __name__ == "__main__":
    random.seed(0)
    main()
