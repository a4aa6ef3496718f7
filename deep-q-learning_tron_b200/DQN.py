"""Drop-in for the replay side of the reference's DQN.py: Transition (DQN.py:78) and ReplayMemory (DQN.py:81-132),
stored in the GPU replay ring (replay_push / replay_gather kernels), plus a batched restatement of the survivor
training loop (DQN.py:135-309) on the vectorised environment."""
import random
from collections import namedtuple

import numpy as np
import torch

from tron.player import Direction, Player

import tron_b200
from tron_b200.replay import ReplayRing

device = 'cuda' if torch.cuda.is_available() else 'cpu'  # DQN.py:16
MEM_CAPACITY = 10000  # DQN.py:32
BATCH_SIZE = 128      # DQN.py:19
GAMMA = 0.9           # DQN.py:20
EPSILON_START, ESPILON_END, DECAY_RATE = 1, 0.003, 0.999  # DQN.py:23-25
GAME_CYCLE = 20       # DQN.py:35

class Ai(Player):
    """Drop-in for DQN.Ai (DQN.py:39-75): a player that owns a Q-net and picks epsilon-greedy moves from its own 1-plane observation.
    (In the reference's fork Game can no longer step such a player -- SURVEY section 0; the drop-in Game can.)"""

    def __init__(self, epsilon=0, net=None):
        super(Ai, self).__init__()
        from Net.DQNNet import Net
        self.net = net if net is not None else Net(in_planes=1).to(device)
        self.epsilon = epsilon

    def action(self, map, id):
        game_map = map.state_for_player(id)
        x = torch.from_numpy(np.reshape(game_map, (1, 1, game_map.shape[0], game_map.shape[1]))).float()
        with torch.no_grad():
            next_action = int(torch.argmax(self.net(x), 1)[0]) + 1
        if random.random() <= self.epsilon:
            next_action = random.randint(1, 4)
        return Direction(next_action)


Transition = namedtuple('Transition', ('old_state', 'action', 'new_state', 'reward', 'terminal'))


class ReplayMemory(object):
    """push(old_state (1,1,R,C) f32, action (1,1) f32, new_state, reward (1,1) f32, terminal bool); sample(k) -> [Transition]."""

    def __init__(self, capacity, frame_dtype=torch.float32, device="cuda"):
        self.capacity = capacity
        self._ring = None
        self._dt, self._dev = frame_dtype, device

    def _ensure(self, frame):
        if self._ring is None:
            self._ring = ReplayRing(self.capacity, tuple(frame.shape[1:]), self._dt, device=self._dev)

    def push(self, *args):
        old_state, action, new_state, reward, terminal = args
        self._ensure(old_state)
        self._ring.push(old_state, new_state, action.reshape(-1), reward.reshape(-1), torch.tensor([int(bool(terminal))], dtype=torch.uint8))

    def push_batch(self, old_state, action, new_state, reward, terminal, done_stride=1):
        self._ensure(old_state)
        self._ring.push(old_state, new_state, action, reward, terminal, done_stride)

    @property
    def position(self):
        return 0 if self._ring is None else self._ring.cursor % self.capacity

    def sample_batch(self, batch_size):
        """-> (old_state [k,...] f32, action i64 [k,1], reward f32 [k,1], new_state, terminal f32 [k,1]) on the device"""
        return self._ring.sample(batch_size)

    def sample(self, batch_size):
        s, a, r, s2, d = self.sample_batch(batch_size)
        return [Transition(s[i:i + 1], a[i:i + 1].float(), s2[i:i + 1], r[i:i + 1], bool(d[i, 0].item())) for i in range(batch_size)]

    def __len__(self):
        return 0 if self._ring is None else len(self._ring)


def train(model=None, n_envs=4096, iterations=100, learn_steps_per_iter=1, device="cuda", seed=0, log=None):
    """Batched restatement of DQN.train (DQN.py:135-309): self-play with the survivor reward (step index / 100 / -25 / 0),
    1-plane observations, one smooth-L1 learn step per cycle on a uniform sample, target r or r + gamma * max Q(s')."""
    import torch.nn.functional as F
    from Net.DQNNet import Net
    model = model or Net(in_planes=1, batch_size=BATCH_SIZE, gamma=GAMMA).to(device)
    opt = torch.optim.Adam(model.parameters())
    env = tron_b200.BatchedTron(n_envs, 10, 10, device=device, obs_dtype=torch.float32, obs_enc="lut1", reward="survivor", seed=seed)
    mem = ReplayMemory(max(MEM_CAPACITY, 2 * n_envs * GAME_CYCLE), device=device)
    obs = env.reset()
    epsilon = float(EPSILON_START)
    losses = []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_fwd = t_env = 0.0
    for it in range(iterations):
        for _ in range(GAME_CYCLE):
            ev[0].record()
            with torch.no_grad():
                q = model(obs.view(2 * n_envs, 1, 12, 12))
            ev[1].record()
            act = env.select_actions(q, epsilon, counter=env.counter)
            res = env.step(act)
            mem.push_batch(obs.view(2 * n_envs, 1, 12, 12), act.view(-1), res.obs.view(2 * n_envs, 1, 12, 12), res.reward.view(-1), res.done, done_stride=2)
            obs = res.obs
            ev[2].record()
            if log:
                torch.cuda.synchronize()
                t_fwd += ev[0].elapsed_time(ev[1]); t_env += ev[1].elapsed_time(ev[2])
            if epsilon * DECAY_RATE > ESPILON_END:
                epsilon *= DECAY_RATE
        for _ in range(learn_steps_per_iter):
            s, a, r, s2, d = mem.sample_batch(min(len(mem), model.batch_size))
            pred = model(s).gather(1, a).squeeze(1)
            with torch.no_grad():
                target = r.squeeze(1) + (1 - d.squeeze(1)) * model.gamma * model(s2).max(1)[0]
            loss = F.smooth_l1_loss(pred, target)
            model.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        if log:
            st = env.stats_dict()
            st.update(q_forward_ms_per_tick=t_fwd / ((it + 1) * GAME_CYCLE), env_replay_ms_per_tick=t_env / ((it + 1) * GAME_CYCLE))
            log(it, losses[-1], st)
    return model, losses
