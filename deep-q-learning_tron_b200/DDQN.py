"""Drop-in for the replay side of the reference's DDQN.py -- ReplayBuffer (DDQN.py:167-203) on the GPU replay ring -- and a
batched restatement of Agent.learn / train (DDQN.py:73-165, 206-346): 65,536+ self-play games per GPU, epsilon-greedy from the
live Q-net, Double-DQN target, MSE, Adam, soft update; optional data-parallel gradient all-reduce over NCCL (one flat bucket)."""
import os

import torch

import tron_b200
from tron_b200.replay import ReplayRing

EPSILON_START, ESPILON_END, DECAY_RATE = 1, 0.003, 0.999  # DDQN.py:18-20
TAU = 0.001             # DDQN.py:21
MEM_CAPACITY = int(1e5)  # DDQN.py:26
UPDATE_EVERY = 4        # DDQN.py:29
GAME_CYCLE = 20         # DDQN.py:30
BATCH_SIZE = 64         # config.py:7
GAMMA = 0.9             # config.py:5


class ReplayBuffer:
    """add(state (1,P,R,C), action int, reward number, next_state, done bool); sample() -> (states, actions i64, rewards f32,
    next_states, dones f32) on the device, batch_size rows without replacement (DDQN.py:191-200)."""

    def __init__(self, action_size, buffer_size, batch_size, frame_dtype=torch.float32, device="cuda"):
        self.action_size = action_size
        self.buffer_size = int(buffer_size)
        self.batch_size = batch_size
        self._ring = None
        self._dt, self._dev = frame_dtype, device

    def _ensure(self, frame):
        if self._ring is None:
            self._ring = ReplayRing(self.buffer_size, tuple(frame.shape[1:]), self._dt, device=self._dev)

    def add(self, state, action, reward, next_state, done):
        state, next_state = torch.as_tensor(state), torch.as_tensor(next_state)
        self._ensure(state)
        self._ring.push(state, next_state, torch.tensor([int(action)], dtype=torch.uint8), torch.tensor([float(reward)]),
                        torch.tensor([int(bool(done))], dtype=torch.uint8))

    def add_batch(self, state, action, reward, next_state, done, done_stride=1):
        self._ensure(state)
        self._ring.push(state, next_state, action, reward, done, done_stride)

    def sample(self, out_dtype=torch.float32):
        return self._ring.sample(self.batch_size, out_dtype)

    def __len__(self):
        return 0 if self._ring is None else len(self._ring)


class Agent:
    """Batched counterpart of DDQN.Agent (DDQN.py:34-165)."""

    def __init__(self, in_planes=3, device="cuda", buffer_size=1 << 20, batch_size=BATCH_SIZE, frame_dtype=torch.bfloat16, lr=1e-3,
                 data_parallel=False):
        from Net.DQNNet import Net
        self.device = torch.device(device)
        self.qnetwork_local = Net(in_planes).to(self.device)
        self.qnetwork_target = Net(in_planes).to(self.device)
        self.qnetwork_target.load_state_dict(self.qnetwork_local.state_dict())
        self.optimizer = torch.optim.Adam(self.qnetwork_local.parameters(), lr=lr)
        self.memory = ReplayBuffer(4, buffer_size, batch_size, frame_dtype, device=self.device)
        self.epsilon = 0.0
        self.totalloss, self.steps = 0.0, 0
        self.data_parallel = data_parallel
        if data_parallel:  # identical initial weights on every rank
            import torch.distributed as dist
            for p in list(self.qnetwork_local.parameters()) + list(self.qnetwork_target.parameters()):
                dist.broadcast(p.data, 0)

    def q_values(self, obs, amp=False):
        """acting forward; amp=True runs it under bf16 autocast (cuDNN/cuBLAS tensor-core paths, observations are already bf16)"""
        self.qnetwork_local.eval()
        with torch.no_grad():
            if amp:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    q = self.qnetwork_local(obs).float()
            else:
                q = self.qnetwork_local(obs)
        self.qnetwork_local.train()
        return q

    def learn(self, experiences, gamma=GAMMA):
        states, actions, rewards, next_state, dones = experiences
        self.qnetwork_local.train(); self.qnetwork_target.eval()
        predicted = self.qnetwork_local(states).gather(1, actions)
        self.qnetwork_local.eval()
        with torch.no_grad():  # Double DQN: argmax from the local net, value from the target net (DDQN.py:131-134)
            a_loc = self.qnetwork_local(next_state).max(1)[1].unsqueeze(1)
            labels_next = self.qnetwork_target(next_state).gather(1, a_loc)
        self.qnetwork_local.train()
        labels = rewards + gamma * labels_next * (1 - dones)
        loss = torch.nn.functional.mse_loss(predicted, labels)
        self.optimizer.zero_grad()
        loss.backward()
        if self.data_parallel:
            allreduce_gradients(self.qnetwork_local)
        self.optimizer.step()
        with torch.no_grad():  # soft update (DDQN.py:154-165)
            for tp, lp in zip(self.qnetwork_target.parameters(), self.qnetwork_local.parameters()):
                tp.mul_(1 - TAU).add_(lp, alpha=TAU)
        self.totalloss += float(loss.detach()); self.steps += 1
        return loss.detach()


def allreduce_gradients(model):
    """One NCCL all-reduce of the flattened gradient (501,924 fp32 = 2.0 MB for the DQN net), averaged over ranks."""
    import torch.distributed as dist
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def train(n_envs=65536, env_steps=64, device=None, seed=0, obs_dtype=torch.bfloat16, learn_every=UPDATE_EVERY, data_parallel=None, log=None,
          amp=False):
    """Batched DDQN loop (DDQN.py:206-346 restated): every tick both players of all envs act epsilon-greedily from the
    local net, 2*n_envs transitions go into the GPU ring, and every `learn_every` ticks one Double-DQN learn step runs.
    Under torchrun each rank owns n_envs envs (env_id_base = rank * n_envs) and gradients are all-reduced."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if data_parallel is None:
        data_parallel = world > 1
    if device is None:
        device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    if data_parallel and not dist.is_initialized():
        dist.init_process_group("nccl")
    torch.manual_seed(seed)
    agent = Agent(3, device, frame_dtype=obs_dtype, data_parallel=data_parallel)
    if obs_dtype == torch.bfloat16:
        pass  # observations stay bf16 in the ring; the net computes in fp32 (conv weights), cast happens in forward()
    env = tron_b200.BatchedTron(n_envs, 10, 10, device=device, obs_dtype=obs_dtype, obs_enc="popup3", reward="ddqn", seed=seed,
                                env_id_base=rank * n_envs)
    obs = env.reset()
    epsilon = float(EPSILON_START)
    t_env = t_learn = 0.0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for t in range(env_steps):
        ev[0].record()
        q = agent.q_values(obs.view(2 * n_envs, 3, 12, 12), amp=amp)
        ev[1].record()
        act = env.select_actions(q, epsilon, counter=env.counter)
        res = env.step(act)
        agent.memory.add_batch(obs.view(2 * n_envs, 3, 12, 12), act.view(-1), res.reward.view(-1), res.obs.view(2 * n_envs, 3, 12, 12), res.done, done_stride=2)
        obs = res.obs
        ev[2].record()
        if (t + 1) % learn_every == 0 and len(agent.memory) > agent.memory.batch_size:
            agent.learn(agent.memory.sample())
        ev[3].record()
        if (t + 1) % GAME_CYCLE == 0 and epsilon * DECAY_RATE > ESPILON_END:
            epsilon *= DECAY_RATE
        if log:
            torch.cuda.synchronize()
            log(t, dict(forward_ms=ev[0].elapsed_time(ev[1]), env_replay_ms=ev[1].elapsed_time(ev[2]), learn_ms=ev[2].elapsed_time(ev[3]),
                        loss=agent.totalloss / max(agent.steps, 1)))
    torch.cuda.synchronize()
    return agent, env
