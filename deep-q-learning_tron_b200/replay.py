"""GPU replay ring behind the reference's replay containers (DQN.py:81-132 ReplayMemory, DDQN.py:167-203 ReplayBuffer).

Storage is a ring of transitions in HBM; push and gather are CUDA kernels (replay_push / replay_gather),
uniform sampling without replacement is replay_sample_indices (Floyd's algorithm on Philox).
"""
import ctypes as C

import torch

from . import _abi as abi
from . import _lib
from .batch_env import _CODE_OF, _TORCH_OF


class ReplayRing:
    def __init__(self, capacity, frame_shape, frame_dtype=torch.bfloat16, device="cuda", seed=0):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.capacity = int(capacity)
        self.frame_shape = tuple(frame_shape)
        self.F = 1
        for d in self.frame_shape:
            self.F *= int(d)
        self.dt = frame_dtype if isinstance(frame_dtype, int) else _CODE_OF[frame_dtype]
        tdt = _TORCH_OF[self.dt]
        self.state = torch.zeros((self.capacity, self.F), dtype=tdt, device=self.device)
        self.next_state = torch.zeros((self.capacity, self.F), dtype=tdt, device=self.device)
        self.action = torch.zeros(self.capacity, dtype=torch.uint8, device=self.device)
        self.reward = torch.zeros(self.capacity, dtype=torch.float32, device=self.device)
        self.done = torch.zeros(self.capacity, dtype=torch.uint8, device=self.device)
        self.cursor = 0  # total transitions ever pushed
        self.seed, self.sample_counter = int(seed), 0
        self.ring = abi.ReplayRing(struct_size=C.sizeof(abi.ReplayRing), frame_elems=self.F, frame_dtype=self.dt, capacity=self.capacity,
                                   state=self.state.data_ptr(), next_state=self.next_state.data_ptr(), action=self.action.data_ptr(),
                                   reward=self.reward.data_ptr(), done=self.done.data_ptr())

    def __len__(self):
        return min(self.cursor, self.capacity)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def push(self, state, next_state, action, reward, done, done_stride=1):
        """Append n transitions.  state/next_state: [n, *frame_shape] (ring dtype), action uint8 [n], reward f32 [n],
        done uint8 [n] (done_stride=1) or [n/2] (done_stride=2: one flag per env, two players per env)."""
        n = action.numel()
        if state.numel() != n * self.F or next_state.numel() != n * self.F or reward.numel() != n:
            raise ValueError("push: state/next_state must hold %d frames of %d elements and reward %d values" % (n, self.F, n))
        if done_stride not in (1, 2) or done.numel() * done_stride != n:
            raise ValueError("push: done must hold n/done_stride flags (n=%d, done_stride=%d, got %d)" % (n, done_stride, done.numel()))
        tdt = _TORCH_OF[self.dt]
        state = state.to(device=self.device, dtype=tdt).contiguous()
        next_state = next_state.to(device=self.device, dtype=tdt).contiguous()
        action = action.to(device=self.device, dtype=torch.uint8).contiguous()
        reward = reward.to(device=self.device, dtype=torch.float32).contiguous()
        done = done.to(device=self.device, dtype=torch.uint8).contiguous()
        off = 0
        while off < n:  # the ABI takes at most `capacity` transitions per call
            m = min(self.capacity, n - off)
            fo = off * self.F
            with torch.cuda.device(self.device):
                _lib.check(self.lib.replay_push(C.byref(self.ring), self.cursor, state.view(-1)[fo:].data_ptr(),
                                                next_state.view(-1)[fo:].data_ptr(), action[off:].data_ptr(), reward.view(-1)[off:].data_ptr(),
                                                done[(off // done_stride):].data_ptr(), done_stride, m, self._stream()), "replay_push")
            self.cursor += m
            off += m

    def sample_indices(self, k, counter=None):
        if counter is None:
            counter, self.sample_counter = self.sample_counter, self.sample_counter + 1
        if not 0 < k <= min(len(self), 4096):
            raise ValueError("sample_indices: need 0 < k <= min(len(ring)=%d, 4096), got %d" % (len(self), k))
        idx = torch.empty(k, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.replay_sample_indices(len(self), k, self.seed, counter, idx.data_ptr(), self._stream()),
                       "replay_sample_indices")
        return idx

    def gather(self, idx, out_dtype=torch.float32):
        """-> (states [k,*frame], actions i64 [k,1], rewards f32 [k,1], next_states, dones f32 [k,1])  (DDQN.py:191-200)"""
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        k = idx.numel()
        if out_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("gather: out_dtype must be float32 or bfloat16")
        s = torch.empty((k,) + self.frame_shape, dtype=out_dtype, device=self.device)
        s2 = torch.empty_like(s)
        a = torch.empty((k, 1), dtype=torch.int64, device=self.device)
        r = torch.empty((k, 1), dtype=torch.float32, device=self.device)
        d = torch.empty((k, 1), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.replay_gather(C.byref(self.ring), idx.data_ptr(), k, s.data_ptr(), s2.data_ptr(), _CODE_OF[out_dtype],
                                              a.data_ptr(), r.data_ptr(), d.data_ptr(), self._stream()), "replay_gather")
        return s, a, r, s2, d

    def sample(self, k, out_dtype=torch.float32):
        return self.gather(self.sample_indices(k), out_dtype)
