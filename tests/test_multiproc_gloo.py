"""world_size-2 checks of the multi-GPU host logic on CPU (gloo): env sharding and the flat gradient all-reduce."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import tron_b200
    from tron_b200 import dropin
    dropin.install()
    import DDQN
    from tron_b200.sharding import env_shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4))
    x = torch.full((3, 6), float(rank + 1))
    model(x).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    DDQN.allreduce_gradients(model)
    gathered = [torch.zeros(sum(g.numel() for g in local)) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local]))
    want = torch.stack(gathered).mean(0)
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    ok = bool(torch.allclose(got, want, atol=1e-6))
    # the Agent keeps every gradient in one flat buffer (parameters' .grad are views): one all-reduce, no packing
    torch.manual_seed(1)
    agent = DDQN.Agent(in_planes=3, device="cpu", frame_dtype=torch.float32, data_parallel=True)
    g = torch.Generator().manual_seed(10 + rank)
    exp = (torch.randn(8, 3, 12, 12, generator=g), torch.randint(0, 4, (8, 1), generator=g), torch.randn(8, 1, generator=g),
           torch.randn(8, 3, 12, 12, generator=g), torch.zeros(8, 1))
    agent.qnetwork_local.dropout.p = 0.0
    w0 = [p.detach().clone() for p in agent.qnetwork_local.parameters()]
    agent._backward(exp, DDQN.GAMMA)
    mine = agent.flat_grad.clone()
    assert all(p.grad.data_ptr() >= agent.flat_grad.data_ptr() for p in agent.qnetwork_local.parameters())
    DDQN.allreduce_gradients(agent.qnetwork_local, agent.flat_grad)
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    ok = ok and bool(torch.allclose(agent.flat_grad, torch.stack(both).mean(0), atol=1e-6)) and float(mine.abs().sum()) > 0
    ok = ok and bool(torch.allclose(torch.cat([p.grad.reshape(-1) for p in agent.qnetwork_local.parameters()]), agent.flat_grad))
    agent._apply()  # identical averaged gradients + identical initial weights (broadcast) -> identical weights on both ranks
    w1 = torch.cat([p.detach().reshape(-1) for p in agent.qnetwork_local.parameters()])
    ws = [torch.zeros_like(w1) for _ in range(world)]
    dist.all_gather(ws, w1)
    ok = ok and bool(torch.equal(ws[0], ws[1])) and not bool(torch.equal(w1, torch.cat([w.reshape(-1) for w in w0])))
    q.put((rank, ok, env_shard(1000003)))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    (b0, n0), (b1, n1) = res[0][2], res[1][2]
    assert b0 == 0 and b1 == n0 and n0 + n1 == 1000003


def test_sharded_rng_streams_match_single_shard_on_the_oracle():
    """The design property the GPU path relies on (tests/test_gpu_parity.py::test_sharding_is_invisible), on the CPU oracle."""
    sys.path.insert(0, ROOT)
    from oracle import c_oracle as oc
    from tron_b200 import abi
    from tron_b200.sharding import env_shard
    N = 1001
    full = oc.OracleEnv(N, 10, 10, obs_dtype=abi.I8, seed=4)
    shards = [oc.OracleEnv(env_shard(N, r, 2)[1], 10, 10, obs_dtype=abi.I8, seed=4, env_id_base=env_shard(N, r, 2)[0]) for r in range(2)]
    a = full.reset(); b = np.concatenate([s.reset() for s in shards])
    assert np.array_equal(a, b)
    for _ in range(10):
        f = full.step(); parts = [s.step() for s in shards]
        for i in range(5):
            assert np.array_equal(f[i], np.concatenate([p[i] for p in parts]))


def test_dropin_game_fails_loudly_without_gpu():
    import pytest
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); import tron_b200; from tron_b200 import dropin, _lib; dropin.install()\n"
            "from tron.game import Game, PositionPlayer; from tron.player import ACPlayer\n"
            "try:\n    Game(10, 10, [PositionPlayer(1, ACPlayer(), [1, 1]), PositionPlayer(2, ACPlayer(), [5, 5])])\n"
            "except _lib.TronError as e:\n    print('LOUD', e)\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "LOUD" in out.stdout and "no CPU fallback" in out.stdout, out.stdout + out.stderr
