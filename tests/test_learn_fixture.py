"""the learn steps that consume the replay output -- DDQN.Agent.learn (reference DDQN.py:115-165: Double-DQN
target, MSE, Adam, soft update tau) and the DQN learn block (reference DQN.py:263-292: smooth-L1, target r or r + gamma max Q) --
against tests/golden/learn.npz, which tests/golden/make_learn_golden.py produced by running the reference's own code on fixed
weights and batches.  Runs on the CPU (incl. the dropout variant: same torch CPU generator stream) and, marked gpu, on the B200."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tron_b200  # noqa: E402,F401
from tron_b200 import dropin  # noqa: E402

from _golden import fill_params, load_npz, summarize  # noqa: E402

RTOL = 1e-5   # fp32; the only arithmetic difference to the reference is F.mish (fused) vs x * tanh(softplus(x))


_SHIMS = ("tron", "config", "DQN", "DDQN", "Net")


@pytest.fixture(autouse=True, scope="module")
def _dropin_on_path():
    """the drop-in mirrors are importable only while this module runs (other tests import the reference's modules of the same names)"""
    saved = list(sys.path)
    for m in [k for k in sys.modules if k.split(".")[0] in _SHIMS]:
        del sys.modules[m]
    yield
    for m in [k for k in sys.modules if k.split(".")[0] in _SHIMS]:
        del sys.modules[m]
    sys.path[:] = saved


def _mods():
    dropin.install()
    import DDQN
    import DQN
    from Net.DQNNet import Net
    return DDQN, DQN, Net


def _close(name, got, fx, key, atol_scale=1.0):
    """compare a tensor with its stored signature (sums + strided sample)"""
    sums, sample = summarize(got.detach().cpu().numpy())
    want_sums, want_sample = fx[key + "/sums"], fx[key + "/sample"]
    scale = max(float(np.abs(want_sample).max()), 1e-12)
    assert np.allclose(sample, want_sample, rtol=RTOL * 20, atol=RTOL * scale * atol_scale), (name, key, np.abs(sample - want_sample).max(), scale)
    assert abs(sums[1] - want_sums[1]) <= RTOL * 20 * max(want_sums[1], 1e-12), (name, key, sums, want_sums)


def _ddqn_case(device, tag):
    DDQN, _, _ = _mods()
    fx = load_npz("learn.npz")
    agent = DDQN.Agent(in_planes=4, device=device, frame_dtype=torch.float32)
    fill_params(agent.qnetwork_local, 7)
    fill_params(agent.qnetwork_target, 8)
    p = 0.0 if tag == "nodrop" else 0.2
    agent.qnetwork_local.dropout.p = p; agent.qnetwork_target.dropout.p = p
    before = {k: v.detach().clone() for k, v in agent.qnetwork_local.named_parameters()}
    exp = (torch.from_numpy(fx["ddqn_states"].astype(np.float32)), torch.from_numpy(fx["ddqn_actions"]), torch.from_numpy(fx["ddqn_rewards"]),
           torch.from_numpy(fx["ddqn_next_states"].astype(np.float32)), torch.from_numpy(fx["ddqn_dones"]))
    exp = tuple(x.to(device) for x in exp)
    assert abs(float(fx["ddqn_gamma"]) - DDQN.GAMMA) < 1e-7 and abs(float(fx["ddqn_tau"]) - DDQN.TAU) < 1e-9
    torch.manual_seed(123)
    loss = agent.learn(exp, DDQN.GAMMA)
    want = float(fx["ddqn_%s_loss" % tag])
    assert abs(float(loss) - want) <= RTOL * abs(want), (float(loss), want)
    for k, prm in agent.qnetwork_local.named_parameters():
        _close("grad", prm.grad, fx, "ddqn_%s_grad_%s" % (tag, k))
        # Adam's first step is -lr * g / (|g| + eps): compare where the gradient is far from zero (elsewhere its sign is noise)
        g_ref = fx["ddqn_%s_grad_%s/sample" % (tag, k)]
        _, d_got = summarize((prm.detach() - before[k]).cpu().numpy())
        _, b0 = summarize(before[k].cpu().numpy())
        d_ref = fx["ddqn_%s_local1_%s/sample" % (tag, k)] - b0
        solid = np.abs(g_ref) > 1e-3 * np.abs(g_ref).max()
        assert np.allclose(d_got[solid], d_ref[solid], rtol=0, atol=2e-6), (k, np.abs(d_got[solid] - d_ref[solid]).max())
        assert np.abs(d_got).max() <= 1.001e-3  # lr
    for k, prm in agent.qnetwork_target.named_parameters():
        _close("target", prm, fx, "ddqn_%s_target1_%s" % (tag, k), atol_scale=0.1)


def _dqn_case(device):
    _, DQN, Net = _mods()
    fx = load_npz("learn.npz")
    model = Net(in_planes=4, batch_size=128, gamma=float(fx["dqn_gamma"])).to(device)
    fill_params(model, 11)
    model.dropout.p = 0.0
    opt = torch.optim.Adam(model.parameters())
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    batch = (torch.from_numpy(fx["dqn_states"].astype(np.float32)), torch.from_numpy(fx["dqn_actions"]).long(), torch.from_numpy(fx["dqn_rewards"]),
             torch.from_numpy(fx["dqn_next_states"].astype(np.float32)), torch.from_numpy(fx["dqn_terminal"].astype(np.float32)).unsqueeze(1))
    loss = DQN.learn_step(model, opt, tuple(x.to(device) for x in batch))
    want = float(fx["dqn_loss"])
    assert abs(float(loss) - want) <= RTOL * abs(want), (float(loss), want)
    for k, prm in model.named_parameters():
        _close("grad", prm.grad, fx, "dqn_grad_" + k)
        g_ref = fx["dqn_grad_%s/sample" % k]
        _, d_got = summarize((prm.detach() - before[k]).cpu().numpy())
        _, b0 = summarize(before[k].cpu().numpy())
        d_ref = fx["dqn_model1_%s/sample" % k] - b0
        solid = np.abs(g_ref) > 1e-3 * np.abs(g_ref).max()
        assert np.allclose(d_got[solid], d_ref[solid], rtol=0, atol=2e-6), (k, np.abs(d_got[solid] - d_ref[solid]).max())


@pytest.mark.parametrize("tag", ["nodrop", "drop"])
def test_ddqn_learn_matches_reference_cpu(tag):
    _ddqn_case("cpu", tag)


def test_dqn_learn_block_matches_reference_cpu():
    _dqn_case("cpu")


@pytest.mark.gpu
def test_learn_steps_match_reference_on_gpu():
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False  # compare at fp32, as the tolerance states
    try:
        _ddqn_case("cuda", "nodrop")
        _dqn_case("cuda")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
