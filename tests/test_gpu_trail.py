"""TRON_LAYOUT_TRAIL (trail-list records, pure ticks) must be indistinguishable from the dense layouts through the API."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import c_oracle as oc  # noqa: E402
from tron_b200 import abi  # noqa: E402

from _golden import load_json, load_npz, ragged_to_tapes  # noqa: E402
from _gpu import GpuEnvNumpy, assert_same_state, assert_same_step, make_pair  # noqa: E402

L = "trail"
KW = dict(layout=L, obs_dtype=abi.I8, obs_enc=abi.ENC_NONE)


def test_kat_and_trajectories_trail():
    for name, case in load_json("kat.json").items():
        env = GpuEnvNumpy(1, 10, 10, auto_reset=False, **KW)
        env.reset(spawn=np.array([case["spawn"]], np.int8))
        for t, snap in enumerate(case["ticks"]):
            if t > 0:
                _, rew, done, winner, _ = env.step(np.array([case["actions"][t - 1]], np.uint8))
                assert bool(done[0]) == snap["done"] and int(winner[0]) == snap["winner"], name
            ex = env.export()
            assert (ex["tiles"][0] == np.array(snap["tiles"], np.int8)).all(), (name, t)
            assert ex["alive"][0].tolist() == [int(x) for x in snap["alive"]] and ex["heads"][0].tolist() == snap["pos"][0] + snap["pos"][1]
    tr = load_npz("traj.npz")
    length = tr["length"]; G = len(length)
    tape, _, row = ragged_to_tapes(length, tr["actions"])
    env = GpuEnvNumpy(G, 10, 10, auto_reset=False, **KW)
    env.reset(spawn=tr["spawn"])
    for t in range(tape.shape[0]):
        _, rew, done, winner, eplen = env.step(tape[t])
        live = row[t + 1] >= 0
        r = row[t + 1][live]
        ex = env.export()
        assert (ex["tiles"][live] == tr["tiles"][r]).all() and (done[live] == tr["done"][r]).all() and (winner[live] == tr["winner"][r]).all()
        assert (ex["alive"][live] == tr["alive"][r]).all() and (ex["heads"][live] == tr["pos"][r]).all()


def test_slide_tape_fixture_trail():
    sl = load_npz("slide.npz")
    length = sl["length"]; G = len(length)
    tape, stape, row = ragged_to_tapes(length, sl["actions"], sl["slide"])
    env = GpuEnvNumpy(G, 10, 10, auto_reset=False, slide_mode=abi.SLIDE_TAPE, **KW)
    env.reset(spawn=sl["spawn"])
    for t in range(tape.shape[0]):
        _, rew, done, winner, _ = env.step(tape[t], slide_tape=stape[t])
        live = row[t + 1] >= 0
        r = row[t + 1][live]
        assert (env.export()["tiles"][live] == sl["tiles"][r]).all() and (done[live] == sl["done"][r]).all() and (winner[live] == sl["winner"][r]).all()


@pytest.mark.parametrize("W,N,mode", [(64, 900, abi.SLIDE_NONE), (3, 500, abi.SLIDE_NONE), (10, 4096, abi.SLIDE_ICE), (21, 1000, abi.SLIDE_TEMPER), (126, 40, abi.SLIDE_NONE)])
def test_trail_rng_mode_matches_oracle(W, N, mode):
    g, o = make_pair(N, W, W, seed=60 + W, slide_mode=mode, slide_rate=0.3, env_id_base=5, **KW)
    if mode == abi.SLIDE_TEMPER:
        prm = np.stack([np.random.default_rng(1).integers(-30, 31, N), np.random.default_rng(2).integers(40, 102, N),
                        np.random.default_rng(3).integers(40, 102, N), np.zeros(N, np.int64)], 1).astype(np.int8)
        g.env.slide_params.copy_(torch.as_tensor(prm)); o.slide_params[...] = prm
    g.reset(); o.reset()
    for t in range(50):
        assert_same_step(g.step(), o.step(), "W=%d tick %d" % (W, t))
        if t % 10 == 9:
            assert_same_state(g, o, "tick %d" % t)
    assert np.array_equal(g.stats, o.stats)
    assert_same_step(g.step_many(12), o.step_many(12))
    assert_same_state(g, o)


def test_trail_long_episodes_beyond_the_hot_window():
    """wall-avoiding tapes -> trails far longer than the 12 ticks kept in registers"""
    rng = np.random.default_rng(8)
    N, W = 600, 14
    g, o = make_pair(N, W, W, auto_reset=False, **KW)
    sp = np.stack([rng.integers(0, 7, N), rng.integers(0, W, N), rng.integers(7, W, N), rng.integers(0, W, N)], 1).astype(np.int8)
    g.reset(spawn=sp); o.reset(spawn=sp)
    D = [(-1, 0), (0, 1), (1, 0), (0, -1)]
    longest = 0
    for t in range(90):
        ex = o.export()
        tiles, heads = ex["tiles"], ex["heads"].astype(int)
        act = np.zeros((N, 2), np.uint8)
        for e in range(N):
            for pl in range(2):
                r, c = heads[e, 2 * pl], heads[e, 2 * pl + 1]
                free = [a for a, (dr, dc) in enumerate(D) if 0 <= r + dr < W and 0 <= c + dc < W and tiles[e, r + 1 + dr, c + 1 + dc] == 0]
                act[e, pl] = free[rng.integers(0, len(free))] if free else 0
        assert_same_step(g.step(act), o.step(act), "tick %d" % t)
        longest = max(longest, int(ex["ep_len"].max()))
        if t % 15 == 14:
            assert_same_state(g, o, "tick %d" % t)
    assert longest >= 30


@pytest.mark.parametrize("W,slide", [(8, abi.SLIDE_NONE), (8, abi.SLIDE_TEMPER), (20, abi.SLIDE_ICE)])
def test_trail_step_many_sees_its_own_bitmap_updates(W, slide):
    """several ticks per launch with long trails: the occupancy bit an earlier tick of the SAME launch set for a cold list entry
    (a reduction performed in L2) must be seen by the later ticks' lookups, or a collision is missed and the lists run over"""
    N = 20000
    g, o = make_pair(N, W, W, seed=3, slide_mode=slide, slide_rate=0.15, policy=abi.POLICY_FREE_EPS, policy_epsilon=0.0, **KW)
    g.reset(); o.reset()
    longest = 0
    for rnd in range(8):
        r, want = g.step_many(7), o.step_many(7)
        assert_same_step(r, want, "round %d" % rnd)
        longest = max(longest, int(want[4].max()))
    assert_same_state(g, o)
    assert longest > 20  # trails well beyond the 12 entries kept in the hot words


@pytest.mark.parametrize("W,slide", [(8, abi.SLIDE_ICE), (10, abi.SLIDE_TEMPER)])
def test_trail_long_trails_with_slide_modes_tick_by_tick(W, slide):
    """the trigger of round 2's reset failure (tools/long_trail_case.py): 20,000 games under the free-neighbour policy with a slide
    mode -- lists of unequal length growing past the hot words, games with cold entries being reset from tick ~10 on"""
    N = 20000
    g, o = make_pair(N, W, W, seed=3, slide_mode=slide, slide_rate=0.15, policy=abi.POLICY_FREE_EPS, policy_epsilon=0.0, **KW)
    g.reset(); o.reset()
    for t in range(40):
        assert_same_step(g.step(), o.step(), "tick %d" % t)
    assert_same_state(g, o)


def test_trail_import_export_and_cross_layout():
    N, W = 700, 12
    a = GpuEnvNumpy(N, W, W, seed=5, **KW)
    b = GpuEnvNumpy(N, W, W, seed=5, layout="tile8", obs_dtype=abi.I8, obs_enc=abi.ENC_NONE)
    a.reset(); b.reset()
    for _ in range(11):
        for x, y in zip(a.step(), b.step()):
            assert (x is None and y is None) or np.array_equal(x, y)
    ex = a.export()
    for k, v in b.export().items():
        assert np.array_equal(ex[k], v), k
    c = GpuEnvNumpy(N, W, W, seed=5, **KW)  # dense export -> trail import continues identically
    c.import_(**b.export()); c.env.counter = b.env.counter
    for _ in range(11):
        for x, y in zip(b.step(), c.step()):
            assert (x is None and y is None) or np.array_equal(x, y)
    for k, v in b.export().items():
        assert np.array_equal(c.export()[k], v), k


@pytest.mark.parametrize("W,N,dt,enc", [(64, 300, abi.BF16, abi.ENC_LUT1), (10, 4096, abi.BF16, abi.ENC_POPUP3), (7, 777, abi.F32, abi.ENC_LUT1),
                                         (31, 500, abi.I8, abi.ENC_POPUP3_CONST), (126, 9, abi.BF16, abi.ENC_LUT1), (12, 1000, abi.F32, abi.ENC_POPUP3)])
def test_trail_fused_observations_match_oracle(W, N, dt, enc):
    g, o = make_pair(N, W, W, layout=L, obs_dtype=dt, obs_enc=enc, const_plane=5.0, seed=80 + W, slide_mode=abi.SLIDE_ICE, slide_rate=0.2)
    assert np.array_equal(g.reset(), o.reset())
    for t in range(30):
        assert_same_step(g.step(), o.step(), "W=%d tick %d" % (W, t))
    assert_same_state(g, o)
    assert np.array_equal(g.observe(), o.observe())
    assert_same_step(g.step_many(6), o.step_many(6))
    assert_same_step(g.step_many(6, obs_every_tick=False), o.step_many(6, obs_every_tick=False))
    assert_same_state(g, o)


def test_trail_explicit_tapes_config2_shape():
    N = 4096
    rng = np.random.default_rng(0)
    g, o = make_pair(N, 10, 10, layout=L, obs_dtype=abi.BF16, obs_enc=abi.ENC_LUT1)
    sp = rng.integers(0, 10, size=(N, 4)).astype(np.int8); sp[:, 2] = (sp[:, 0] + 1 + rng.integers(0, 9, N)) % 10
    assert np.array_equal(g.reset(spawn=sp), o.reset(spawn=sp))
    for t in range(64):
        act = rng.integers(0, 4, size=(N, 2)).astype(np.uint8)
        sp = rng.integers(0, 10, size=(N, 4)).astype(np.int8); sp[:, 2] = (sp[:, 0] + 1 + rng.integers(0, 9, N)) % 10
        assert_same_step(g.step(act, spawn=sp), o.step(act, spawn=sp), "tick %d" % t)
    assert_same_state(g, o)
    assert np.array_equal(g.stats, o.stats)
