"""The bit-plane layouts -- TRON_LAYOUT_BITS10 (two 128-bit planes per 10x10 game, game-major) and TRON_LAYOUT_BITS (three dense
plane arrays, any W*H <= 128, slide modes included) -- must be indistinguishable from the int8 layout through the API: the same
golden fixtures and oracle comparisons as tests/test_gpu_parity.py, bit for bit, including exported grids."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import c_oracle as oc  # noqa: E402
from tron_b200 import abi  # noqa: E402

from _golden import digest_with, load_json, load_npz, ragged_to_tapes  # noqa: E402
from _gpu import GpuEnvNumpy, assert_same_state, assert_same_step, make_pair  # noqa: E402

@pytest.fixture(params=["bits10", "bits"])
def L(request):
    return request.param


def one_env_of(L):
    def one_env(W, **kw):
        kw.setdefault("obs_dtype", abi.I8); kw.setdefault("auto_reset", False)
        return GpuEnvNumpy(1, W, W, layout=L, **kw)
    return one_env


def test_kat_table_bits10(L):
    for name, case in load_json("kat.json").items():
        env = one_env_of(L)(10)
        obs = env.reset(spawn=np.array([case["spawn"]], np.int8))
        for t, snap in enumerate(case["ticks"]):
            if t > 0:
                obs, rew, done, winner, _ = env.step(np.array([case["actions"][t - 1]], np.uint8))
                assert bool(done[0]) == snap["done"] and int(winner[0]) == snap["winner"], name
            ex = env.export()
            assert (ex["tiles"][0] == np.array(snap["tiles"], np.int8)).all(), (name, t)
            assert (obs[0, 0, 0] == np.array(snap["obs1"])).all() and (obs[0, 1, 0] == np.array(snap["obs2"])).all(), (name, t)
            assert ex["alive"][0].tolist() == [int(x) for x in snap["alive"]] and ex["heads"][0].tolist() == snap["pos"][0] + snap["pos"][1]


@pytest.mark.parametrize("seed", [0, 1])
def test_digest_bits10(L, seed):
    want = load_json("digests.json")[str(seed)]
    got = digest_with(one_env_of(L), seed, 1000)
    assert got[0] == want["sha256"] and got[1] == want["steps"] and got[2] == want["wins"]


def test_full_trajectories_bits10(L):
    tr = load_npz("traj.npz")
    length = tr["length"]; G = len(length)
    tape, _, row = ragged_to_tapes(length, tr["actions"])
    env = GpuEnvNumpy(G, 10, 10, obs_dtype=abi.I8, auto_reset=False, layout=L)
    obs = env.reset(spawn=tr["spawn"])
    assert (obs[:, 0, 0] == tr["obs1"][row[0]]).all()
    for t in range(tape.shape[0]):
        obs, rew, done, winner, eplen = env.step(tape[t])
        live = row[t + 1] >= 0
        r = row[t + 1][live]
        ex = env.export()
        assert (ex["tiles"][live] == tr["tiles"][r]).all()
        assert (obs[live, 0, 0] == tr["obs1"][r]).all() and (obs[live, 1, 0] == tr["obs2"][r]).all()
        assert (done[live] == tr["done"][r]).all() and (winner[live] == tr["winner"][r]).all()
        assert (ex["alive"][live] == tr["alive"][r]).all() and (ex["heads"][live] == tr["pos"][r]).all()


@pytest.mark.parametrize("dt,enc,ticks", [(abi.BF16, abi.ENC_LUT1, 256), (abi.F32, abi.ENC_LUT1, 48), (abi.I8, abi.ENC_LUT1, 48),
                                          (abi.BF16, abi.ENC_POPUP3, 48), (abi.F32, abi.ENC_POPUP3_CONST, 32), (abi.BF16, abi.ENC_NONE, 64)])
def test_config2_4096_envs_tapes_bits10(L, dt, enc, ticks):
    N = 4096
    rng = np.random.default_rng(0)
    g, o = make_pair(N, 10, 10, layout=L, obs_dtype=dt, obs_enc=enc, const_plane=5.0, reward="ddqn")

    def spawn_tape():
        s = rng.integers(0, 10, size=(N, 4)).astype(np.int8)
        while True:
            same = (s[:, 0] == s[:, 2]) & (s[:, 1] == s[:, 3])
            if not same.any():
                return s
            s[same, :2] = rng.integers(0, 10, size=(int(same.sum()), 2))
    sp = spawn_tape()
    a, b = g.reset(spawn=sp), o.reset(spawn=sp)
    if enc != abi.ENC_NONE:
        assert np.array_equal(a, b)
    for t in range(ticks):
        act = rng.integers(0, 4, size=(N, 2)).astype(np.uint8)
        sp = spawn_tape()
        assert_same_step(g.step(act, spawn=sp), o.step(act, spawn=sp), "tick %d" % t)
        if t % 16 == 0:
            assert_same_state(g, o)
    assert_same_state(g, o)
    assert np.array_equal(g.stats, o.stats)


@pytest.mark.parametrize("N", [1, 127, 129, 5000])
def test_rng_mode_long_episodes_and_tails_bits10(L, N):
    g, o = make_pair(N, 10, 10, layout=L, obs_dtype=abi.BF16, seed=77, env_id_base=11, auto_reset=True)
    assert np.array_equal(g.reset(), o.reset())
    for t in range(40):
        assert_same_step(g.step(), o.step(), "tick %d" % t)
    assert_same_state(g, o)


def test_long_games_fill_the_board_bits10(L):
    """wall-avoiding tape from the reference fixtures -> long trails crossing the 64-bit word boundary"""
    tr = load_npz("traj.npz")
    length = tr["length"]
    long_ones = np.argsort(length)[-64:]
    assert length[long_ones].max() >= 20
    starts = np.concatenate([[0], np.cumsum(length + 1)[:-1]])
    G = len(long_ones); T = int(length[long_ones].max())
    tape = np.zeros((T, G, 2), np.uint8)
    for i, gi in enumerate(long_ones):
        tape[:length[gi], i] = tr["actions"][starts[gi] + 1: starts[gi] + 1 + length[gi]]
    g, o = make_pair(G, 10, 10, layout=L, obs_dtype=abi.I8, auto_reset=False)
    g.reset(spawn=tr["spawn"][long_ones]); o.reset(spawn=tr["spawn"][long_ones])
    for t in range(T):
        assert_same_step(g.step(tape[t]), o.step(tape[t]), "tick %d" % t)
        assert_same_state(g, o, "tick %d" % t)


def test_step_many_frozen_bad_actions_masked_reset_bits10(L):
    N, T = 3000, 24
    rng = np.random.default_rng(3)
    g, o = make_pair(N, 10, 10, layout=L, obs_dtype=abi.BF16, seed=11)
    g.reset(); o.reset()
    act = rng.integers(0, 4, size=(T, N, 2)).astype(np.uint8)
    assert_same_step(g.step_many(T, actions=act), o.step_many(T, actions=act))
    assert_same_step(g.step_many(T, obs_every_tick=False), o.step_many(T, obs_every_tick=False))
    assert_same_state(g, o)
    g, o = make_pair(N, 10, 10, layout=L, obs_dtype=abi.I8, auto_reset=False, seed=2)
    g.reset(); o.reset()
    for t in range(12):
        a = rng.integers(0, 4, size=(N, 2)).astype(np.int64)
        if t == 3:
            a[::7, 0] = 9; a[1::7, 1] = -1
        assert_same_step(g.step(a), o.step(a), "tick %d" % t)
    assert np.array_equal(g.stats, o.stats)
    mask = (np.arange(N) % 3 == 0).astype(np.uint8)
    assert np.array_equal(g.reset(mask=mask), o.reset(mask=mask))
    assert_same_state(g, o)


def test_export_import_round_trip_and_cross_layout(L):
    N = 2000
    a = GpuEnvNumpy(N, 10, 10, obs_dtype=abi.I8, seed=5, layout=L)
    b = GpuEnvNumpy(N, 10, 10, obs_dtype=abi.I8, seed=5, layout="tile8")
    a.reset(); b.reset()
    for _ in range(9):
        ra, rb = a.step(), b.step()
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
    ex = a.export()
    for k, v in b.export().items():
        assert np.array_equal(ex[k], v), k
    c = GpuEnvNumpy(N, 10, 10, obs_dtype=abi.I8, seed=5, layout=L)  # tile8 export -> bits10 import continues identically
    c.import_(**b.export()); c.env.counter = b.env.counter
    for _ in range(9):
        rb, rc = b.step(), c.step()
        for x, y in zip(rb, rc):
            assert np.array_equal(x, y)


def test_sharding_and_host_env_bits10(L):
    from tron_b200.batch_env import HostTron
    N = 2048
    full = GpuEnvNumpy(N, 10, 10, obs_dtype=abi.I8, seed=21, layout=L)
    lo = GpuEnvNumpy(N // 2, 10, 10, obs_dtype=abi.I8, seed=21, env_id_base=0, layout=L)
    hi = GpuEnvNumpy(N // 2, 10, 10, obs_dtype=abi.I8, seed=21, env_id_base=N // 2, layout=L)
    assert np.array_equal(full.reset(), np.concatenate([lo.reset(), hi.reset()]))
    for t in range(10):
        f = full.step(); l = lo.step(); h = hi.step()
        for x, y, z in zip(f, l, h):
            assert np.array_equal(x, np.concatenate([y, z]))
    N = 10000
    h = HostTron(N, 10, 10, obs_dtype=abi.BF16, n_chunks=7, seed=31, layout=L)
    o = oc.OracleEnv(N, 10, 10, obs_dtype=abi.BF16, seed=31)
    assert np.array_equal(h.reset(), o.reset())
    rng = np.random.default_rng(9)
    for t in range(8):
        act = rng.integers(0, 4, size=(N, 2)).astype(np.uint8)
        obs, rew, done, winner = h.step(act)
        wo, wr, wd, ww, _ = o.step(act)
        assert np.array_equal(obs, wo) and np.array_equal(rew, wr) and np.array_equal(done, wd) and np.array_equal(winner, ww)
    h.close()


def test_bit_layouts_refuse_what_they_cannot_represent():
    from tron_b200 import _lib
    from tron_b200.batch_env import BatchedTron
    with pytest.raises(_lib.TronError):
        BatchedTron(16, 12, 12, layout="bits10")
    with pytest.raises(_lib.TronError):
        BatchedTron(16, 12, 11, layout="bits")  # 132 cells > 128 bits
    env = BatchedTron(16, 10, 10, layout="bits10", slide_mode="ice")
    env.reset()
    with pytest.raises(_lib.TronError):
        env.step()


# ------------------------------------------------------------------ TRON_LAYOUT_BITS only: slide modes and other boards
def test_slide_tape_fixture_bits():
    """reference-generated ice-mode trajectories (explicit Bernoulli tape) on the three-plane layout"""
    sl = load_npz("slide.npz")
    W = int(sl["W"]); length = sl["length"]; G = len(length)
    tape, stape, row = ragged_to_tapes(length, sl["actions"], sl["slide"])
    env = GpuEnvNumpy(G, W, W, obs_dtype=abi.I8, auto_reset=False, slide_mode=abi.SLIDE_TAPE, layout="bits")
    env.reset(spawn=sl["spawn"])
    for t in range(tape.shape[0]):
        obs, rew, done, winner, _ = env.step(tape[t], slide_tape=stape[t])
        live = row[t + 1] >= 0
        r = row[t + 1][live]
        assert (env.export()["tiles"][live] == sl["tiles"][r]).all()
        assert (obs[live, 0, 0] == sl["obs1"][r]).all() and (obs[live, 1, 0] == sl["obs2"][r]).all()
        assert (done[live] == sl["done"][r]).all() and (winner[live] == sl["winner"][r]).all()


@pytest.mark.parametrize("mode,rate", [(abi.SLIDE_ICE, 0.15), (abi.SLIDE_ICE, 0.6), (abi.SLIDE_TEMPER, 0.0)])
@pytest.mark.parametrize("dt,enc", [(abi.I8, abi.ENC_LUT1), (abi.BF16, abi.ENC_POPUP3), (abi.F32, abi.ENC_POPUP3_CONST), (abi.BF16, abi.ENC_NONE)])
def test_slide_rng_modes_bits(mode, rate, dt, enc):
    N = 3000
    g, o = make_pair(N, 10, 10, layout="bits", obs_dtype=dt, obs_enc=enc, seed=99, slide_mode=mode, slide_rate=rate, const_plane=5.0)
    a, b = g.reset(), o.reset()
    if enc != abi.ENC_NONE:
        assert np.array_equal(a, b)
    slid = False
    for t in range(40):
        assert_same_step(g.step(), o.step(), "tick %d" % t)
        slid |= bool((o.export()["tiles"] >= 5).any())
    assert slid
    assert_same_state(g, o)
    if mode == abi.SLIDE_TEMPER:
        assert np.array_equal(g.env.slide_params.cpu().numpy(), o.slide_params)


@pytest.mark.parametrize("W,H,N", [(3, 3, 1000), (5, 5, 777), (7, 7, 4096), (11, 11, 130), (6, 11, 500), (8, 16, 2049), (2, 64, 300), (16, 8, 129)])
@pytest.mark.parametrize("dt", [abi.BF16, abi.F32, abi.I8])
def test_other_boards_bits(W, H, N, dt):
    g, o = make_pair(N, W, H, layout="bits", obs_dtype=dt, seed=W * 100 + H, slide_mode=abi.SLIDE_ICE, slide_rate=0.3)
    assert np.array_equal(g.reset(), o.reset())
    for t in range(24):
        assert_same_step(g.step(), o.step(), "tick %d" % t)
    assert_same_step(g.step_many(6), o.step_many(6))
    assert_same_state(g, o)
    assert np.array_equal(g.stats, o.stats)
