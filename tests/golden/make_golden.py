#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE REFERENCE ITSELF.

Runs only in the authoring container (needs /root/reference, which never travels to the GPU box).
It imports the reference's own tron.game / tron.util / DQN / DDQN modules (read-only, nothing is
copied) behind a 7-line `orderedset` shim (the only missing third-party import; it is dead code on
this path) and writes small fixtures next to this file:

  kat.json        pinned edge cases (wall, head-on, swap, trail, reverse, history directions)
  digests.json    SHA-256 over observations of 2 x 1000 seeded random games (SURVEY 8c script)
  traj.npz        512 full 10x10 trajectories: spawn + action tape -> tiles/obs/done/winner per tick
  fuzz.json       300 games, W=H in [3,16], random + wall-avoiding policies, per-game SHA-256
  slide.npz       "ice" mode trajectories driven by an explicit Bernoulli tape
  popup.npz       pop_up planes for observations taken from traj.npz
  minimax.npz     MinimaxPlayer(2, voronoi) decisions + root child values on ~300 mid-game states, both players
  misc.json       get_reward table, get_rate table, replay container semantics, timing of the reference

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import random
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Deep-Q-learning_TRON"


def import_reference():
    shim = tempfile.mkdtemp(prefix="orderedset_shim_")
    with open(os.path.join(shim, "orderedset.py"), "w") as f:
        f.write("class OrderedSet(dict):\n"
                "    def add(self, x): self[x] = None\n"
                "    def remove(self, x): del self[x]\n"
                "    def __getitem__(self, i): return list(self.keys())[i]\n")
    for m in [k for k in sys.modules if k.split(".")[0] in ("tron", "config", "DQN", "DDQN", "Net")]:
        del sys.modules[m]  # never mix the reference's modules with the drop-in mirrors of the same names
    # the reference's tron/ has no __init__.py (namespace package): a regular package of the same name ANYWHERE on sys.path would
    # win, so the drop-in mirror's directory must not be importable while the reference is
    sys.path[:] = [p for p in sys.path if os.path.basename(os.path.normpath(p or ".")) != "deep-q-learning_tron_b200"]
    sys.path.insert(0, shim)
    sys.path.insert(0, REF)
    try:  # tron.game imports torchvision; import it BEFORE the file-less namespace module `tron` exists (torchvision's op registration
        import torchvision  # noqa: F401  walks sys.modules with inspect.getmodule, which raises on namespace packages)
    except Exception:
        pass
    import tron.game as game  # noqa
    import tron.util as util  # noqa
    import tron.player as player  # noqa
    return game, util, player


def tiles_of(m):
    """Map -> int8 (W+2,H+2) of Tile.value"""
    return np.array([[t.value for t in row] for row in m._data], dtype=np.int8)


def new_game(G, P, W, spawn, mode=None, slide=None):
    x1, y1, x2, y2 = (int(v) for v in spawn)
    return G.Game(W, W, [G.PositionPlayer(1, P.ACPlayer(), [x1, y1]), G.PositionPlayer(2, P.ACPlayer(), [x2, y2])],
                  mode, slide)


def snapshot(g):
    m = g.history[-1].map
    return dict(tiles=tiles_of(m).tolist(),
                obs1=np.asarray(m.state_for_player(1)).astype(int).tolist(),
                obs2=np.asarray(m.state_for_player(2)).astype(int).tolist(),
                alive=[bool(p.alive) for p in g.pps], pos=[list(map(int, p.position)) for p in g.pps],
                done=bool(g.done), winner=int(g.winner or 0), history_len=len(g.history))


def make_kat(G, P):
    cases = {
        "wall_up": ((0, 4, 5, 5), [(0, 1)]),
        "wall_down": ((9, 4, 5, 5), [(2, 1)]),
        "wall_left": ((4, 0, 5, 5), [(3, 1)]),
        "wall_right": ((4, 9, 5, 5), [(1, 3)]),
        "both_wall": ((0, 4, 9, 5), [(0, 2)]),
        "head_on_same_cell": ((3, 3, 3, 5), [(1, 3)]),
        "swap_adjacent": ((3, 3, 3, 4), [(1, 3)]),
        "p1_into_p2_old_head": ((3, 3, 3, 4), [(1, 1)]),
        "p2_into_p1_old_head": ((3, 3, 3, 4), [(3, 3)]),
        "reverse_into_own_trail": ((3, 3, 7, 7), [(1, 1), (3, 1)]),
        "two_quiet_ticks": ((2, 2, 7, 7), [(1, 2), (0, 3)]),
        "p2_wall_p1_survives": ((5, 5, 9, 9), [(0, 2)]),
        "p1_trail_kills_p2": ((3, 3, 4, 5), [(1, 3), (1, 0)]),
        "corner_double": ((0, 0, 9, 9), [(3, 1)]),
    }
    out = {}
    for name, (spawn, tape) in cases.items():
        g = new_game(G, P, 10, spawn)
        ticks = [snapshot(g)]
        for a1, a2 in tape:
            g.step(a1, a2)
            ticks.append(snapshot(g))
        dirs = [[h.player_one_direction.value if h.player_one_direction else 0,
                 h.player_two_direction.value if h.player_two_direction else 0] for h in g.history]
        out[name] = dict(spawn=list(spawn), actions=[list(t) for t in tape], ticks=ticks, history_dirs=dirs)
    return out


def digest(G, P, seed, ngames, W=10):
    rng = np.random.default_rng(seed)
    h = hashlib.sha256()
    steps = 0
    wins = [0, 0, 0]
    for _ in range(ngames):
        while True:
            x1, y1, x2, y2 = (int(v) for v in rng.integers(0, [W, W, W, W]))
            if (x1, y1) != (x2, y2):
                break
        g = new_game(G, P, W, (x1, y1, x2, y2))
        h.update(g.map().state_for_player(1).astype(np.int8).tobytes())
        done = False
        while not done:
            a1, a2 = (int(v) for v in rng.integers(0, 4, size=2))
            s1, s2, done = g.step(a1, a2)
            steps += 1
            h.update(s1.astype(np.int8).tobytes())
            h.update(s2.astype(np.int8).tobytes())
            h.update(bytes([int(done), g.winner or 0]))
        wins[g.winner or 0] += 1
    return h.hexdigest(), steps, wins


def free_neighbour_policy(rng, tiles, pos, eps):
    """with prob 1-eps a uniformly random free neighbour if any, else uniform (SURVEY 8d proxy)"""
    if rng.random() >= eps:
        free = [a for a, (dr, dc) in enumerate([(-1, 0), (0, 1), (1, 0), (0, -1)])
                if tiles[pos[0] + 1 + dr, pos[1] + 1 + dc] == 0]
        if free:
            return int(free[rng.integers(0, len(free))])
    return int(rng.integers(0, 4))


def make_traj(G, P, ngames=512, W=10, seed=7):
    rng = np.random.default_rng(seed)
    spawns, lens, acts, tiles, obs1, obs2, done, winner, alive, pos = [], [], [], [], [], [], [], [], [], []
    for gi in range(ngames):
        while True:
            sp = tuple(int(v) for v in rng.integers(0, W, size=4))
            if sp[:2] != sp[2:]:
                break
        eps = [1.0, 0.5, 0.1, 0.0][gi % 4]
        g = new_game(G, P, W, sp)
        spawns.append(sp)
        tiles.append(tiles_of(g.history[-1].map)); obs1.append(g.map().state_for_player(1).astype(np.int8))
        obs2.append(g.map().state_for_player(2).astype(np.int8)); done.append(0); winner.append(0)
        alive.append([1, 1]); pos.append([sp[0], sp[1], sp[2], sp[3]]); acts.append([255, 255])
        n = 0
        while not g.done:
            t = tiles_of(g.history[-1].map)
            a1 = free_neighbour_policy(rng, t, g.pps[0].position, eps)
            a2 = free_neighbour_policy(rng, t, g.pps[1].position, eps)
            s1, s2, d = g.step(a1, a2)
            n += 1
            acts.append([a1, a2]); tiles.append(tiles_of(g.history[-1].map))
            obs1.append(s1.astype(np.int8)); obs2.append(s2.astype(np.int8))
            done.append(int(d)); winner.append(int(g.winner or 0))
            alive.append([int(p.alive) for p in g.pps])
            pos.append([int(g.pps[0].position[0]), int(g.pps[0].position[1]), int(g.pps[1].position[0]), int(g.pps[1].position[1])])
        lens.append(n)
    return dict(W=np.int32(W), spawn=np.array(spawns, np.int8), length=np.array(lens, np.int32),
                actions=np.array(acts, np.uint8), tiles=np.array(tiles, np.int8), obs1=np.array(obs1, np.int8),
                obs2=np.array(obs2, np.int8), done=np.array(done, np.uint8), winner=np.array(winner, np.uint8),
                alive=np.array(alive, np.uint8), pos=np.array(pos, np.int8))


def make_fuzz(G, P, ngames=300, seed=11):
    rng = np.random.default_rng(seed)
    games = []
    for gi in range(ngames):
        W = int(rng.integers(3, 17))
        while True:
            sp = tuple(int(v) for v in rng.integers(0, W, size=4))
            if sp[:2] != sp[2:]:
                break
        eps = [1.0, 0.3, 0.05][gi % 3]
        g = new_game(G, P, W, sp)
        h = hashlib.sha256()
        tape = []
        while not g.done:
            t = tiles_of(g.history[-1].map)
            a1 = free_neighbour_policy(rng, t, g.pps[0].position, eps)
            a2 = free_neighbour_policy(rng, t, g.pps[1].position, eps)
            s1, s2, d = g.step(a1, a2)
            tape.append([a1, a2])
            h.update(tiles_of(g.history[-1].map).tobytes()); h.update(s1.astype(np.int8).tobytes())
            h.update(s2.astype(np.int8).tobytes())
            h.update(bytes([int(p.alive) for p in g.pps] + [int(d), g.winner or 0]))
            h.update(np.array([g.pps[0].position, g.pps[1].position], np.int8).tobytes())
        games.append(dict(W=W, spawn=list(sp), actions=tape, sha256=h.hexdigest(), winner=int(g.winner or 0)))
    return games


def make_slide(G, P, ngames=200, W=10, seed=13):
    """mode="ice": the Bernoulli draw `random.random() <= rate` (game.py:169) is driven by an explicit tape."""
    import tron.game as tg
    rng = np.random.default_rng(seed)
    cur = {"tape": (0, 0)}

    def fake_random():
        pid = sys._getframe(1).f_locals["id"]  # enumerate index of the player asking (game.py:158)
        return 0.0 if cur["tape"][pid] else 1.0
    real = tg.random.random
    spawns, lens, acts, slides, tiles, obs1, obs2, done, winner = [], [], [], [], [], [], [], [], []
    try:
        tg.random.random = fake_random
        for gi in range(ngames):
            while True:
                sp = tuple(int(v) for v in rng.integers(0, W, size=4))
                if sp[:2] != sp[2:]:
                    break
            g = new_game(G, P, W, sp, mode="ice", slide=0.15)
            spawns.append(sp); tiles.append(tiles_of(g.history[-1].map)); acts.append([255, 255]); slides.append([0, 0])
            obs1.append(g.map().state_for_player(1).astype(np.int8)); obs2.append(g.map().state_for_player(2).astype(np.int8))
            done.append(0); winner.append(0)
            n = 0
            while not g.done:
                t = tiles_of(g.history[-1].map)
                a1 = free_neighbour_policy(rng, t, g.pps[0].position, 0.2)
                a2 = free_neighbour_policy(rng, t, g.pps[1].position, 0.2)
                cur["tape"] = (int(rng.random() < 0.4), int(rng.random() < 0.4))
                s1, s2, d = g.step(a1, a2)
                n += 1
                acts.append([a1, a2]); slides.append(list(cur["tape"])); tiles.append(tiles_of(g.history[-1].map))
                obs1.append(s1.astype(np.int8)); obs2.append(s2.astype(np.int8)); done.append(int(d)); winner.append(int(g.winner or 0))
            lens.append(n)
    finally:
        tg.random.random = real
    return dict(W=np.int32(W), spawn=np.array(spawns, np.int8), length=np.array(lens, np.int32),
                actions=np.array(acts, np.uint8), slide=np.array(slides, np.uint8), tiles=np.array(tiles, np.int8),
                obs1=np.array(obs1, np.int8), obs2=np.array(obs2, np.int8), done=np.array(done, np.uint8),
                winner=np.array(winner, np.uint8))


def make_misc(G, U, P):
    out = {}

    class W_:  # get_reward only reads .winner
        def __init__(self, w): self.winner = w
    out["get_reward"] = {str(w): [list(map(float, U.get_reward(W_(w), c))) for c in ([10, -10], [10, -20], [20.0, -10.0])]
                         for w in (None, 1, 2)}
    g = new_game(G, P, 10, (1, 1, 5, 5), mode="temper")
    rates = []
    for degree in range(-30, 31, 5):
        for w in (40, 55, 70, 85, 101):
            g.degree = degree; g.weight = [w, 141 - w]
            rates.append([degree, w, 141 - w, g.get_rate(0), g.get_rate(1)])
    out["get_rate"] = rates
    out["degree_slide"] = [[s, G.Game(10, 10, [G.PositionPlayer(1, P.ACPlayer(), [0, 0]), G.PositionPlayer(2, P.ACPlayer(), [1, 1])], None, s).get_degree_silde()]
                           for s in (0.15, 0.0, 0.3)]
    # make_game spawn rule under a seeded global RNG (util.py:70-78): positions distinct, inside the grid
    random.seed(123)
    sp = []
    for _ in range(200):
        gm = U.make_game(True, True)
        sp.append([int(gm.pps[0].position[0]), int(gm.pps[0].position[1]), int(gm.pps[1].position[0]), int(gm.pps[1].position[1])])
    out["make_game_spawns_seed123"] = sp
    # make_game(mode="fair") box bounds (util.py:48-62): drive random.randint and record the (lo, hi) it is asked for
    import tron.util as tu
    fair = []
    real_randint = tu.random.randint
    try:
        for py in range(10):
            for px in range(10):
                calls = []

                def fake(a, b, _c=calls, _p=(py, px)):
                    _c.append((a, b))
                    return _p[len(_c) - 1] if len(_c) <= 2 else a + (len(_c) % 2) * (b - a)
                tu.random.randint = fake
                tu.make_game(True, True, mode="fair")
                fair.append([px, py] + [v for ab in calls[2:6] for v in ab])  # x1, y1, x2, y2 ranges
    finally:
        tu.random.randint = real_randint
    out["fair_bounds"] = fair
    # replay containers (plain python; importable even though Net() construction is broken)
    import DQN as RD
    import DDQN as RDD
    mem = RD.ReplayMemory(5)
    for i in range(8):
        mem.push(i, i * 10, i + 1, float(i), i % 3 == 0)
    out["ReplayMemory_cap5_push8"] = dict(memory=[list(t) for t in mem.memory], position=mem.position, length=len(mem))
    random.seed(5)
    out["ReplayMemory_sample3_distinct"] = len({t.old_state for t in mem.sample(3)}) == 3
    buf = RDD.ReplayBuffer(4, 5, 3)
    for i in range(8):
        buf.add(np.full((1, 3, 2, 2), i, np.float32), i % 4, float(-i), np.full((1, 3, 2, 2), i + 1, np.float32), i % 2 == 0)
    out["ReplayBuffer_cap5_add8_states"] = [int(e.state.flat[0]) for e in buf.memory]
    random.seed(6)
    s, a, r, s2, d = buf.sample()
    out["ReplayBuffer_sample"] = dict(shapes=[list(x.shape) for x in (s, a, r, s2, d)], dtypes=[str(x.dtype) for x in (s, a, r, s2, d)],
                                      distinct=len(set(s[:, 0, 0, 0].tolist())) == 3,
                                      consistent=bool(((s2[:, 0, 0, 0] - s[:, 0, 0, 0]) == 1).all() and (r[:, 0] == -s[:, 0, 0, 0]).all()))
    # timing of the reference loop in this container (context for the python port's speed; not a fixture that is asserted)
    random.seed(0)
    t0 = time.perf_counter(); steps = 0
    for _ in range(300):
        gm = U.make_game(True, True)
        d = False
        while not d:
            _, _, d = gm.step(random.randrange(4), random.randrange(4)); steps += 1
    dt = time.perf_counter() - t0
    out["reference_timing_10x10_random"] = dict(env_steps=steps, seconds=dt, steps_per_s_per_core=steps / dt)
    return out


def make_minimax(G, P, n_states=300, seed=17):
    """MinimaxPlayer(2, voronoi).action (tron/minimax.py:296-310) on mid-game states from traj.npz, both perspectives.
    random.choice -> first element, random.randint -> low bound, so ties are deterministic ("tie_mode 0")."""
    import tron.minimax as tm
    from tron.map import Map, Tile
    tr = np.load(os.path.join(HERE, "traj.npz"))
    live = np.nonzero(tr["done"] == 0)[0]
    rng = np.random.default_rng(seed)
    # prefer crowded boards: sort candidates by number of trail cells, take a spread
    crowd = (tr["tiles"][live] > 0).sum((1, 2))
    order = live[np.argsort(crowd)]
    pick = np.unique(np.concatenate([order[-n_states // 2:], rng.choice(order, n_states // 2, replace=False)]))
    real_choice, real_randint = tm.random.choice, tm.random.randint
    tiles_out, acts, vals = [], [], []
    try:
        tm.random.choice = lambda seq: seq[0]
        tm.random.randint = lambda a, b: a
        for row in pick:
            codes = tr["tiles"][row]
            m = Map(10, 10, Tile.EMPTY, Tile.WALL)
            m._data = np.array([[Tile(int(v)) for v in r] for r in codes])
            a, v = [], []
            for pid in (1, 2):
                pl = tm.MinimaxPlayer(2)
                d = pl.action(m, pid)
                a.append(d.value - 1)
                cv = [-(2 ** 31)] * 4
                for ch in pl.minimax.root._children:
                    cv[ch.get_action() - 1] = int(ch.get_value())
                v.append(cv)
            tiles_out.append(codes); acts.append(a); vals.append(v)
    finally:
        tm.random.choice, tm.random.randint = real_choice, real_randint
    return dict(tiles=np.array(tiles_out, np.int8), actions=np.array(acts, np.uint8), values=np.array(vals, np.int64))


def main():
    G, U, P = import_reference()
    kat = make_kat(G, P)
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"))
    dg = {str(s): dict(zip(("sha256", "steps", "wins"), digest(G, P, s, 1000))) for s in (0, 1)}
    json.dump(dg, open(os.path.join(HERE, "digests.json"), "w"), indent=1)
    print("digests", dg)
    np.savez_compressed(os.path.join(HERE, "traj.npz"), **make_traj(G, P))
    json.dump(make_fuzz(G, P), open(os.path.join(HERE, "fuzz.json"), "w"))
    np.savez_compressed(os.path.join(HERE, "slide.npz"), **make_slide(G, P))
    tr = np.load(os.path.join(HERE, "traj.npz"))
    sel = np.linspace(0, tr["obs1"].shape[0] - 1, 64).astype(int)
    np.savez_compressed(os.path.join(HERE, "popup.npz"), obs=tr["obs1"][sel],
                        planes=np.array([U.pop_up(o.astype(np.int64)) for o in tr["obs1"][sel]], np.float32))
    np.savez_compressed(os.path.join(HERE, "minimax.npz"), **make_minimax(G, P))
    misc = make_misc(G, U, P)
    json.dump(misc, open(os.path.join(HERE, "misc.json"), "w"), indent=1)
    print("reference timing", misc["reference_timing_10x10_random"])


if __name__ == "__main__":
    main()
