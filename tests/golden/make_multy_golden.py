#!/usr/bin/env python
"""Generate tests/golden/multy.json by running the reference (authoring container only): the per-game side features
Game.get_multy / get_rate (tron/game.py:96-102,137-139) for seeded games, and the `extra` arguments Game.main_loop hands to
model.act / model2.act (tron/game.py:296-304), captured with stub models.

Usage:  python tests/golden/make_multy_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, new_game  # noqa: E402


def main():
    G, U, P = import_reference()
    import torch

    class Stub:
        def __init__(self):
            self.extras = []

        def act(self, x, extra=None):
            self.extras.append([float(v) for v in torch.as_tensor(extra).flatten().tolist()])
            return 1  # both run right until somebody hits the wall

    games = []
    for seed in range(24):
        random.seed(seed)
        g = new_game(G, P, 10, (2, 1, 6, 3), "temper")
        rec = dict(seed=seed, weight=list(g.weight), degree=g.degree, multy0=g.get_multy(0), multy1=g.get_multy(1), rate=g.get_rate(),
                   rate0=g.get_rate(0), rate1=g.get_rate(1))
        m1, m2 = Stub(), Stub()
        random.seed(1000 + seed)
        g.main_loop(m1, pop=U.pop_up, model2=m2)
        rec.update(act_extra_model=m1.extras[0], act_extra_model2=m2.extras[0], ticks=len(m1.extras))
        games.append(rec)
    with open(os.path.join(HERE, "multy.json"), "w") as f:
        json.dump(dict(games=games), f)
    print("wrote multy.json:", games[0])


if __name__ == "__main__":
    main()


def scripted_games():
    """tests/golden/scripted.json: make_game(True, False, gamemode="ice") -- P1 driven by an action tape, P2 the reference's scripted
    MinimaxPlayer, slide mode on -- under a fixed global seed.  The scripted player's random tie-breaks and the slide draws come
    from the same global RNG, interleaved per player (tron/game.py:158-198), so reproducing these trajectories needs the same order."""
    import hashlib
    import numpy as np
    from make_golden import tiles_of
    G, U, P = import_reference()
    games = []
    for seed in range(10):
        random.seed(seed)
        g = U.make_game(True, False, gamemode="ice", slide_pram=0.4)
        rng = np.random.default_rng(seed)
        spawn = [list(p.position) for p in g.pps]
        ticks, acts = [], []
        D = ((-1, 0), (0, 1), (1, 0), (0, -1))
        while not g.done and len(ticks) < 60:
            t = tiles_of(g.map()); r, c = g.pps[0].position
            free = [k for k in range(4) if t[r + D[k][0] + 1, c + D[k][1] + 1] == 0]  # P1 avoids walls and trails so the game lasts
            a1 = int(free[int(rng.integers(0, len(free)))]) if free else int(rng.integers(0, 4))
            g.step(a1, 0)
            acts.append(a1)
            ticks.append(dict(pos=[list(map(int, p.position)) for p in g.pps], alive=[bool(p.alive) for p in g.pps], done=bool(g.done), winner=g.winner or 0,
                              tiles_sha1=hashlib.sha1(tiles_of(g.map()).tobytes()).hexdigest()))
        games.append(dict(seed=seed, spawn=spawn, actions=acts, ticks=ticks))
    with open(os.path.join(HERE, "scripted.json"), "w") as f:
        json.dump(dict(games=games), f)
    print("wrote scripted.json:", [len(x["ticks"]) for x in games])


if __name__ == "__main__" and "--scripted" in sys.argv:
    scripted_games()
