"""Differential fuzz on the CPU: the C oracle against the pure-Python port (itself validated cell-for-cell against the live
reference) on grid sizes, episode lengths and slide tapes beyond the committed fixtures."""
import numpy as np
import pytest

from oracle import c_oracle as oc
from oracle import py_port as pp
from tron_b200 import abi


def _tiles(board):
    return np.array([[t.value for t in row] for row in board.cells], np.int8)


@pytest.mark.parametrize("seed", range(6))
def test_c_oracle_equals_python_port(seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(25):
        W = int(rng.integers(2, 21))
        slide = bool(rng.integers(0, 2))
        while True:
            sp = [int(v) for v in rng.integers(0, W, size=4)]
            if sp[:2] != sp[2:]:
                break
        tape = {"cur": (0, 0)}
        game = pp.PyGame(W, W, sp[:2], sp[2:], mode="ice" if slide else None, bernoulli=lambda i: bool(tape["cur"][i]))
        env = oc.OracleEnv(1, W, W, obs_dtype=abi.I8, auto_reset=False, slide_mode=abi.SLIDE_TAPE if slide else abi.SLIDE_NONE)
        obs = env.reset(spawn=np.array([sp], np.int8))
        assert (obs[0, 0, 0] == game.board().view_for(1)).all()
        eps = float(rng.choice([1.0, 0.3, 0.0]))
        while not game.done:
            t = _tiles(game.history[-1].board)
            acts = []
            for i in (0, 1):
                p = game.pos[i]
                free = [a for a, (dr, dc) in enumerate([(-1, 0), (0, 1), (1, 0), (0, -1)]) if t[p[0] + 1 + dr, p[1] + 1 + dc] == 0]
                acts.append(int(free[rng.integers(0, len(free))]) if (free and rng.random() >= eps) else int(rng.integers(0, 4)))
            tape["cur"] = (int(rng.random() < 0.4), int(rng.random() < 0.4))
            o1, o2, done = game.step(*acts)
            obs, rew, dn, wn, _ = env.step(np.array([acts], np.uint8), slide_tape=np.array([tape["cur"]], np.uint8))
            ex = env.export()
            assert (ex["tiles"][0] == _tiles(game.history[-1].board)).all()
            assert (obs[0, 0, 0] == o1).all() and (obs[0, 1, 0] == o2).all()
            assert bool(dn[0]) == done and int(wn[0]) == (game.winner or 0)
            assert ex["heads"][0].tolist() == list(game.pos[0]) + list(game.pos[1]) and ex["alive"][0].tolist() == [int(x) for x in game.alive]
