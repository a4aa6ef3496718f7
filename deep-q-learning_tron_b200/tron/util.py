"""Drop-in for the reference's tron/util.py: pop_up (:11-37), prob_map (:38-45), make_game (:46-84), get_reward (:87-94)."""
import random

import numpy as np

from config import *  # noqa: F401,F403
import config as _config
from tron.game import *  # noqa: F401,F403
from tron.game import Game, PositionPlayer
from tron.player import ACPlayer
from tron import _gpu


def pop_up(map):
    """(R,C) observation -> (3,R,C) float64 planes [wall, my, enemy]; evaluated by tron_pop_up on the GPU."""
    return _gpu.pop_up_gpu(np.asarray(map))


def prob_map(x):
    return np.full((_config.MAP_WIDTH + 2, _config.MAP_HEIGHT + 2), float(x))


def make_game(p1, p2, mode=None, gamemode=None, slide_pram=None):
    W, H = _config.MAP_WIDTH, _config.MAP_HEIGHT
    if mode == "fair":
        point_y = random.randint(0, H - 1)
        point_x = random.randint(0, W - 1)
        lo1x, hi1x = max(0, point_x - 1), min(W - 1, point_x + 1)
        lo1y, hi1y = max(0, point_y - 1), min(H - 1, point_y + 1)
        lo2x, hi2x = W - 1 - hi1x, W - 1 - lo1x
        lo2y, hi2y = H - 1 - hi1y, H - 1 - lo1y
    else:
        lo1x = lo1y = lo2x = lo2y = 0
        hi1x = hi2x = W - 1
        hi1y = hi2y = H - 1
    x1 = random.randint(lo1x, hi1x)
    y1 = random.randint(lo1y, hi1y)
    x2 = random.randint(lo2x, hi2x)
    y2 = random.randint(lo2y, hi2y)
    while x1 == x2 and y1 == y2:
        x1 = random.randint(lo1x, hi1x)
        y1 = random.randint(lo1y, hi1y)
    from tron.minimax import MinimaxPlayer
    return Game(W, H, [PositionPlayer(1, ACPlayer() if p1 else MinimaxPlayer(2, "voronoi"), [x1, y1]),
                       PositionPlayer(2, ACPlayer() if p2 else MinimaxPlayer(2, "voronoi"), [x2, y2])], gamemode, slide_pram)


def get_reward(game, constants):
    if game.winner is None:
        return 0, 0
    if game.winner == 1:
        return constants[0], constants[1]
    return constants[1], constants[0]
