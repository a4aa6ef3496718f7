"""Drop-in for the reference's tron/minimax.py: MinimaxPlayer(depth, mode) -- the depth-2 minimax / Voronoi opponent.

The search itself runs in the CUDA library (tron_minimax_actions, csrc/minimax.cu): one launch evaluates the 16 leaves of
this game (or of every game of a batch, BatchedTron.minimax_actions).  Python's global `random` is consumed exactly where the
reference consumes it -- one random.choice / random.randint per finished depth-1 node, then one at the root
(tron/minimax.py:233-234,266-267; depth-2 searches never prune because the root's minimax action is still 0 while its children
are searched) -- so seeded runs match, also when slide draws from the same RNG are interleaved (tron/game.py:158-198).
Only depth 2 with the Voronoi heuristic is implemented (the only configuration the reference instantiates:
tron/util.py:82-83 `MinimaxPlayer(2, "voronoi")`, tron/game.py).
"""
import random
from enum import Enum

import numpy as np
import torch

from tron.player import Direction, Player
from tron import _gpu

_UNEXPANDED = -(2 ** 31)


class Mode(Enum):
    DISTWALL = 1
    VORNOI = 2


class MinimaxPlayer(Player):
    def __init__(self, depth, mode=Mode.VORNOI):
        super(MinimaxPlayer, self).__init__()
        if depth != 2 or mode == Mode.DISTWALL:
            raise NotImplementedError("only MinimaxPlayer(2, voronoi) is implemented (the configuration the reference uses)")
        self.mode = mode
        self.depth = depth
        self.direction = None

    def _search(self, map, id):
        env = _gpu.env_for(map.width, map.height)
        env.import_(tiles=torch.as_tensor(np.ascontiguousarray(map._codes, np.int8)).view(1, map.width + 2, map.height + 2))
        _, vals, ties = env.minimax_actions(id, tie_mode=0, counter=0, want_values=True, want_ties=True)
        return [None if v == _UNEXPANDED else int(v) for v in vals[0].tolist()], ties[0].tolist()

    def root_values(self, map, id):
        """minimax value of each of the 4 moves for player `id` (None = move not expanded), evaluated on the GPU"""
        return self._search(map, id)[0]

    def action(self, map, id):
        vals, ties = self._search(map, id)
        for t in ties:  # the depth-1 nodes finish in move order; each draws once (the chosen enemy move itself is never used)
            if t == 0:
                random.randint(1, 4)                                # enemy boxed in (minimax.py:233-234)
            elif t > 0:
                random.choice(range(t))                             # random.choice(minimax_acts) (minimax.py:266-267)
        expanded = [v for v in vals if v is not None]
        if not expanded:
            next_action = random.randint(1, 4)                      # minimax.py:233-234
        else:
            best = max(expanded)
            next_action = random.choice([i + 1 for i, v in enumerate(vals) if v == best])  # minimax.py:266-267
        return Direction(next_action)

    def next_position_and_direction(self, current_position, id, map, action=None):
        direction = action if action is not None else self.action(map, id)
        return self.next_position(current_position, direction), direction

    def next_position(self, current_position, direction):
        d = {Direction.UP: (-1, 0), Direction.RIGHT: (0, 1), Direction.DOWN: (1, 0), Direction.LEFT: (0, -1)}[direction]
        return current_position[0] + d[0], current_position[1] + d[1]
