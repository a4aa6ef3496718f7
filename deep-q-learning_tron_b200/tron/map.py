"""Drop-in for the reference's tron/map.py: Tile (map.py:9-17) and Map (map.py:43-92).

Storage is an int8 array of Tile.value codes (what the CUDA kernels use); `_data` exposes the reference's
object-dtype view of Tile members.  state_for_player is evaluated by the GPU (tron_observe), or returned from
the observation the fused step already produced when the map came out of Game.history.
"""
from enum import Enum

import numpy as np


class Tile(Enum):
    EMPTY = 0
    WALL = -1
    PLAYER_ONE_BODY = 1
    PLAYER_ONE_HEAD = 2
    PLAYER_TWO_BODY = 3
    PLAYER_TWO_HEAD = 4
    PLAYER_ONE_slide = 5
    PLAYER_TWO_slide = 6

    def color(self):  # RGB used by the pygame window (map.py:21-41)
        return {Tile.EMPTY: (0, 0, 0), Tile.WALL: (255, 255, 255), Tile.PLAYER_ONE_BODY: (0, 17, 128),
                Tile.PLAYER_ONE_HEAD: (0, 34, 255), Tile.PLAYER_ONE_slide: (0, 180, 250), Tile.PLAYER_TWO_BODY: (128, 17, 0),
                Tile.PLAYER_TWO_HEAD: (255, 34, 0), Tile.PLAYER_TWO_slide: (250, 100, 0)}.get(self)


_TILE_BY_CODE = np.empty(8, dtype=object)
for _t in Tile:
    _TILE_BY_CODE[_t.value + 1] = _t


def _code(x):
    return x.value if isinstance(x, Tile) else int(x)


def is_on_border(i, j, w, h):
    return i == 0 or i == w - 1 or j == 0 or j == h - 1


class Map:
    def __init__(self, w, h, empty, wall):
        self.width = w
        self.height = h
        codes = np.full((w + 2, h + 2), _code(empty), dtype=np.int8)
        codes[0, :] = codes[-1, :] = codes[:, 0] = codes[:, -1] = _code(wall)
        self._codes = codes
        self._obs_cache = None  # (obs_p1, obs_p2) produced by the fused CUDA step for exactly these codes

    @classmethod
    def _from_codes(cls, w, h, codes, obs=None):
        m = cls.__new__(cls)
        m.width, m.height = w, h
        m._codes = np.array(codes, dtype=np.int8).reshape(w + 2, h + 2)
        m._obs_cache = obs
        return m

    @property
    def _data(self):
        return _TILE_BY_CODE[self._codes.astype(np.int64) + 1]

    def clone(self):
        return Map._from_codes(self.width, self.height, self._codes.copy(), self._obs_cache)

    def apply(self, converter):
        out = Map._from_codes(self.width, self.height, self._codes.copy())
        conv = np.array([[converter(t) for t in row] for row in self._data])
        out._converted = conv
        return out

    def array(self):
        return getattr(self, "_converted", self._data)

    def clone_array(self):
        return self.clone()._data

    def color(self, t, p):  # map.py:67-81, kept for callers that colour single tiles
        if t == Tile.EMPTY:
            return 1
        if t == Tile.WALL:
            return -1
        if t in (Tile.PLAYER_ONE_BODY, Tile.PLAYER_ONE_slide):
            return -2 if p == 1 else -3
        if t == Tile.PLAYER_ONE_HEAD:
            return 10 if p == 1 else -10
        if t in (Tile.PLAYER_TWO_BODY, Tile.PLAYER_TWO_slide):
            return -3 if p == 1 else -2
        if t == Tile.PLAYER_TWO_HEAD:
            return 10 if p == 2 else -10
        return None

    def state_for_player(self, p):
        """(W+2, H+2) int64 observation for player p (map.py:83-84), computed on the GPU."""
        if self._obs_cache is None:
            from ._gpu import observe_codes
            self._obs_cache = observe_codes(self._codes, self.width, self.height)
        return self._obs_cache[0 if p == 1 else 1].copy()

    def __getitem__(self, index):
        (i, j) = index
        return _TILE_BY_CODE[int(self._codes[i + 1][j + 1]) + 1]

    def __setitem__(self, position, other):
        (i, j) = position
        self._codes[i + 1][j + 1] = _code(other)
        self._obs_cache = None
