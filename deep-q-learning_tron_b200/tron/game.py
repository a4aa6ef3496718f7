"""Drop-in for the reference's tron/game.py: PositionPlayer (:36-58), HistoryElement (:61-65), Game (:70-328).

Game keeps the reference's attributes (width, height, pps, history, winner, done, next_p1, next_p2, mode, slide,
weight, degree) and methods (map, step, next_frame, main_loop, get_rate, prob_map, ...), but every tick runs in the
CUDA library: Game owns a 1-env BatchedTron; step() launches the fused tick kernel and copies the new grid, both
observations and the flags back to maintain `history`, `pps[i].position/alive` and `winner`.
Deviations (documented in DESIGN.md): stepping a finished game is a no-op (the reference keeps moving dead heads
and wraps negative indices); main_loop(pop=None) passes the raw 1-plane observation instead of crashing.
"""
import random
from enum import Enum
from time import sleep

import numpy as np
import torch

from config import *  # noqa: F401,F403  (MAP_WIDTH, MAP_HEIGHT, slide, device ... like the reference)
import config as _config
from tron.map import Map, Tile
from tron.player import ACPlayer, Direction
from tron import _gpu


class Winner(Enum):
    PLAYER_ONE = 1
    PLAYER_TWO = 2


class PositionPlayer:
    def __init__(self, id, player, position):
        self.id = id
        self.player = player
        self.position = position
        self.alive = True

    def body(self):
        return Tile.PLAYER_ONE_BODY if self.id == 1 else Tile.PLAYER_TWO_BODY if self.id == 2 else None

    def slide(self):
        return Tile.PLAYER_ONE_slide if self.id == 1 else Tile.PLAYER_TWO_slide if self.id == 2 else None

    def head(self):
        return Tile.PLAYER_ONE_HEAD if self.id == 1 else Tile.PLAYER_TWO_HEAD if self.id == 2 else None


class HistoryElement:
    def __init__(self, mmap, player_one_direction, player_two_direction):
        self.map = mmap
        self.player_one_direction = player_one_direction
        self.player_two_direction = player_two_direction


_DELTA = ((-1, 0), (0, 1), (1, 0), (0, -1))


def _takes_extra(act):
    """does act(obs, extra) accept the side-feature argument?  (decided from the signature, never by catching TypeError)"""
    import inspect
    try:
        params = list(inspect.signature(act).parameters.values())
    except (TypeError, ValueError):
        return True
    if any(p.kind == p.VAR_POSITIONAL for p in params):
        return True
    return len([p for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]) >= 2


class Game:
    def __init__(self, width, height, pps, mode=None, slide_pram=None):
        self.width = width
        self.height = height
        self.pps = pps
        self.winner = None
        self.next_p1 = []
        self.next_p2 = []
        self.weight = [random.randint(40, 101), random.randint(40, 101)]  # same global-RNG draws as game.py:83,87
        self.done = False
        self.mode = mode
        self.degree = random.randint(-30, 30)
        self.slide = _config.slide if slide_pram is None else slide_pram
        self._env = _gpu.new_env(width, height, slide_mode="tape" if mode in ("ice", "temper") else None)
        spawn = np.array([[pps[0].position[0], pps[0].position[1], pps[1].position[0], pps[1].position[1]]], np.int8)
        self._finished = False
        self.history = [HistoryElement(self._map_from_device(self._env.reset(spawn)), None, None)]

    def __del__(self):
        env = getattr(self, "_env", None)
        if env is not None:
            try:
                env.release()  # the 1-env GPU context goes back to the pool for the next Game of this size
            except Exception:
                pass

    # ------------------------------------------------------------------ device <-> history
    def _map_from_device(self, snap):
        self._heads, self._alive = snap["heads"], snap["alive"]
        return Map._from_codes(self.width, self.height, snap["tiles"], (snap["obs1"], snap["obs2"]))

    def map(self):
        return self.history[-1].map.clone()

    # ------------------------------------------------------------------ ice / temper helpers (game.py:96-147)
    def get_rate(self, player_num=None):
        if player_num is None:
            return -((self.degree - 30) * 0.6) / 100
        return (-((self.degree - 30) * 0.6) / 100) - ((70 - self.get_weight(player_num)) / 100)

    def get_degree(self):
        return float(self.degree)

    def get_degree_silde(self):
        return float((-self.slide * 100) * (10 / 6) + 30)

    def change_degree(self):
        if random.random() > 0.5:
            self.degree = min(30, self.degree + random.randint(0, 3))
        else:
            self.degree = max(-30, self.degree - random.randint(1, 5))

    def prob_map(self):
        return np.full((_config.MAP_WIDTH + 2, _config.MAP_HEIGHT + 2), self.get_degree_silde())

    def get_weight(self, player_num):
        return self.weight[player_num]

    def get_multy(self, player_num):
        return [self.get_degree(), self.get_weight(player_num)]

    def degree_map(self):
        return np.full((_config.MAP_WIDTH + 2, _config.MAP_HEIGHT + 2), self.get_degree())

    def _slide_draw(self, codes, i, action):
        """Consume the global RNG exactly like game.py:163-178 for player i: one random.random() iff its first move lands on a
        free in-bounds cell of the scratch grid (P2 sees P1's slide tile).  Only the Bernoulli outcome goes to the GPU."""
        pp = self.pps[i]
        r, c = pp.position[0] + _DELTA[action][0], pp.position[1] + _DELTA[action][1]
        if 0 <= r < self.width and 0 <= c < self.height and codes[r + 1, c + 1] == 0:
            rate = self.slide if self.mode == "ice" else self.get_rate(i)
            if random.random() <= rate:
                codes[r + 1, c + 1] = 5 if i == 0 else 6
                return 1
        return 0

    # ------------------------------------------------------------------ ticks
    def next_frame(self, action_p1, action_p2, window=None):
        actions = [action_p1, action_p2]
        sliding = self.mode in ("ice", "temper")
        tape = [0, 0]
        codes = None
        if sliding and not self._finished:  # scratch grid of game.py:151-156: both old heads are bodies before anyone moves
            codes = self.history[-1].map._codes.copy()
            for i, pp in enumerate(self.pps):
                codes[pp.position[0] + 1, pp.position[1] + 1] = 1 if i == 0 else 3
        # the reference interleaves per player: move decision (a scripted player may draw from the global RNG), then that
        # player's slide draw, P1 before P2 (game.py:158-198)
        for i, pp in enumerate(self.pps):
            if isinstance(pp.player, ACPlayer):
                actions[i] = int(actions[i])
                pp.player.direction = pp.player.get_direction(actions[i])
            elif hasattr(pp.player, "action"):  # scripted player (MinimaxPlayer): it looks at the current map (game.py:182)
                pp.player.direction = pp.player.action(self.map(), i + 1)
                actions[i] = pp.player.direction.value - 1
            else:
                raise NotImplementedError("unsupported player type %r" % type(pp.player).__name__)
            if codes is not None:
                tape[i] = self._slide_draw(codes, i, actions[i])
        if self._finished:
            return True  # finished game: frozen (documented deviation)
        snap = self._env.step(actions, slide_tape=tape if sliding else None)
        self.history[-1].player_one_direction = self.pps[0].player.direction
        self.history[-1].player_two_direction = self.pps[1].player.direction
        self.history.append(HistoryElement(self._map_from_device(snap), None, None))
        for i, pp in enumerate(self.pps):
            pp.position = (int(self._heads[2 * i]), int(self._heads[2 * i + 1]))
            pp.alive = bool(self._alive[i])
        self._last_done = self._finished = bool(snap["done"])
        self._last_winner = snap["winner"]
        self.next_p1 = self.history[-1].map.state_for_player(1)
        self.next_p2 = self.history[-1].map.state_for_player(2)
        if window:
            window.render_map(self.map())
        return True

    def _resolve(self):
        if getattr(self, "_last_done", False):
            if self._last_winner:
                self.winner = self._last_winner
            return True
        return False

    def step(self, action_p1, action_p2):
        if not self.next_frame(action_p1, action_p2):
            self.done = True
            return self.next_p1, self.next_p2, self.done
        if self._resolve():
            self.done = True
        return self.next_p1, self.next_p2, self.done

    def main_loop(self, model=None, pop=None, window=None, model2=None):
        if window:
            window.render_map(self.map())
        if model is None:  # players that decide for themselves (DQN.Ai, MinimaxPlayer): tick until the game is over
            while True:
                if window:
                    sleep(0.3)
                if not self.next_frame(None, None, window) or self._resolve():
                    self.done = self.done or self._finished
                    return
        if not model2:
            model2 = model
        dev = _config.device
        while True:
            if window:
                sleep(0.3)
            m = self.map()
            with torch.no_grad():
                acts = []
                for pid, mdl in ((1, model), (2, model2)):
                    o = m.state_for_player(pid)
                    x = torch.tensor(pop(o) if pop is not None else o[None]).float()
                    if getattr(mdl, "wants_prob_map", False):  # MapNet-style input: extra constant plane (game.py:297)
                        x = torch.cat([x, torch.tensor(self.prob_map()).unsqueeze(0).float()], 0)
                        a = mdl.act(x.unsqueeze(0))
                    else:  # game.py:299,304: model gets [[degree, weight_p1]], model2 gets [[rate]]
                        extra = torch.tensor([self.get_multy(0)] if pid == 1 else [[self.get_rate()]]).to(dev)
                        a = mdl.act(x.unsqueeze(0), extra) if _takes_extra(mdl.act) else mdl.act(x.unsqueeze(0))
                    acts.append(int(a))
            if not self.next_frame(acts[0], acts[1], window):
                break
            if self._resolve():
                self.done = True
                break
