"""Drop-in for the reference's tron/player.py: Direction (player.py:4-8), Player (:11-42), ACPlayer (:95-132).
KeyboardPlayer / pygame handling is UI and out of scope; Mode is kept so imports resolve."""
from enum import Enum


class Direction(Enum):
    UP = 1
    RIGHT = 2
    DOWN = 3
    LEFT = 4


class Player(object):
    def __init__(self):
        pass

    def find_file(self, name):
        pass

    def next_position(self, current_position, direction):
        pass

    def get_direction(self, current_position, direction):
        pass

    def next_position_and_direction(self, current_position, action):
        pass

    def action(self, map, id):
        pass

    def step(self, state, action, reward, next_step, done):
        pass

    def learn(self, experiences, gamma):
        pass

    def soft_update(self, local_model, target_model, tau):
        pass

    def manage_event(self, event):
        pass


class Mode(Enum):
    ARROWS = 1
    ZQSD = 2


_DELTA = {Direction.UP: (-1, 0), Direction.RIGHT: (0, 1), Direction.DOWN: (1, 0), Direction.LEFT: (0, -1)}


class ACPlayer(Player):
    """Player driven by integer actions 0..3 (UP, RIGHT, DOWN, LEFT)."""

    def get_direction(self, next_action):
        if next_action not in (0, 1, 2, 3):
            raise UnboundLocalError("action must be 0..3, got %r" % (next_action,))  # the reference's failure mode (player.py:107-118)
        return Direction(int(next_action) + 1)

    def next_position_and_direction(self, current_position, action):
        direction = self.get_direction(action)
        return self.next_position(current_position, direction), direction

    def next_position(self, current_position, direction):
        d = _DELTA[direction]
        return current_position[0] + d[0], current_position[1] + d[1]
