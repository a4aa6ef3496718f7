"""install(): make `import config`, `from tron.game import Game`, `import DQN`, `import DDQN` resolve to this package's
drop-in mirrors of the reference's modules (same names as in Deep-Q-learning_TRON/)."""
import os
import sys


def install():
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (root, here):
        if p not in sys.path:
            sys.path.insert(0, p)
    return here
