"""Drop-in for the reference's config.py (config.py:1-41): `from config import *` yields the same names with the same
values.  MAP_WIDTH / MAP_HEIGHT are the grid-size source of the drop-in Game / make_game, exactly as in the reference."""
import torch

_SETTINGS = {
    # where the networks live (config.py:3)
    "device": "cuda" if torch.cuda.is_available() else "cpu",
    # DDQN hyper-parameters (config.py:5-7)
    "GAMMA": 0.9, "BATCH_SIZE": 64,
    # A2C / ACKTR optimiser + loss constants (config.py:10-21); unused by the DQN path, kept so imports resolve
    "lr": 3e-3, "eps": 1e-5, "alpha": 0.99, "NUM_PROCESSES": 16, "NUM_ADVANCED_STEP": 5,
    "value_loss_coef": 0.5, "entropy_coef": 0.01, "policy_loss_coef": 1, "max_grad_norm": 0.5,
    # board size (config.py:23-24)
    "MAP_WIDTH": 10, "MAP_HEIGHT": 10,
    # logging / evaluation cadence (config.py:26-28)
    "SHOW_ITER": 20, "PLAY_WITH_MINIMAX": 200,
    # stochastic-slide game modes (config.py:32-34)
    "slide": 0.15, "GAME_MODE": "temper",
    # terminal reward constants [win, lose] (config.py:37-41)
    "reward_cons1": [10, -10], "reward_cons2": [10, -20], "reward_cons3": [20.0, -10.0],
}
globals().update(_SETTINGS)
__all__ = sorted(_SETTINGS)
