"""Drop-in for the reference's config.py (config.py:1-41): same names, same values.  MAP_WIDTH / MAP_HEIGHT are
the grid-size source of the drop-in Game / make_game, exactly as in the reference."""
import torch

device = 'cuda' if torch.cuda.is_available() else 'cpu'

GAMMA = 0.9
BATCH_SIZE = 64

lr = 3e-3
eps = 1e-5
alpha = 0.99

NUM_PROCESSES = 16
NUM_ADVANCED_STEP = 5

value_loss_coef = 0.5
entropy_coef = 0.01
policy_loss_coef = 1
max_grad_norm = 0.5

MAP_WIDTH = 10
MAP_HEIGHT = 10

SHOW_ITER = 20
PLAY_WITH_MINIMAX = 200

slide = 0.15
GAME_MODE = "temper"

reward_cons1 = [10, -10]
reward_cons2 = [10, -20]
reward_cons3 = [20.0, -10.0]
