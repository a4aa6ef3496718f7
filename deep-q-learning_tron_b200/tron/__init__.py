"""Drop-in mirror of the reference's `tron` package (tron/map.py, tron/player.py, tron/game.py, tron/util.py).

Same class and function names, argument order and return types as the reference; every tick, observation and
pop_up is computed by the CUDA library through the C ABI (N=1 slice of the batched environment).  Put this
package's parent directory on sys.path (tron_b200.dropin.install()) to use `from tron.game import Game` unchanged.
"""
