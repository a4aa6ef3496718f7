"""long trails with slide modes on the trail layout, tick by tick against the oracle: python tools/long_trail_case.py <slide> <W> <N> <eps>"""
import sys, os; sys.path.insert(0,"tests"); sys.path.insert(0,".")
import torch
import numpy as np
from tron_b200 import abi, _lib
from _gpu import make_pair, assert_same_step, assert_same_state
slide = int(sys.argv[1]); W = int(sys.argv[2]); N = int(sys.argv[3]); eps = float(sys.argv[4])
g,o = make_pair(N, W, W, layout="trail", obs_dtype=abi.I8, obs_enc=abi.ENC_NONE, seed=3, slide_mode=slide, slide_rate=0.15, policy=abi.POLICY_FREE_EPS, policy_epsilon=eps)
g.reset(); o.reset()
t = -1
try:
    for t in range(56):
        ex = o.export()
        assert_same_step(g.step(), o.step(), "tick %d" % t)
        torch.cuda.synchronize()
    assert_same_state(g, o)
    print("slide", slide, "W", W, "N", N, "eps", eps, "OK")
except Exception as e:
    tiles = ex["tiles"]; n_trail = ((tiles == 1) | (tiles == 3) | (tiles == 5) | (tiles == 6)).reshape(N, -1).sum(1)
    p1 = ((tiles == 1) | (tiles == 5)).reshape(N, -1).sum(1); p2 = ((tiles == 3) | (tiles == 6)).reshape(N, -1).sum(1)
    print("slide", slide, "W", W, "N", N, "eps", eps, "FAIL at tick", t, type(e).__name__, str(e)[:80].replace("\n"," "))
    print("  before the tick: max trail cells per game", n_trail.max(), "max P1", p1.max(), "max P2", p2.max(), "ep_len max", ex["ep_len"].max())
