import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from tron_b200 import abi, _lib
from _gpu import make_pair, assert_same_step, assert_same_state
import ctypes as C

def viol():
    c = C.c_uint64(); f = C.c_int32()
    rc = _lib.load().tron_debug_violations(C.byref(c), C.byref(f))
    return rc, c.value, f.value

W = int(os.environ.get("W", 8)); N = int(os.environ.get("N", 3000)); eps = float(os.environ.get("EPS", 0.05))
for slide in (abi.SLIDE_TEMPER, abi.SLIDE_NONE, abi.SLIDE_ICE):
    g, o = make_pair(N, W, W, layout="trail", obs_dtype=abi.F32, obs_enc=abi.ENC_NONE, seed=77, slide_mode=slide, slide_rate=0.15,
                     policy=abi.POLICY_FREE_EPS, policy_epsilon=eps)
    g.reset(); o.reset()
    try:
        for t in range(60):
            assert_same_step(g.step(), o.step(), "slide %d tick %d" % (slide, t))
            torch.cuda.synchronize()
        for T in (2, 5, 9):
            assert_same_step(g.step_many(T), o.step_many(T), "slide %d step_many %d" % (slide, T))
            torch.cuda.synchronize()
        assert_same_state(g, o)
        print("slide", slide, "ok", viol())
    except AssertionError as e:
        print("slide", slide, "MISMATCH", str(e)[:200], viol())
