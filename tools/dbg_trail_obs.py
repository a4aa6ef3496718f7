import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tron_b200 import abi
from _gpu import make_pair
for W, N, dt in ((64, 300, abi.BF16), (126, 9, abi.BF16), (40, 64, abi.BF16), (32,64,abi.F32)):
    g, o = make_pair(N, W, W, layout="trail", obs_dtype=dt, obs_enc=abi.ENC_LUT1, seed=80 + W)
    a, b = g.reset(), o.reset()
    a = np.asarray(a, dtype=np.float32); b = np.asarray(b, dtype=np.float32)
    d = np.argwhere(a != b)
    print("W", W, "N", N, "shape", a.shape, "mismatches", len(d))
    if len(d):
        print(" first", d[:8].tolist(), "envs", sorted(set(d[:, 0].tolist()))[:20], "n envs", len(set(d[:, 0].tolist())))
        i = tuple(d[0]); print(" got", a[i], "want", b[i])
        flat = np.argwhere(a.reshape(a.shape[0], -1) != b.reshape(b.shape[0], -1))
        print(" flat offsets min/max", flat[:, 1].min(), flat[:, 1].max(), "distinct", len(set(flat[:, 1].tolist())))
        for e in sorted(set(flat[:, 0].tolist()))[:4]:
            offs = flat[flat[:, 0] == e][:, 1]
            print("  env", e, "n", len(offs), "min", offs.min(), "max", offs.max())
