"""The pure-Python loop port (bench.py's reference arm) against the reference's digests and, in the authoring
container, against the live reference cell for cell."""
import hashlib
import os
import sys
import time

import numpy as np
import pytest

from oracle import py_port as pp

from _golden import load_json, load_npz


def digest_port(seed, ngames, W=10):
    rng = np.random.default_rng(seed)
    h = hashlib.sha256(); steps = 0; wins = [0, 0, 0]
    for _ in range(ngames):
        while True:
            x1, y1, x2, y2 = (int(v) for v in rng.integers(0, [W, W, W, W]))
            if (x1, y1) != (x2, y2):
                break
        g = pp.PyGame(W, W, (x1, y1), (x2, y2))
        h.update(g.board().view_for(1).astype(np.int8).tobytes())
        done = False
        while not done:
            a1, a2 = (int(v) for v in rng.integers(0, 4, size=2))
            s1, s2, done = g.step(a1, a2); steps += 1
            h.update(s1.astype(np.int8).tobytes()); h.update(s2.astype(np.int8).tobytes())
            h.update(bytes([int(done), g.winner or 0]))
        wins[g.winner or 0] += 1
    return h.hexdigest(), steps, wins


def test_port_reproduces_reference_digest():
    want = load_json("digests.json")["0"]
    got = digest_port(0, 1000)
    assert got[0] == want["sha256"] and got[1] == want["steps"] and got[2] == want["wins"]


def test_port_pop_up_matches_reference_fixture():
    pu = load_npz("popup.npz")
    for o, p in zip(pu["obs"][:16], pu["planes"][:16]):
        assert (pp.pop_up(o.astype(np.int64)) == p).all()


def test_port_slide_mode_matches_fixture():
    sl = load_npz("slide.npz")
    starts = np.concatenate([[0], np.cumsum(sl["length"] + 1)[:-1]])
    for gi in range(0, len(starts), 5):
        s = int(starts[gi]); sp = sl["spawn"][gi]
        tape = {"cur": (0, 0)}
        g = pp.PyGame(10, 10, (int(sp[0]), int(sp[1])), (int(sp[2]), int(sp[3])), mode="ice", bernoulli=lambda i: bool(tape["cur"][i]))
        for t in range(int(sl["length"][gi])):
            row = s + 1 + t
            tape["cur"] = tuple(int(x) for x in sl["slide"][row])
            o1, o2, done = g.step(int(sl["actions"][row][0]), int(sl["actions"][row][1]))
            assert (o1.astype(np.int8) == sl["obs1"][row]).all() and (o2.astype(np.int8) == sl["obs2"][row]).all()
            assert int(done) == int(sl["done"][row]) and (g.winner or 0) == int(sl["winner"][row])


@pytest.mark.reference
def test_port_equals_live_reference_and_costs_the_same():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg
    G, U, P = mg.import_reference()
    rng = np.random.default_rng(3)
    t_ref = t_port = 0.0
    steps = 0
    for _ in range(150):
        W = int(rng.integers(3, 13))
        while True:
            sp = tuple(int(v) for v in rng.integers(0, W, size=4))
            if sp[:2] != sp[2:]:
                break
        ref = mg.new_game(G, P, W, sp)
        port = pp.PyGame(W, W, sp[:2], sp[2:])
        while not ref.done:
            a1, a2 = (int(v) for v in rng.integers(0, 4, size=2))
            t0 = time.perf_counter(); r1, r2, rd = ref.step(a1, a2); t1 = time.perf_counter()
            p1, p2, pd = port.step(a1, a2); t2 = time.perf_counter()
            t_ref += t1 - t0; t_port += t2 - t1; steps += 1
            assert (np.asarray(r1) == np.asarray(p1)).all() and (np.asarray(r2) == np.asarray(p2)).all() and rd == pd
            assert (ref.winner or 0) == (port.winner or 0)
            assert [list(p.position) for p in ref.pps] == [list(p) for p in port.pos]
            assert mg.tiles_of(ref.history[-1].map).tolist() == [[t.value for t in row] for row in port.history[-1].board.cells]
    print("\n[py_port] %d ticks: reference %.0f steps/s, port %.0f steps/s (ratio %.2f)" % (steps, steps / t_ref, steps / t_port, t_ref / t_port))
    assert 0.5 < t_ref / t_port < 2.0  # same order of cost as the thing it stands in for
