"""CPU-only checks of the C-ABI library: it loads, exports every declared symbol, agrees with the header on
struct layout, and refuses to compute without a GPU (no silent fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tron_b200  # noqa: E402
from tron_b200 import abi  # noqa: E402

HEADER = os.path.join(ROOT, "include", "tron_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from tron_b200 import _lib
    return _lib.load()


def test_header_declares_exactly_the_exported_symbols(lib):
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = set(re.findall(r"\b((?:tron|replay)_[a-z_]+)\s*\(", src))
    assert declared == set(abi.EXPORTED_SYMBOLS), declared ^ set(abi.EXPORTED_SYMBOLS)
    for name in abi.EXPORTED_SYMBOLS:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    prog = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "tron_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu\n", sizeof(tron_step_args), sizeof(replay_ring), sizeof(tron_meta), sizeof(tron_reward_t));
      printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", offsetof(tron_step_args, state), offsetof(tron_step_args, obs), offsetof(tron_step_args, lut),
             offsetof(tron_step_args, reward_table), offsetof(tron_step_args, seed), offsetof(tron_step_args, slide_tape),
             offsetof(tron_step_args, stats), offsetof(tron_step_args, n_ticks));
      printf("%zu %zu\n", offsetof(replay_ring, capacity), offsetof(replay_ring, done));
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(replay_frames), offsetof(replay_frames, rows), offsetof(replay_frames, frames), offsetof(replay_frames, done),
             offsetof(tron_step_args, obs_terminal), offsetof(tron_step_args, extra));
      return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "p"), os.path.join(d, "p.c")])
        out = subprocess.check_output([os.path.join(d, "p")]).decode().split()
    v = list(map(int, out))
    S, R = abi.StepArgs, abi.ReplayRing
    assert v[:4] == [C.sizeof(S), C.sizeof(R), 8, C.sizeof(abi.Reward)]
    assert v[4:12] == [S.state.offset, S.obs.offset, S.lut.offset, S.reward_table.offset, S.seed.offset, S.slide_tape.offset,
                       S.stats.offset, S.n_ticks.offset]
    assert v[12:14] == [R.capacity.offset, R.done.offset]
    F = abi.ReplayFrames
    assert v[14:] == [C.sizeof(F), F.rows.offset, F.frames.offset, F.done.offset, S.obs_terminal.offset, S.extra.offset]


def test_geometry_helpers_host_only(lib):
    n = C.c_size_t()
    assert lib.tron_state_bytes(4096, 10, 10, 0, C.byref(n)) == 0 and n.value == abi.state_bytes(4096, 10, 10)
    assert lib.tron_state_bytes(5, 64, 64, 0, C.byref(n)) == 0 and n.value == abi.state_bytes(5, 64, 64)
    assert lib.tron_state_bytes(5, 1, 10, 0, C.byref(n)) == abi.ERR_INVALID
    assert lib.tron_cells_per_env(10, 10) == 144 and lib.tron_cells_per_env(64, 64) == 4356
    assert [lib.tron_enc_planes(e) for e in range(4)] == [0, 1, 3, 4]
    assert [lib.tron_dtype_size(d) for d in (abi.U8, abi.I32, abi.I64, abi.BF16, abi.F32, abi.I8)] == [1, 4, 8, 2, 4, 1]
    assert lib.tron_abi_version() == abi.ABI_VERSION


def test_state_bytes_of_every_layout_match_the_library(lib):
    """the Python mirror of tron_state_bytes (used to size torch allocations) agrees with the C side for all four layouts"""
    n = C.c_size_t()
    for layout in (abi.LAYOUT_TILE8, abi.LAYOUT_BITS10, abi.LAYOUT_TRAIL, abi.LAYOUT_BITS):
        for (w, h) in ((10, 10), (2, 2), (3, 4), (7, 9), (12, 12), (33, 20), (64, 64), (126, 126)):
            if layout == abi.LAYOUT_BITS10 and (w, h) != (10, 10):
                continue
            if layout == abi.LAYOUT_BITS and w * h > 128:
                continue
            for n_envs in (1, 17, 4096):
                assert lib.tron_state_bytes(n_envs, w, h, layout, C.byref(n)) == 0, (layout, w, h)
                assert n.value == abi.state_bytes(n_envs, w, h, layout), (layout, w, h, n_envs)


def test_auto_layout_choices():
    """layout="auto": bit planes on config.py's board, trail lists for pure ticks on large grids and for fused observations from
    12x12 up wherever the bulk-store kernel applies, the int8 grid otherwise"""
    import torch
    from tron_b200.batch_env import auto_layout
    assert auto_layout(10, 10, "lut1", None) == "bits10" and auto_layout(10, 10, "popup3", "temper") == "bits"
    assert auto_layout(64, 64, "none", None) == "trail" and auto_layout(8, 8, "none", None) == "tile8"
    assert auto_layout(8, 8, "lut1", None) == "tile8" and auto_layout(11, 11, "lut1", None) == "tile8"
    assert auto_layout(12, 12, "lut1", None) == "trail" and auto_layout(21, 21, "lut1", None, torch.int8) == "trail"
    assert auto_layout(64, 64, "popup3", "ice", torch.float32) == "trail"
    assert abi.trail_bulk_ok(126, 126, abi.ENC_LUT1, abi.BF16) and not abi.trail_bulk_ok(126, 126, abi.ENC_POPUP3_CONST, abi.F32)
    assert auto_layout(126, 126, "popup3_const", None, torch.float32) == "tile8"  # rows too long for the shared-memory group buffer


def test_plane_tables_match_oracle(lib):
    from oracle import c_oracle as oc
    for enc in (abi.ENC_LUT1, abi.ENC_POPUP3, abi.ENC_POPUP3_CONST):
        for lut in ((0,) * 6, (0, -1, -2, -3, 10, -10), (5, 4, 3, 2, 1, -7)):
            l6 = (C.c_int8 * 6)(*lut)
            a = (C.c_int8 * 48)(); b = (C.c_int8 * 48)()
            ra = lib.tron_build_plane_tables(l6, enc, a)
            rb = oc.lib().oracle_build_plane_tables(l6, enc, b)
            assert ra == rb and list(a) == list(b)
    t = (C.c_int8 * 48)()
    lib.tron_build_plane_tables((C.c_int8 * 6)(), abi.ENC_LUT1, t)
    assert list(t)[:8] == [-1, 1, -2, 10, -3, -10, -2, -3]  # Tile.value -1..6 from P1's side (reference map.py:67-81)
    assert list(t)[8:16] == [-1, 1, -3, -10, -2, 10, -3, -2]


def test_no_silent_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tron_b200 import _lib
    assert lib.tron_device_count() <= 0
    with pytest.raises(_lib.TronError):
        tron_b200.BatchedTron(16)
    with pytest.raises(_lib.TronError):
        tron_b200.ReplayRing(16, (4,))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deep-q-learning_tron_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(dirpath, f)
