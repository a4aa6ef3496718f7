import sys; sys.path.insert(0, "tools")
from sweep import run
M = 1 << 20
for v in (0, 8):
    run("6x6 bf16 1-plane tile8 variant %d" % v, 8 * M, 6, "bf16", "lut1", layout="tile8", actions="rng", variant=v)
    run("8x8 bf16 1-plane tile8 variant %d" % v, 4 * M, 8, "bf16", "lut1", layout="tile8", actions="rng", variant=v)
    run("8x8 bf16 pop_up3 tile8 variant %d" % v, 2 * M, 8, "bf16", "popup3", layout="tile8", actions="rng", variant=v)
    run("8x8 f32 1-plane tile8 variant %d" % v, 2 * M, 8, "f32", "lut1", layout="tile8", actions="rng", variant=v)
    run("8x8 bf16 1-plane bits variant %d" % v, 4 * M, 8, "bf16", "lut1", layout="bits", actions="rng", variant=v)
    run("12x12 bf16 1-plane tile8 variant %d" % v, 2 * M, 12, "bf16", "lut1", layout="tile8", actions="rng", variant=v)
    run("16x16 bf16 1-plane tile8 variant %d" % v, 1 * M, 16, "bf16", "lut1", layout="tile8", actions="rng", variant=v)
    run("20x20 bf16 1-plane tile8 variant %d" % v, 1 * M, 20, "bf16", "lut1", layout="tile8", actions="rng", variant=v)
