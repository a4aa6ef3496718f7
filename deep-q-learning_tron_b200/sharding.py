"""Env sharding across ranks: contiguous ranges of global env ids, no collective on the step path.

RNG streams are keyed by global env id (env_id_base + local index), so a run's trajectories do not depend on the number
of GPUs.  Only the optional data-parallel learn step communicates (DDQN.allreduce_gradients)."""
import os


def env_shard(n_total, rank=None, world=None):
    """-> (env_id_base, n_local): rank r owns [r*ceil(n/w), min(n, (r+1)*ceil(n/w)))"""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    per = (n_total + world - 1) // world
    base = min(n_total, rank * per)
    return base, max(0, min(n_total, base + per) - base)
