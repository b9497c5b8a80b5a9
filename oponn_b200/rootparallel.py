"""Multi-GPU modes of the rollout path: one process per GPU over ``torch.distributed``.

* playout sharding: rank r plays a contiguous slice of every leaf's playouts; game ids are global, so the union of
  the shards is exactly the single-GPU batch.
* root-parallel merge: per-leaf (per root child) integer sums are combined with ONE all-reduce per move, the only
  exchange the path has.  Integer sums keep every running mean of ``MCTSTree.backpropagation``
  (src/com/mld46/oponn/MCTSTree.java:412-427) exact after the merge.

The rollout provider is anything with ``rollout_batch(leaves, playouts_per_leaf, seed, first_game, per_game=False)``
returning ``(agg, games)`` -- :class:`oponn_b200.engine.RolloutEngine` in production.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .state import LEAF_RESULT_DTYPE


def shard(total: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of rank's contiguous slice of ``total`` items; slices differ by at most one item."""
    lo = total * rank // world
    hi = total * (rank + 1) // world
    return lo, hi - lo


def merge_leaf_results(agg: np.ndarray, device=None) -> np.ndarray:
    """Sum ``OrxLeafResult`` records over all ranks (one all-reduce of 20 int64 per leaf)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return agg
    flat = torch.from_numpy(np.ascontiguousarray(agg).view(np.int64).copy())
    if device is not None:
        flat = flat.to(device)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat.cpu().numpy().view(LEAF_RESULT_DTYPE).reshape(agg.shape)


class ShardedRollout:
    """``rollout(leaves, playouts_per_leaf, seed)`` over all ranks: every rank returns the merged aggregates."""

    def __init__(self, provider, rank: int, world: int, device=None):
        self.provider, self.rank, self.world, self.device = provider, rank, world, device

    def rollout(self, leaves: np.ndarray, playouts_per_leaf: int, seed: int, first_game: int = 0) -> np.ndarray:
        leaves = np.ascontiguousarray(leaves)
        n = leaves.shape[0]
        lo, cnt = shard(playouts_per_leaf, self.world, self.rank)
        agg = np.zeros(n, dtype=LEAF_RESULT_DTYPE)
        if cnt:
            # leaf i, playout k has game id first_game + i * playouts_per_leaf + k on every layout: one call per
            # leaf keeps the ids of this rank's slice identical to the unsharded batch
            for i in range(n):
                a, _ = self.provider.rollout_batch(leaves[i:i + 1], cnt, seed, first_game + i * playouts_per_leaf + lo)
                agg[i] = a[0]
        return merge_leaf_results(agg, self.device)
