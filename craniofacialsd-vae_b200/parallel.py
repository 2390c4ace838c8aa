"""Sharding rules of the data-parallel training step (pure host logic, no CUDA).

The global batch is the ``bs x bs`` swap grid (element ``i*bs+j`` = base mesh i with the swapped
region of mesh j, swap_batch_transform.py:27-38).  It is sharded by GRID ROWS, never re-squared per
rank, because the latent-consistency loss (model_manager.py:360-393) is defined on the whole grid.
"""
from __future__ import annotations

from typing import Tuple


def grid_rows(bs: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows ``[i0, i1)`` of the swap grid owned by ``rank``."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError('bad rank %d / world %d' % (rank, world))
    if bs % world:
        raise ValueError('grid rows (batch_size=%d) must divide over %d ranks' % (bs, world))
    rows = bs // world
    return rank * rows, (rank + 1) * rows


def local_meshes(bs: int, world: int, rank: int) -> Tuple[int, int]:
    """Slice ``[lo, hi)`` of the ``bs*bs`` swapped meshes (and of ``z``) owned by ``rank``."""
    i0, i1 = grid_rows(bs, world, rank)
    return i0 * bs, i1 * bs


def mean_loss_scale(world: int) -> float:
    """Factor that turns a rank-local mean (MSE, KL, Laplacian) into its share of the global
    mean, so that SUM all-reduce of losses and gradients reproduces the single-GPU values."""
    return 1.0 / float(world)
