"""tools/shard.py -- how a multi-GPU run splits the work: the table is replicated, sequences are
sharded by rank, there is no data-path collective; only the timing / counters are reduced at the end."""
from __future__ import annotations

import numpy as np


def weak_shard(rank: int, per_rank: int) -> tuple[int, int]:
    """Weak scaling (bench.py): rank r owns synthetic proteins [r*per_rank, (r+1)*per_rank)."""
    return rank * per_rank, per_rank


def balanced_cuts(offsets: np.ndarray, world: int) -> np.ndarray:
    """Strong scaling of a real FASTA: world contiguous shards balanced by residues (sequences are independent units,
    KGJ:528/540).  Returns world+1 sequence indices; shard r = sequences [cuts[r], cuts[r+1])."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    total = int(offsets[-1])
    targets = (np.arange(1, world, dtype=np.float64) * total / world).astype(np.uint64)
    inner = np.searchsorted(offsets, targets, side="left").astype(np.int64)
    return np.concatenate([[0], np.clip(inner, 0, n), [n]]).astype(np.int64)


def merge_records(parts: list[np.ndarray], cuts: np.ndarray, field: str = "seq") -> np.ndarray:
    """Result gather: per-shard record arrays (sequence indices relative to the shard) -> one array in global order."""
    out = []
    for r, p in enumerate(parts):
        q = p.copy()
        q[field] = q[field] + np.uint32(cuts[r])
        out.append(q)
    return np.concatenate(out) if out else np.zeros(0)
