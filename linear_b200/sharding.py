"""Read sharding across GPUs: reads are independent units, so every rank maps a contiguous
range of the block with its own replica of the index and genome features; there is no data-path collective.
Results are concatenated in input order on the host (the order `map_` appends them in, mapper.cpp:850-865)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_ranges(offsets: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous read ranges [lo, hi) per rank, balanced by bases (work is ~linear in read length)."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    total = int(offsets[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        cuts.append(int(np.searchsorted(offsets, np.uint64(target), side="left")))
    cuts.append(n)
    cuts = [min(max(c, 0), n) for c in cuts]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def take_shard(bases: np.ndarray, offsets: np.ndarray, lo: int, hi: int):
    offsets = np.asarray(offsets, dtype=np.uint64)
    b0, b1 = int(offsets[lo]), int(offsets[hi])
    return bases[b0:b1], (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint64)


def merge_cords(parts: Sequence[Tuple[np.ndarray, np.ndarray]]):
    """parts[r] = (cords, cords_off) of rank r, in rank order -> (cords, cords_off) of the whole block."""
    cords = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.uint64)
    offs = [np.zeros(1, np.uint64)]
    base = np.uint64(0)
    for c, o in parts:
        offs.append(np.asarray(o[1:], dtype=np.uint64) + base)
        base = base + np.uint64(len(c))
    return cords.astype(np.uint64), np.concatenate(offs)
