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


# ---- hash-range sharded DIndex build ---------------------------------------------------------------------
N_BUCKETS = 1 << 26


def assemble_dindex(parts, xp=np):
    """parts[s] = (dir_s, hs_s) of shard s (lnr_index_build_shard): dir_s is the full 2^26+1 exclusive prefix over the
    shard's own records, hs_s its records. Returns the (dir, hs) of the whole index. `xp` is numpy or torch: with torch
    device tensors this is the local assembly step after the all-gather."""
    n = len(parts)
    per = N_BUCKETS // n
    counts = [int(p[1].shape[0]) for p in parts]
    base = 0
    slices = []
    if sum(counts) >= 1 << 31:
        raise ValueError("assembled hs exceeds int32 bucket offsets (index_util.h:101)")
    for s, (d, _) in enumerate(parts):
        slices.append(d[s * per:(s + 1) * per] + base)
        base += counts[s]
    if xp is np:
        dir_ = np.concatenate(slices + [np.array([base], dtype=slices[0].dtype)])
        hs = np.concatenate([p[1] for p in parts])
    else:
        dir_ = xp.cat(slices + [xp.tensor([base], dtype=slices[0].dtype, device=slices[0].device)])
        hs = xp.cat([p[1] for p in parts])
    return dir_, hs


def build_index_sharded(lb, ctx, genome, threads, rank, world, torch, dist, device):
    """Every rank builds the buckets of its minimizer range, then ONE exchange step over NCCL: all-gather of the record
    counts, of the (padded) hs slices and of the rebased dir slices. Returns an Index holding the whole DIndex."""
    part = lb.Index(ctx, genome, 1, threads, shard=rank, n_shards=world)
    d_loc, hs_loc = part.export_device(torch, device)
    part.close()
    n_loc = torch.tensor([hs_loc.numel()], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_loc) for _ in range(world)]
    dist.all_gather(counts, n_loc)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    pad = torch.zeros(n_max, dtype=torch.int64, device=device)
    pad[: hs_loc.numel()] = hs_loc
    hs_all = torch.empty(world * n_max, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(hs_all, pad)
    per = N_BUCKETS // world
    base = sum(counts[:rank])
    my_slice = (d_loc[rank * per:(rank + 1) * per] + base).contiguous()
    dir_all = torch.empty(N_BUCKETS + 1, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(dir_all[:N_BUCKETS], my_slice)
    dir_all[N_BUCKETS] = sum(counts)
    hs = torch.cat([hs_all[r * n_max: r * n_max + counts[r]] for r in range(world)])
    if sum(counts) >= 1 << 31:
        raise ValueError("assembled hs exceeds int32 bucket offsets (index_util.h:101)")
    # the library copies on its own (non-blocking) stream: the gathers / cat queued on torch's streams must be complete
    torch.cuda.synchronize(device)
    return lb.Index.from_device(ctx, dir_all, hs)
