"""CPU, world_size 2 over gloo: the N > 1 host path (shard reads by bases -> map each shard independently ->
concatenate in input order) gives exactly the unsharded result. The per-shard map is done by the CPU oracle here
(no GPU in this container); on the GPU box bench.py runs the same sharding with the CUDA path per rank."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from cases import make_case
    from cpu_checkers import Oracle
    from linear_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, reads, bases, offs, T, preset = make_case("clean_hifi")
    O = Oracle(g, threads=T, preset=preset)           # index replicated on every rank
    lo, hi = sharding.shard_ranges(offs, world)[rank]
    sb, so = sharding.take_shard(bases, offs, lo, hi)
    part = O.map_batch(sb, so, map_threads=1)
    gathered = [None] * world
    dist.all_gather_object(gathered, (part[0], part[1], lo, hi))
    dist.barrier()
    if rank == 0:
        cords, coff = sharding.merge_cords([(c, o) for c, o, _, _ in gathered])
        full_c, full_o = O.map_batch(bases, offs, map_threads=2)
        ranges = [(a, b) for _, _, a, b in gathered]
        q.put((bool(np.array_equal(cords, full_c) and np.array_equal(coff, full_o)), ranges, len(offs) - 1))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, ranges, n = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == n


def test_shard_ranges_cover_and_balance():
    from linear_b200 import sharding
    rng = np.random.default_rng(0)
    lens = rng.integers(1, 50000, size=1000)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 3, 8):
        rs = sharding.shard_ranges(offs, world)
        assert rs[0][0] == 0 and rs[-1][1] == 1000
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        per = [int(offs[b] - offs[a]) for a, b in rs]
        assert max(per) - min(per) <= 2 * 50000
    assert sharding.shard_ranges(np.zeros(1, np.uint64), 4) == [(0, 0)] * 4


def test_assemble_dindex_from_hash_range_shards():
    """host logic of the multi-GPU index assembly: split a full DIndex by minimizer range the way lnr_index_build_shard
    does, reassemble, compare"""
    from cases import make_case
    from cpu_checkers import Oracle
    from linear_b200 import sharding
    g, reads, bases, offs, T, preset = make_case("clean_hifi")
    d, hs = Oracle(g, threads=T, preset=preset).dindex()
    cnt = np.diff(d).astype(np.int64)
    NB = sharding.N_BUCKETS
    for n in (1, 2, 4, 8):
        per = NB // n
        parts = []
        for s in range(n):
            c = np.zeros(NB, dtype=np.int64)
            c[s * per:(s + 1) * per] = cnt[s * per:(s + 1) * per]
            ds = np.concatenate([[0], np.cumsum(c)]).astype(np.int32)
            parts.append((ds, hs[d[s * per]:d[(s + 1) * per]]))
        d2, h2 = sharding.assemble_dindex(parts)
        assert np.array_equal(d2, d) and np.array_equal(h2, hs)


def test_hindex_shard_cuts_balance_and_reassemble():
    """host logic of the sharded HIndex build (lnr_index_build_sharded, index_type 2): the X ranges every rank derives from the
    per-X pair histogram are ordered, cover the 18-bit X axis, balance the pairs up to one X's worth, and -- a block being one
    X -- the ranks' slices of ysa, concatenated in rank order, are ysa itself"""
    from cases import make_case
    from cpu_checkers import Oracle
    import linear_b200 as lb
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    ysa, empty_dir, kv, tl = Oracle(g, threads=T, preset=preset, index_type=2).hindex()
    words = ysa[:empty_dir]
    heads = np.flatnonzero((words >> np.uint64(63)) == 0)          # every body carries bit 63
    X = (words[heads] & np.uint64((1 << 40) - 1)).astype(np.int64)
    n_body = (words[heads] >> np.uint64(40)).astype(np.int64) - 1  # head = (bodies + 1) << 40 | X
    assert np.all(np.diff(X) > 0) and int(n_body.sum()) + len(heads) == empty_dir
    NX = 1 << 18
    hist = np.zeros(NX, dtype=np.uint32)
    hist[X] = n_body
    total = int(hist.sum())
    for n in (1, 2, 3, 8, 64):
        cuts = lb.hindex_shard_cuts(hist, n).astype(np.int64)
        assert cuts[0] == 0 and cuts[n] == NX and np.all(np.diff(cuts) >= 0)
        per = [int(hist[cuts[r]:cuts[r + 1]].sum()) for r in range(n)]
        assert sum(per) == total
        assert max(per) <= -(-total // n) + int(hist.max())
        pieces = []
        for r in range(n):
            blk = np.flatnonzero((X >= cuts[r]) & (X < cuts[r + 1]))
            if len(blk):
                pieces.append(words[heads[blk[0]]:heads[blk[-1]] + n_body[blk[-1]] + 1])
                assert len(pieces[-1]) == per[r] + len(blk)         # pairs + blocks: what the rank reports in the all-gather
        assert np.array_equal(np.concatenate(pieces), words)
    # degenerate inputs
    assert list(lb.hindex_shard_cuts(np.zeros(16, np.uint32), 4)) == [0, 0, 0, 0, 16]
    assert list(lb.hindex_shard_cuts(np.array([5, 0, 0, 0], np.uint32), 2)) == [0, 1, 4]
