"""GPU (-m gpu): BASELINE.json full sizes (3.1-Gbase, 24-contig genome) through size-independent properties -- the CPU
oracle needs minutes at this size, so here the index is checked by what must hold for ANY correct DIndex, plus one
cross-check that does not need an oracle: the hash-range sharded build must assemble to the very same arrays."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dindex_3gbase_properties_and_sharded_equivalence():
    import torch
    import linear_b200 as lb
    from linear_b200 import sharding
    sys.path.insert(0, ROOT)
    import bench
    dev = torch.device("cuda", 0)
    lens = bench.contig_lengths()
    genome = bench.gen_genome(torch, dev, lens)
    G = int(sum(lens))
    assert G > 3_000_000_000 and len(lens) == 24 and max(lens) < 299_000_000
    ctx = lb.Context(0)
    gen = lb.Genome(ctx, device_ptr=genome.data_ptr(), lens=[int(x) for x in lens])
    T = 16
    index = lb.create_index(ctx, gen, 1, T)
    d, hs = index.export_device(torch, dev)
    n_hs = hs.numel()
    cnt = d[1:] - d[:-1]
    # bucket offsets: exclusive prefix, no bucket above the omission threshold, total = n_hs
    assert int(d[0]) == 0 and int(d[-1]) == n_hs and int(cnt.min()) >= 0 and int(cnt.max()) <= 400
    # one record per ~9 bases minus omitted repeats
    assert 0.09 * G < n_hs < 0.112 * G
    # ascending inside every bucket: hs[i] < hs[i+1] unless i+1 starts a bucket
    starts = torch.zeros(n_hs + 1, dtype=torch.bool, device=dev)
    starts[d.long()] = True
    asc = hs[1:] > hs[:-1]
    assert bool((asc | starts[1:n_hs]).all())
    # every record decodes to (contig, sampled position): id < 24, position inside the contig's sampled range, Y < 256,
    # and the position lies on its chunk's 9-spaced sampling grid (index_util.cpp:1654-1685)
    x = (hs >> 20) & ((1 << 30) - 1)
    cid = (hs >> 50) & 1023
    y = hs & 0xFFFFF
    lens_t = torch.tensor(lens, dtype=torch.int64, device=dev)
    assert int(cid.max()) < 24 and int(y.max()) < 256
    pos = x - (1 << 20)
    L = lens_t[cid]
    assert bool((pos >= 21 + 8).all()) and bool((pos < L - 42).all())
    chunk = torch.minimum(pos // (L // T), torch.tensor(T - 1, device=dev))
    # a position may belong to chunk c or (if it is before c's first sample) to c-1; check both candidates
    ok = torch.zeros_like(pos, dtype=torch.bool)
    for dc in (0, 1):
        c = (chunk - dc).clamp_min(0)
        t_str = (L // T) * c + 21
        ok |= ((pos - t_str - 8) % 9 == 0) & (pos >= t_str + 8)
    assert bool(ok.all())
    del starts, asc, x, cid, y, pos, L, chunk, ok
    # hash-range sharded build (what N GPUs do) assembles to identical arrays
    for n_shards in (8,):
        parts = []
        for s in range(n_shards):
            p = lb.Index(ctx, gen, 1, T, shard=s, n_shards=n_shards)
            parts.append(p.export_device(torch, dev))
            p.close()
        d2, h2 = sharding.assemble_dindex(parts, xp=torch)
        assert torch.equal(d2, d) and torch.equal(h2, hs)
    # genome features: entry count and field sanity (every 6-bit field <= 48, 47 two-mers minus TT per entry)
    feats = lb.create_features(ctx, gen, 2, T)
    f = torch.from_numpy(feats.download(0)).to(dev)
    assert f.shape[0] == ((lens[0] - 48) >> 4) + 1
    tot = torch.zeros(f.shape[0], dtype=torch.int64, device=dev)
    for k in range(3):
        for sft in range(0, 30, 6):
            fld = (f[:, k].long() >> sft) & 63
            assert int(fld.max()) <= 48
            tot += fld
    assert int(tot.max()) <= 48 and int(tot.min()) >= 20


def test_config2_cords_and_dindex_vs_reference():
    """BASELINE configs[1] + configs[2] against the UNMODIFIED reference (oracle/_ref) on the very inputs bench.py times:
    the whole 3.1-Gbase DIndex (dir + hs, -t 16) and the cords of ONT-like reads incl. planted-SV reads. Regimes the small
    cases never reach: contig ids up to 23, binningFilter bins up to ~9970 (pmpfinder.cpp:1984-1996), ~27 bucket records
    per seed, lookup tails beyond 24 keys."""
    import torch
    import linear_b200 as lb
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    from cpu_checkers import RefImpl, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref (the compiled reference) did not travel")
    dev = torch.device("cuda", 0)
    lens = bench.contig_lengths()
    genome = bench.gen_genome(torch, dev, lens)
    n_reads = 768
    bases_t, offs = bench.gen_reads(torch, dev, genome, lens, n_reads, seed=4242)
    bases = bases_t.cpu().numpy()
    gh = genome.cpu().numpy()
    coff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    contigs = [gh[coff[i]:coff[i + 1]] for i in range(len(lens))]
    T = 16
    ctx = lb.Context(0)
    gen = lb.Genome(ctx, device_ptr=genome.data_ptr(), lens=[int(x) for x in lens])
    del genome, bases_t
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    cords, c_off = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=1)
    ref = RefImpl(contigs, threads=T, preset=1)
    r_cords, r_off = ref.map_batch(bases, offs, map_threads=os.cpu_count() or 1)
    par = bench.check_parity(torch, dev, ref, index, r_cords, r_off, cords, c_off)
    # genome features of the last (24th) contig as well
    f_gpu = feats.download(len(lens) - 1)
    f_ref = ref.genome_features(len(lens) - 1)
    ref.close()
    assert par["cords_equal"], par
    assert par["dindex_equal"], par
    assert par["cords"] > 50 * n_reads          # the reads really mapped
    assert np.array_equal(f_gpu, f_ref)


def test_hindex_at_200_mbase_returns_and_equals_the_oracle():
    """-i 2 beyond the small genomes. From ~200 Mbase on HIndex blocks reach the 1024-record limit and get virtual heads; a
    read seed that meets one with YValue == 0 sends the reference's getXDir (index_util.cpp:1076-1090) round the same probe
    sequence for ever -- on this very input `ref_read_stage(read 8, stage 1)` did not return within 150 s (round 2 log) --
    while the (key -> value) map has no such entry. The library answers with a miss, like the oracle's exact-match map:
    raw anchors and cords of all reads equal the oracle's, and the call returns."""
    import linear_b200 as lb
    from linear_b200 import datagen
    from cpu_checkers import Oracle
    lens = datagen.contig_lengths(200_000_000, 4, seed=31)
    rng = np.random.default_rng(5)
    g = [rng.integers(0, 4, size=int(l), dtype=np.uint8) for l in lens]
    rs = datagen.simulate_reads(9, g, 16, mean_len=15000, sd_len=3000, err=0.1, sv_frac=0.2)
    ctx = lb.Context(0)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, 16)
    index = lb.create_index(ctx, gen, 2, 16)
    cords, coff, dbg = lb.apx_map_batch(ctx, index, feats, rs.bases, rs.offsets, preset=1, debug=True)
    O = Oracle(g, threads=16, preset=1, index_type=2)
    for i in range(rs.n):
        a = O.stage(rs.read(i), 1)[1:]
        got = dbg["ra"][int(dbg["ra_off"][i]):int(dbg["ra_off"][i + 1])]
        assert np.array_equal(a, got), f"raw anchors of read {i}"
    oc, oo = O.map_batch(rs.bases, rs.offsets, map_threads=8)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords)
