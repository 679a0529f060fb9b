"""GPU (-m gpu): randomised parity sweep. Small seeded genomes and read sets that vary everything the fixed cases hold
constant -- repeat density (comparator ties, dense DP windows, big buckets), N runs in genome and reads, read length and
error mix, -t (chunk seams) and preset (stop ratio) -- CUDA path vs the CPU oracle, bit-exact, stage by stage."""
import numpy as np
import pytest

from cases import _junk_and_chimeras, _pack
from cpu_checkers import Oracle
from linear_b200 import datagen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import linear_b200
    linear_b200.load_library()
    return linear_b200


@pytest.fixture(scope="module")
def ctx(lb):
    return lb.Context(0)


def make_fuzz_case(seed: int, genome_n: bool = True):
    rng = np.random.default_rng(1000 + seed)
    total = int(rng.integers(300_000, 900_000))
    n_contigs = int(rng.integers(1, 5))
    fam = int(rng.integers(0, 8))
    copies = int(rng.integers(50, 600)) if fam else 0
    tandem = int(rng.integers(0, 120))
    lens = datagen.contig_lengths(total, n_contigs, seed=seed)
    g = datagen.make_genome(seed, lens, n_families=fam, copies=copies, n_tandem=tandem)
    # N runs in the genome (hashInit N-skip at chunk starts, N arithmetic inside windows)
    for c in g:
        for _ in range(int(rng.integers(0, 4)) if genome_n else 0):
            p = int(rng.integers(0, max(len(c) - 400, 1)))
            c[p:p + int(rng.integers(1, 300))] = 4
    err = float(rng.choice([0.0, 0.01, 0.05, 0.12]))
    rs = datagen.simulate_reads(seed + 7, g, int(rng.integers(20, 40)), mean_len=int(rng.integers(1500, 9000)), sd_len=2000,
                                err=err, mix=(4, 3, 3), sv_frac=float(rng.choice([0.0, 0.3, 0.6])), lognormal=bool(seed & 1))
    reads = [rs.read(i) for i in range(rs.n)] + _junk_and_chimeras(rs, seed + 3, int(rng.integers(0, 5)))
    for r in reads[::5]:   # N inside reads (fast-path fallbacks of seeding and features)
        if len(r) > 600:
            p = int(rng.integers(0, len(r) - 300))
            r[p:p + int(rng.integers(1, 40))] = 4
    reads.append(reads[0][:int(rng.integers(180, 260))].copy())
    bases, offs = _pack(reads)
    T = int(rng.choice([1, 2, 3, 4, 7, 16]))
    preset = int(rng.choice([0, 1, 2]))
    return g, bases, offs, T, preset


@pytest.mark.parametrize("seed", range(32))
def test_random_case_is_bit_exact(lb, ctx, seed):
    g, bases, offs, T, preset = make_fuzz_case(seed)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    O = Oracle(g, threads=T, preset=preset)
    d0, h0 = O.dindex()
    d1, h1 = index.export_dindex()
    assert np.array_equal(d0, d1) and np.array_equal(h0, h1)
    cords, coff, dbg = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset, debug=True)
    oc, oo = O.map_batch(bases, offs, map_threads=4)
    for r in range(0, len(offs) - 1, 5):
        read = bases[int(offs[r]):int(offs[r + 1])]
        if len(read) <= 200:
            continue
        ra = dbg["ra"][int(dbg["ra_off"][r]):int(dbg["ra_off"][r + 1])]
        assert np.array_equal(ra, O.stage(read, 1)[1:]), (seed, "raw anchors", r)
        h = dbg["h"][int(dbg["h_off"][r]):int(dbg["h_off"][r + 1])]
        assert np.array_equal(h, O.stage(read, 3)), (seed, "hits", r)
    assert np.array_equal(oo, coff), seed
    assert np.array_equal(oc, cords), seed
    index.close(); feats.close(); gen.close()


@pytest.mark.parametrize("seed", range(100, 108))
def test_random_case_hindex_is_bit_exact(lb, ctx, seed):
    """-i 2 (HIndex) on the same kind of random cases (ACGT-only genomes: N genomes are not built for -i 2)"""
    g, bases, offs, T, preset = make_fuzz_case(seed, genome_n=False)
    T = min(T, 8)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 2, T)
    O = Oracle(g, threads=T, preset=preset, index_type=2)
    cords, coff = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    oc, oo = O.map_batch(bases, offs, map_threads=4)
    assert np.array_equal(oo, coff), seed
    assert np.array_equal(oc, cords), seed
    index.close(); feats.close(); gen.close()
