"""CPU: the oracle restatement against (a) golden digests produced by the real reference and (b) the real
reference itself when oracle/_ref is present. Parity is bit-exact (integer / byte work)."""
import json
import os

import numpy as np
import pytest

from cases import CASES, digest, make_case
from cpu_checkers import Oracle, RefImpl, have_ref

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    g, reads, bases, offs, T, preset = make_case(request.param)
    return request.param, g, reads, bases, offs, T, preset, Oracle(g, threads=T, preset=preset)


def test_oracle_index_and_features_match_golden(case):
    name, g, reads, bases, offs, T, preset, O = case
    gd = GOLDEN[name]
    d, hs = O.dindex()
    assert len(d) == (1 << 26) + 1
    nz = np.flatnonzero(np.diff(d))
    assert len(hs) == gd["n_hs"] and digest(hs) == gd["hs"]
    assert digest(nz.astype(np.int64)) == gd["dir_nonzero"]
    assert digest(np.diff(d)[nz].astype(np.int32)) == gd["dir_counts"]
    for i in range(len(g)):
        assert digest(O.genome_features(i)[:-1]) == gd["genome_features"][i]


def test_oracle_stages_match_golden(case):
    name, g, reads, bases, offs, T, preset, O = case
    gd = GOLDEN[name]
    for i, r in enumerate(reads):
        if not gd["stable"][i]:
            continue
        c = O.cords(r) if len(r) > 200 else np.zeros(0, np.uint64)
        assert len(c) == gd["n_cords"][i] and digest(c) == gd["cords"][i], f"read {i}"
        if len(r) > 200:
            assert digest(O.stage(r, 1)[1:]) == gd["raw_anchors"][i], f"anchors read {i}"
            assert digest(O.stage(r, 3)) == gd["hits"][i], f"hits read {i}"


def test_oracle_batch_matches_per_read(case):
    name, g, reads, bases, offs, T, preset, O = case
    cords, coff = O.map_batch(bases, offs, map_threads=2)
    gd = GOLDEN[name]
    assert [int(coff[i + 1] - coff[i]) for i in range(len(reads))] == [
        n if len(r) > 200 else 0 for n, r in zip(gd["n_cords"], reads)]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_reference_all_stages():
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    O = Oracle(g, threads=T, preset=preset)
    R = RefImpl(g, threads=T, preset=preset)
    d0, h0 = R.dindex()
    d1, h1 = O.dindex()
    assert np.array_equal(d0, d1) and np.array_equal(h0, h1)
    for r in reads[:40]:
        if len(r) <= 200:
            continue
        for st in (0, 1):
            assert np.array_equal(R.read_features(r, st), O.read_features(r, st))
        for stage in (1, 2, 5, 3, 4, 0):
            assert np.array_equal(R.stage(r, stage), O.stage(r, stage)), f"stage {stage}"
        # a re-map shaped seeding call (str > 0, step 7): splice + selector bias
        a = R.stage(r, 1, len(r) // 3, len(r) - 100, 1)
        b = O.stage(r, 1, len(r) // 3, len(r) - 100, 1)
        assert np.array_equal(a, b)
    rc, ro = R.map_batch(bases, offs, map_threads=2)
    oc, oo = O.map_batch(bases, offs, map_threads=2)
    assert np.array_equal(ro, oo) and np.array_equal(rc, oc)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_reference_with_N():
    """N handling: ord(N)=4 is added, not masked (shape_extend.cpp:176-180); N two-mers add nothing to features."""
    from linear_b200 import datagen
    lens = datagen.contig_lengths(600_000, 2, seed=4)
    g = datagen.make_genome(21, lens, n_families=1, copies=30, n_tandem=4, n_runs=40)
    rs = datagen.simulate_reads(5, g, 30, mean_len=5000, sd_len=1000, err=0.03)
    O = Oracle(g, threads=4)
    R = RefImpl(g, threads=4)
    d0, h0 = R.dindex()
    d1, h1 = O.dindex()
    assert np.array_equal(d0, d1) and np.array_equal(h0, h1)
    for i in range(len(g)):
        assert np.array_equal(R.genome_features(i)[:-1], O.genome_features(i)[:-1])
    for i in range(rs.n):
        r = rs.read(i)
        assert np.array_equal(R.stage(r, 1), O.stage(r, 1))
        assert np.array_equal(R.cords(r), O.cords(r))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("threads", [1, 4, 8])
def test_oracle_hindex_equals_reference(threads):
    """-i 2: ysa byte-exact (incl. the chunk-tail mislabel), emptyDir, and the open-addressing directory
    as a sorted (val1, val2) list -- its physical layout is nondeterministic in the reference itself."""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    O = Oracle(g, threads=threads, preset=preset, index_type=2)
    R = RefImpl(g, threads=threads, preset=preset, index_type=2)
    y0, e0, kv0, tl0 = R.hindex()
    y1, e1, kv1, tl1 = O.hindex()
    assert np.array_equal(y0, y1) and e0 == e1 and tl0 == tl1
    assert np.array_equal(kv0, kv1)
    if threads == 4:
        for r in reads[:25]:
            if len(r) <= 200:
                continue
            assert np.array_equal(R.stage(r, 1), O.stage(r, 1))           # getHIndexMatchAll incl. head words read as bodies
            assert np.array_equal(R.stage(r, 1, len(r) // 3, len(r) - 100, 1), O.stage(r, 1, len(r) // 3, len(r) - 100, 1))
            assert np.array_equal(R.cords(r), O.cords(r))


# ---- -f 1 (1-mer / 32-base features): canonical rule of oracle/ref_harness.cpp ------------------------------------------
GOLDEN_F1 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_f1.json")))


@pytest.mark.parametrize("name", ["clean_hifi", "repeat_ont", "repeat_t1_p0"])
def test_oracle_f1_matches_golden(name):
    g, reads, bases, offs, T, preset = make_case(name)
    gd = GOLDEN_F1[name]
    O = Oracle(g, threads=T, preset=preset, feature_type=1)
    for i in range(len(g)):
        assert digest(O.genome_features(i)) == gd["genome_features"][i]
    for k, r in enumerate(reads[:16]):
        if len(r) > 200:
            assert digest(np.concatenate([O.read_features(r, 0), O.read_features(r, 1)])) == gd["read_features"][k]
    cords, coff = O.map_batch(bases, offs, map_threads=2)
    assert all(gd["stable"])
    for i, r in enumerate(reads):
        c = cords[int(coff[i]):int(coff[i + 1])]
        assert len(c) == (gd["n_cords"][i] if len(r) > 200 else 0)
        if len(r) > 200:
            assert digest(c) == gd["cords"][i], f"read {i}"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_f1_equals_reference_and_reference_is_stable():
    """with zeroed feature slack the reference's -f 1 output is a function of its input (two passes agree) and the
    oracle restates it: features incl. the zeroed tail entries, hits, cords after the first apxMap_, final cords"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    O = Oracle(g, threads=T, preset=preset, feature_type=1)
    R = RefImpl(g, threads=T, preset=preset, feature_type=1)
    for i in range(len(g)):
        assert np.array_equal(R.genome_features(i), O.genome_features(i))
    for r in reads[:30]:
        if len(r) <= 200:
            continue
        for st in (0, 1):
            assert np.array_equal(R.read_features(r, st), O.read_features(r, st))
        for stage in (3, 4, 0):
            assert np.array_equal(R.stage(r, stage), O.stage(r, stage)), f"stage {stage}"
    a = R.map_batch(bases, offs, map_threads=2)
    b = R.map_batch(bases, offs, map_threads=2)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    oc, oo = O.map_batch(bases, offs, map_threads=2)
    assert np.array_equal(a[1], oo) and np.array_equal(a[0], oc)


# ---- SAM* / BAM* record construction: cords2BamLink f_io.cpp:883 ----------------------------------------
BAM_PARMS = [((1 << 60) - 1, (1 << 60) - 1, 96), (80, 200, 96), (5, 3, 96), (7, 1, 192)]   # defaults; -p 0; two that force the split path


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["clean_hifi", "repeat_ont"])
def test_oracle_cords2bam_equals_reference(name):
    """records (contig, position, flag, score) and cigar* elements of every read, for the default thresholds, the -p 0
    thresholds (thd_DI 80 / thd_X 200, mapper.cpp:185) and small ones that drive cord2cigar_'s split branch (:795-824)"""
    g, reads, bases, offs, T, preset = make_case(name)
    O = Oracle(g, threads=T, preset=preset)
    R = RefImpl(g, threads=T, preset=preset, build_index=False)
    cords, coff = O.map_batch(bases, offs, map_threads=2)
    n_rec = n_split = 0
    for di, x, w in BAM_PARMS:
        for i, r in enumerate(reads):
            c = cords[int(coff[i]):int(coff[i + 1])]
            a, b = O.cords2bam(len(r), c, w, 8000, di, x), R.cords2bam(len(r), c, w, 8000, di, x)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (i, di, x, w)
            n_rec += len(a[0])
            n_split += int(np.count_nonzero((a[1] >> np.uint64(32)) == ord("X")))
    assert n_rec > 100 and n_split > 0


def test_product_record_walk_equals_oracle():
    """lnr_bamrec.h (the walk the CUDA kernels instantiate) compiled for the host: count pass + fill pass == oracle"""
    from cpu_checkers import HostEmu
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    O = Oracle(g, threads=T, preset=preset)
    E = HostEmu(g, threads=T, preset=preset, build_index=False)
    cords, coff = O.map_batch(bases, offs, map_threads=2)
    for di, x, w in BAM_PARMS:
        for i, r in enumerate(reads):
            c = cords[int(coff[i]):int(coff[i + 1])]
            a, b = O.cords2bam(len(r), c, w, 8000, di, x), E.cords2bam(len(r), c, w, 8000, di, x)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (i, di, x, w)


# ---- -c 0 (apxMap with f_chain = 0, alg_type 1: getDAnchorList / getDHitList / path_dst_1) -----------------
GOLDEN_C0 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_c0.json")))


@pytest.mark.parametrize("name", ["clean_hifi", "repeat_ont", "repeat_t1_p0"])
@pytest.mark.parametrize("ft", [2, 1])
def test_oracle_c0_matches_golden(name, ft):
    g, reads, bases, offs, T, preset = make_case(name)
    O = Oracle(g, threads=T, preset=preset, feature_type=ft)
    for st in (0, 1):
        gd = GOLDEN_C0[name][f"f{ft}_s{st}"]
        assert all(gd["stable"])
        cords, coff = O.map_batch(bases, offs, map_threads=2, no_chain=True, gdl_state=st)
        for i in range(len(reads)):
            c = cords[int(coff[i]):int(coff[i + 1])]
            assert len(c) == gd["n_cords"][i] and digest(c) == gd["cords"][i], f"read {i} state {st}"
    # the mode is a different algorithm, not a variant of the default: most reads get other cords than with f_chain = 1
    c1, o1 = O.map_batch(bases, offs, map_threads=2)
    assert not (np.array_equal(o1, coff) and np.array_equal(c1, cords))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_c0_equals_reference():
    """the unmodified reference's apxMap with f_chain = 0 (fresh PMPParms per read, both GetDHitListParms states) against the
    restatement, incl. -i 2 seeding"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    for it, ft in ((1, 2), (2, 2), (1, 1)):
        O = Oracle(g, threads=T, preset=preset, feature_type=ft, index_type=it)
        R = RefImpl(g, threads=T, preset=preset, feature_type=ft, index_type=it)
        for st in (0, 1):
            oc, oo = O.map_batch(bases, offs, map_threads=2, no_chain=True, gdl_state=st)
            rc, ro = R.map_batch(bases, offs, map_threads=2, no_chain=True, gdl_state=st)
            assert np.array_equal(oo, ro) and np.array_equal(oc, rc), (it, ft, st)
            # -i 2 maps nothing in this mode: getHIndexMatchAll reads its record window from map_end's x field, and
            # apxMap passes the bare read length there (pmpfinder.cpp:1933, :2778)
            assert int(np.count_nonzero(np.diff(oo.astype(np.int64)))) > len(reads) // 2 if it == 1 else len(oc) == 0
