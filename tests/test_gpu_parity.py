"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle (and, where oracle/_ref travelled
with the snapshot, the unmodified reference) and against the committed golden digests. Bit-exact everywhere."""
import json
import os

import numpy as np
import pytest

from cases import CASES, digest, make_case
from cpu_checkers import Oracle, RefImpl, have_ref

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


@pytest.fixture(scope="module")
def lb():
    import linear_b200
    linear_b200.load_library()   # fails loudly if the CUDA extension is missing
    return linear_b200


@pytest.fixture(scope="module")
def ctx(lb):
    return lb.Context(0)


@pytest.fixture(scope="module", params=list(CASES))
def case(request, lb, ctx):
    g, reads, bases, offs, T, preset = make_case(request.param)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    O = Oracle(g, threads=T, preset=preset)
    return dict(name=request.param, g=g, reads=reads, bases=bases, offs=offs, T=T, preset=preset, gen=gen, feats=feats,
                index=index, O=O)


def test_dindex_bit_exact(case):
    d0, h0 = case["O"].dindex()
    d1, h1 = case["index"].export_dindex()
    assert d1.shape == d0.shape and np.array_equal(d0, d1)
    assert np.array_equal(h0, h1)
    gd = GOLDEN[case["name"]]
    assert len(h1) == gd["n_hs"] and digest(h1) == gd["hs"]


def test_index_properties(case):
    """size-independent properties: buckets ascending, <= 400 entries, every record decodes to a sampled position"""
    d, hs = case["index"].export_dindex()
    cnt = np.diff(d)
    assert cnt.min() >= 0 and cnt.max() <= 400 and d[-1] == len(hs)
    same_bucket = np.ones(len(hs) - 1, dtype=bool)
    same_bucket[d[1:-1][(d[1:-1] > 0) & (d[1:-1] < len(hs))] - 1] = False
    assert np.all(hs[1:][same_bucket] > hs[:-1][same_bucket])
    x = (hs >> np.uint64(20)) & np.uint64((1 << 30) - 1)
    cid = (hs >> np.uint64(50)) & np.uint64(1023)
    lens = np.array([len(c) for c in case["g"]], dtype=np.uint64)
    assert np.all(cid < len(lens)) and np.all(x - np.uint64(1 << 20) < lens[cid.astype(np.int64)])


def test_genome_features_bit_exact(case):
    for i in range(len(case["g"])):
        assert np.array_equal(case["O"].genome_features(i), case["feats"].download(i))
        assert digest(case["feats"].download(i)[:-1]) == GOLDEN[case["name"]]["genome_features"][i]


def test_read_features_bit_exact(case, lb, ctx):
    for r in case["reads"][:12]:
        if len(r) <= 200:
            continue
        f, rc = lb.read_features(ctx, r)
        assert np.array_equal(f, case["O"].read_features(r, 0))
        assert np.array_equal(rc, case["O"].read_features(r, 1))


def test_apxmap_stages_and_cords_bit_exact(case, lb, ctx):
    O, reads = case["O"], case["reads"]
    cords, coff, dbg = lb.apx_map_batch(ctx, case["index"], case["feats"], case["bases"], case["offs"], preset=case["preset"], debug=True)
    gd = GOLDEN[case["name"]]
    bad = []
    for i, r in enumerate(reads):
        mine = cords[int(coff[i]):int(coff[i + 1])]
        if len(r) <= 200:
            assert len(mine) == 0
            continue
        ra = dbg["ra"][int(dbg["ra_off"][i]):int(dbg["ra_off"][i + 1])]
        assert np.array_equal(ra, O.stage(r, 1)[1:]), f"raw anchors, read {i}"
        h = dbg["h"][int(dbg["h_off"][i]):int(dbg["h_off"][i + 1])]
        assert np.array_equal(h, O.stage(r, 3)), f"hits, read {i}"
        c1 = dbg["c1"][int(dbg["c1_off"][i]):int(dbg["c1_off"][i + 1])]
        assert np.array_equal(c1, O.stage(r, 4)), f"cords after first apxMap_, read {i}"
        if not np.array_equal(mine, O.cords(r)):
            bad.append(i)
        if gd["stable"][i]:
            assert digest(mine) == gd["cords"][i], f"golden cords, read {i}"
    assert not bad, f"cords differ for reads {bad}"
    # cords_end reconstruction (pmpfinder.cpp:2790-2801)
    assert np.array_equal(lb.cords_end(cords), cords + np.uint64((96 << 20) | 96))


def test_batch_without_debug_and_repeatability(case, lb, ctx):
    a, ao = lb.apx_map_batch(ctx, case["index"], case["feats"], case["bases"], case["offs"], preset=case["preset"])
    b, bo = lb.apx_map_batch(ctx, case["index"], case["feats"], case["bases"], case["offs"], preset=case["preset"])
    assert np.array_equal(ao, bo) and np.array_equal(a, b)
    oc, oo = case["O"].map_batch(case["bases"], case["offs"], map_threads=4)
    assert np.array_equal(oo, ao) and np.array_equal(oc, a)


def test_empty_and_tiny_batches(case, lb, ctx):
    c, off = lb.apx_map_batch(ctx, case["index"], case["feats"], np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(c) == 0 and list(off) == [0]
    r = case["reads"][0]
    c, off = lb.apx_map_batch(ctx, case["index"], case["feats"], r, np.array([0, len(r)], np.uint64), preset=case["preset"])
    assert np.array_equal(c, case["O"].cords(r))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref did not travel")
def test_against_unmodified_reference(lb, ctx):
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    R = RefImpl(g, threads=T, preset=preset)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    d0, h0 = R.dindex()
    d1, h1 = index.export_dindex()
    assert np.array_equal(d0, d1) and np.array_equal(h0, h1)
    cords, coff = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    rc, ro = R.map_batch(bases, offs, map_threads=4)
    assert np.array_equal(ro, coff) and np.array_equal(rc, cords)


def test_genome_with_N_runs(lb, ctx):
    """N arithmetic of the rolling hash (shape_extend.cpp:176-180) and of the features, closed-form on the GPU"""
    from linear_b200 import datagen
    lens = datagen.contig_lengths(600_000, 2, seed=4)
    g = datagen.make_genome(21, lens, n_families=1, copies=30, n_tandem=4, n_runs=40)
    rs = datagen.simulate_reads(5, g, 30, mean_len=5000, sd_len=1000, err=0.03)
    O = Oracle(g, threads=4)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, 4)
    index = lb.create_index(ctx, gen, 1, 4)
    d0, h0 = O.dindex()
    d1, h1 = index.export_dindex()
    assert np.array_equal(d0, d1) and np.array_equal(h0, h1)
    for i in range(len(g)):
        assert np.array_equal(O.genome_features(i), feats.download(i))
    cords, coff = lb.apx_map_batch(ctx, index, feats, rs.bases, rs.offsets)
    oc, oo = O.map_batch(rs.bases, rs.offsets, map_threads=4)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords)


def test_limits_are_rejected(lb, ctx):
    with pytest.raises(lb.LnrError):
        lb.Genome(ctx, [np.zeros(10, np.uint8)] * 1025)


@pytest.mark.parametrize("n_shards", [2, 8])
def test_hash_range_sharded_index_build(lb, ctx, n_shards):
    """multi-GPU index build, emulated on one GPU: every shard builds the buckets of its minimizer range; concatenating
    the shards in order and rebasing dir reproduces the whole DIndex bit for bit (the NCCL all-gather only moves these
    arrays between ranks; bench.py --gpus N runs it for real)"""
    from linear_b200 import sharding
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    gen = lb.Genome(ctx, g)
    full = lb.create_index(ctx, gen, 1, T)
    d0, h0 = full.export_dindex()
    parts = []
    for s in range(n_shards):
        p = lb.Index(ctx, gen, 1, T, shard=s, n_shards=n_shards)
        parts.append(p.export_dindex())
        p.close()
    assert sum(len(p[1]) for p in parts) == len(h0)
    d1, h1 = sharding.assemble_dindex(parts)
    assert np.array_equal(d0, d1) and np.array_equal(h0, h1)
    # device-side round trip used by the NCCL path
    import torch
    dev = torch.device("cuda", 0)
    dt, ht = full.export_device(torch, dev)
    again = lb.Index.from_device(ctx, dt, ht)
    d2, h2 = again.export_dindex()
    assert np.array_equal(d0, d2) and np.array_equal(h0, h2)


def test_scratch_overflow_falls_back_to_big_arena(lb, monkeypatch):
    """reads whose scratch does not fit the per-warp arena are re-run by the big-arena pass; results are unchanged"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    monkeypatch.setenv("LNR_ARENA_KB", "16")
    small = lb.Context(0)
    gen = lb.Genome(small, g)
    feats = lb.create_features(small, gen, 2, T)
    index = lb.create_index(small, gen, 1, T)
    cords, coff = lb.apx_map_batch(small, index, feats, bases, offs, preset=preset)
    d = small.diag()
    assert d["heavy_lane_tasks"] > 0 and d["finish_big_reads"] > 0, d   # the big-arena warps / passes really took tasks
    # the same with the heavy lane switched off: the tasks go through the separate big-arena launch instead
    monkeypatch.setenv("LNR_NO_HEAVY_LANE", "1")
    c2, o2 = lb.apx_map_batch(small, index, feats, bases, offs, preset=preset)
    assert small.diag()["hits_big_tasks"] > 0 and small.diag()["heavy_lane_tasks"] == 0
    assert np.array_equal(o2, coff) and np.array_equal(c2, cords)
    oc, oo = Oracle(g, threads=T, preset=preset).map_batch(bases, offs, map_threads=4)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords)


def test_one_kernel_hit_stage_gives_the_same_cords(lb, ctx, monkeypatch):
    """the primary pass runs the hit stage as three kernels (sort / chain / blocks); the single-kernel form of the same
    sections (used by the re-map and big-arena passes) must give identical stages and cords"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    ctx.set_profiling(True)
    ctx.reset_kernel_times()
    c3, o3, d3 = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset, debug=True)
    kt3 = ctx.kernel_times()
    assert "k_hits_sort" in kt3 and "k_hits_chain" in kt3 and "k_hits_blocks" in kt3 and "k_map_hits" not in kt3, sorted(kt3)
    ctx.reset_kernel_times()
    monkeypatch.setenv("LNR_MONOLITHIC_HITS", "1")
    c1, o1, d1 = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset, debug=True)
    kt1 = ctx.kernel_times()
    ctx.set_profiling(False)
    assert "k_map_hits" in kt1 and "k_hits_sort" not in kt1, sorted(kt1)
    assert np.array_equal(o1, o3) and np.array_equal(c1, c3)
    for k in d3:
        assert np.array_equal(np.asarray(d1[k]), np.asarray(d3[k])), k


@pytest.mark.parametrize("threads", [1, 4, 8])
def test_hindex_build_bit_exact(lb, ctx, threads):
    """-i 2: ysa byte-exact (heads, descending bodies, zeroed Y of small blocks, chunk-tail mislabel), emptyDir, table
    length and the directory as a sorted key -> value list"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    O = Oracle(g, threads=threads, preset=preset, index_type=2)
    y0, e0, kv0, tl0 = O.hindex()
    gen = lb.Genome(ctx, g)
    index = lb.create_index(ctx, gen, 2, threads)
    y1, e1, kv1, tl1 = index.export_hindex()
    assert len(y0) == len(y1) and e0 == e1 and tl0 == tl1
    assert np.array_equal(y0, y1)
    assert np.array_equal(kv0, kv1)


def test_hindex_apxmap_bit_exact(lb, ctx):
    """-i 2 end to end: getHIndexMatchAll anchors (incl. head words consumed as bodies) and final cords"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    O = Oracle(g, threads=T, preset=preset, index_type=2)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 2, T)
    cords, coff, dbg = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset, debug=True)
    for i, r in enumerate(reads):
        if len(r) <= 200:
            continue
        ra = dbg["ra"][int(dbg["ra_off"][i]):int(dbg["ra_off"][i + 1])]
        assert np.array_equal(ra, O.stage(r, 1)[1:]), f"raw anchors, read {i}"
    oc, oo = O.map_batch(bases, offs, map_threads=4)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords)
    if have_ref():
        R = RefImpl(g, threads=T, preset=preset, index_type=2)
        rc, ro = R.map_batch(bases, offs, map_threads=4)
        assert np.array_equal(ro, coff) and np.array_equal(rc, cords)


def test_warp_sort_reproduces_std_sort_tie_order(lb, ctx):
    """chainAnchorsHits sorts with a key-only comparator (pmpfinder.cpp:2465); tied anchors must come out in libstdc++'s
    introsort order. The 32-lane gnu_sort_w is checked against std::sort itself on tie-heavy and adversarial inputs."""
    import ctypes as C
    from cpu_checkers import build_emu
    emu = C.CDLL(build_emu())
    emu.emu_std_sort_hi.argtypes = [C.POINTER(C.c_uint64), C.c_int]
    rng = np.random.default_rng(7)
    inputs = []
    for n in (1, 2, 16, 17, 31, 32, 33, 64, 65, 100, 1000, 5000, 40000):
        for nkeys in (1, 2, 5, 50, 1000, 2 ** 30):
            inputs.append(rng.integers(0, nkeys, size=n, dtype=np.uint64))
    for n in (1000, 20000):
        inputs += [np.arange(n, dtype=np.uint64) // 3, np.arange(n, dtype=np.uint64)[::-1] // 3,
                   np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]).astype(np.uint64) // 3]
    for keys in inputs:
        a = (keys << np.uint64(32)) | np.arange(len(keys), dtype=np.uint64)
        want = a.copy()
        emu.emu_std_sort_hi(want.ctypes.data_as(C.POINTER(C.c_uint64)), len(want))
        got = lb.selftest_sort(ctx, a)
        assert np.array_equal(got, want), (len(keys), int(keys.max()))


def test_concurrent_contexts_share_index_and_features(lb, ctx):
    """p_calRecords is called from -t host threads: several contexts (own stream + workspace each) map at the same time
    against ONE genome / feature / index object; every thread must get exactly the single-threaded result."""
    import threading
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    want_c, want_o = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    ctxs = [lb.Context(0) for _ in range(4)]
    out, errs = [None] * 4, []

    def work(k):
        try:
            for _ in range(3):
                out[k] = lb.apx_map_batch(ctxs[k], index, feats, bases, offs, preset=preset)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for k in range(4):
        assert np.array_equal(out[k][1], want_o) and np.array_equal(out[k][0], want_c), k


def test_exhausted_match_mask_pool_falls_back_to_rescan(lb, monkeypatch):
    """k_seed_count records matches as bitmasks from a pooled allocator; a sample that gets no words is re-scanned by
    k_seed_fill with the Y-key rule itself. With one word per pool nearly every sample takes that path."""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    monkeypatch.setenv("LNR_MASK_WORDS", "256")
    c = lb.Context(0)
    gen = lb.Genome(c, g)
    feats = lb.create_features(c, gen, 2, T)
    index = lb.create_index(c, gen, 1, T)
    cords, coff = lb.apx_map_batch(c, index, feats, bases, offs, preset=preset)
    assert c.diag()["seed_rescans"] > 0          # the re-scan branch of the fill pass really ran
    oc, oo = Oracle(g, threads=T, preset=preset).map_batch(bases, offs, map_threads=4)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords)


@pytest.mark.parametrize("with_n", [False, True])
def test_packed_reads_give_identical_cords(lb, ctx, with_n):
    """lnr_apxmap_batch_packed (2-bit bases + optional N bitmap, a quarter of the PCIe bytes) == lnr_apxmap_batch on the
    Dna5 string == the oracle; with N runs inside reads, at read edges and at positions that are not multiples of 4/16"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    bases = bases.copy()
    if with_n:
        rng = np.random.default_rng(5)
        for r in range(0, len(offs) - 1, 3):
            a, b = int(offs[r]), int(offs[r + 1])
            if b - a < 400:
                continue
            for _ in range(3):
                p = int(rng.integers(a, b - 40))
                bases[p:p + int(rng.integers(1, 37))] = 4
        bases[int(offs[1]) - 1] = 4          # last base of a read
        bases[int(offs[2])] = 4              # first base of a read
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    packed, nmask = lb.pack_dna5(bases)
    assert (nmask is not None) == with_n
    assert len(packed) == (len(bases) + 3) // 4
    c1, o1 = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    c2, o2 = lb.apx_map_batch_packed(ctx, index, feats, packed, nmask, offs, preset=preset)
    assert np.array_equal(o1, o2) and np.array_equal(c1, c2)
    oc, oo = Oracle(g, threads=T, preset=preset).map_batch(bases, offs, map_threads=4)
    assert np.array_equal(oo, o2) and np.array_equal(oc, c2)


# ---- -f 1: 1-mer / 32-base features, window 192 (createFeatures1_32, __scriptDist16_3; BASELINE configs[3]) -----------
GOLDEN_F1 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_f1.json")))


@pytest.mark.parametrize("name", list(CASES))
def test_f1_features_and_cords_bit_exact(lb, ctx, name):
    """-f 1 under the canonical rule (unwritten / out-of-range feature entries are 0): genome features, read features and
    final cords equal to the oracle, to the golden digests made from the reference, and to the reference itself"""
    g, reads, bases, offs, T, preset = make_case(name)
    gd = GOLDEN_F1[name]
    O = Oracle(g, threads=T, preset=preset, feature_type=1)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 1, T)
    index = lb.create_index(ctx, gen, 1, T)
    for i in range(len(g)):
        f = feats.download(i)
        assert f.dtype == np.int16
        assert np.array_equal(f.astype(np.int32), O.genome_features(i))
        assert digest(f.astype(np.int32)) == gd["genome_features"][i]
    for k, r in enumerate(reads[:16]):
        if len(r) <= 200:
            continue
        f, rc = lb.read_features(ctx, r, 1)
        assert np.array_equal(f.astype(np.int32), O.read_features(r, 0)) and np.array_equal(rc.astype(np.int32), O.read_features(r, 1))
        assert digest(np.concatenate([f.astype(np.int32), rc.astype(np.int32)])) == gd["read_features"][k]
    cords, coff = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    oc, oo = O.map_batch(bases, offs, map_threads=4)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords)
    n_checked = 0
    for i, r in enumerate(reads):
        if not gd["stable"][i]:
            continue
        c = cords[int(coff[i]):int(coff[i + 1])]
        assert len(c) == (gd["n_cords"][i] if len(r) > 200 else 0) and (len(r) <= 200 or digest(c) == gd["cords"][i]), f"read {i}"
        n_checked += 1
    assert n_checked == len(reads)            # no read had to be excluded as oracle-unstable
    if have_ref():
        rc_, ro_ = RefImpl(g, threads=T, preset=preset, feature_type=1).map_batch(bases, offs, map_threads=4)
        assert np.array_equal(ro_, coff) and np.array_equal(rc_, cords)


def test_f1_with_hindex_and_mismatched_params(lb, ctx):
    """-f 1 together with -i 2, and the argument check: lnr_params.feature_type must agree with the genome features"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 1, T)
    index = lb.create_index(ctx, gen, 2, T)
    cords, coff = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    oc, oo = Oracle(g, threads=T, preset=preset, feature_type=1, index_type=2).map_batch(bases, offs, map_threads=4)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords)
    import ctypes as C
    from linear_b200.api import Params, u64p
    prm = Params(preset=preset, feature_type=2)
    out = np.zeros(1 << 16, np.uint64)
    co = np.zeros(len(offs), np.uint64)
    rc = ctx.lib.lnr_apxmap_batch(ctx.h, index.h, feats.h, C.byref(prm), len(offs) - 1, bases.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(u64p),
                                  out.ctypes.data_as(C.c_void_p), co.ctypes.data_as(u64p), len(out), None)
    assert rc == lb.api.LNR_E_ARG


# ---- SAM* / BAM* records from cords (lnr_cords_to_records; cords2BamLink f_io.cpp:883) ---------------------------------
@pytest.mark.parametrize("name", ["clean_hifi", "repeat_ont"])
def test_cords_to_records_bit_exact(lb, ctx, name):
    g, reads, bases, offs, T, preset = make_case(name)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    cords, coff = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    read_len = np.diff(offs).astype(np.uint64)
    O = Oracle(g, threads=T, preset=preset, build_index=False)
    R = RefImpl(g, threads=T, preset=preset, build_index=False) if have_ref() else None
    for di, x, w in [((1 << 60) - 1, (1 << 60) - 1, 96), (80, 200, 96), (5, 3, 96), (7, 1, 192)]:
        recs, roff, cig, goff = lb.cords_to_records(ctx, cords, coff, read_len, window=w, thd_di=di, thd_x=x)
        assert len(recs) == int(roff[-1]) and len(cig) == int(goff[-1]) and len(recs) > 40
        for i, r in enumerate(reads):
            c = cords[int(coff[i]):int(coff[i + 1])]
            want = O.cords2bam(len(r), c, w, 8000, di, x)
            if R is not None:
                wr = R.cords2bam(len(r), c, w, 8000, di, x)
                assert np.array_equal(want[0], wr[0]) and np.array_equal(want[1], wr[1])
            got = recs[int(roff[i]):int(roff[i + 1])]
            assert len(got) == len(want[0]), (i, di)
            for k in range(len(got)):
                assert [int(got[k][f]) for f in ("rid", "begin_pos", "flag", "s1", "s2", "s3", "cigar_begin", "cigar_end")] == [int(v) for v in want[0][k]], (i, k)
            assert np.array_equal(cig[int(goff[i]):int(goff[i + 1])], want[1]), (i, di)
    # a buffer that is too small is refused and the needed sizes are reported
    import ctypes as C
    from linear_b200.api import BamParms, u64p
    prm = BamParms(window=96, thd_large_x=8000, thd_di=(1 << 60) - 1, thd_x=(1 << 60) - 1)
    ro, go = np.zeros(len(coff), np.uint64), np.zeros(len(coff), np.uint64)
    small = np.zeros(4, lb.api.BAM_REC_DTYPE)
    rc = ctx.lib.lnr_cords_to_records(ctx.h, len(coff) - 1, C.c_void_p(cords.ctypes.data), coff.ctypes.data_as(u64p), read_len.ctypes.data_as(u64p), C.byref(prm),
                                      C.c_void_p(small.ctypes.data), 4, ro.ctypes.data_as(u64p), None, 0, go.ctypes.data_as(u64p))
    assert rc == lb.api.LNR_E_CAPACITY and int(ro[-1]) > 40 and int(go[-1]) > 1000


# ---- -c 0 (lnr_params.no_chain; apxMap with f_chain = 0, pmpfinder.cpp:2773-2787) ---------------------------
GOLDEN_C0 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_c0.json")))


@pytest.mark.parametrize("name", ["clean_hifi", "repeat_ont", "repeat_t1_p0"])
@pytest.mark.parametrize("ft", [2, 1])
def test_c0_cords_bit_exact(lb, ctx, name, ft):
    """k_map_c0 (sorted anchor runs -> hits -> path_dst_1, second attempt at step 7) against the oracle, the golden digests
    made from the reference and the reference itself, for both GetDHitListParms states"""
    g, reads, bases, offs, T, preset = make_case(name)
    O = Oracle(g, threads=T, preset=preset, feature_type=ft)
    R = RefImpl(g, threads=T, preset=preset, feature_type=ft) if have_ref() else None
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, ft, T)
    index = lb.create_index(ctx, gen, 1, T)
    for st in (0, 1):
        gd = GOLDEN_C0[name][f"f{ft}_s{st}"]
        cords, coff = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset, no_chain=True, gdl_state=st)
        oc, oo = O.map_batch(bases, offs, map_threads=4, no_chain=True, gdl_state=st)
        bad = [i for i in range(len(reads)) if not np.array_equal(cords[int(coff[i]):int(coff[i + 1])], oc[int(oo[i]):int(oo[i + 1])])]
        assert not bad, (st, bad[:8])
        assert np.array_equal(oo, coff) and np.array_equal(oc, cords)
        for i in range(len(reads)):
            c = cords[int(coff[i]):int(coff[i + 1])]
            assert len(c) == gd["n_cords"][i] and digest(c) == gd["cords"][i], f"read {i} state {st}"
        if R is not None:
            rc_, ro_ = R.map_batch(bases, offs, map_threads=4, no_chain=True, gdl_state=st)
            assert np.array_equal(ro_, coff) and np.array_equal(rc_, cords)


def test_c0_second_attempt_big_arena_and_hindex(lb, monkeypatch):
    """-c 0: the junk / chimeric reads of the case take the second attempt (step 7); with a 48 KB per-warp arena most reads go
    through the big-arena launch and the cords stay the same; with -i 2 the mode maps nothing, as in the reference"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    oc, oo = Oracle(g, threads=T, preset=preset).map_batch(bases, offs, map_threads=4, no_chain=True)
    ctx = lb.Context(0)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    cords, coff = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset, no_chain=True)
    assert np.array_equal(oo, coff) and np.array_equal(oc, cords)
    assert ctx.counters()["remap_tasks"] > 0               # reads that needed the second attempt
    hindex = lb.create_index(ctx, gen, 2, T)
    ch, oh = lb.apx_map_batch(ctx, hindex, feats, bases, offs, preset=preset, no_chain=True)
    ohc, oho = Oracle(g, threads=T, preset=preset, index_type=2).map_batch(bases, offs, map_threads=4, no_chain=True)
    assert len(ch) == 0 and len(ohc) == 0 and np.array_equal(oh, oho)
    monkeypatch.setenv("LNR_ARENA_KB", "48")
    ctx2 = lb.Context(0)
    gen2 = lb.Genome(ctx2, g)
    feats2 = lb.create_features(ctx2, gen2, 2, T)
    index2 = lb.create_index(ctx2, gen2, 1, T)
    c2, o2 = lb.apx_map_batch(ctx2, index2, feats2, bases, offs, preset=preset, no_chain=True)
    assert np.array_equal(oo, o2) and np.array_equal(oc, c2)
    assert ctx2.diag()["hits_big_tasks"] > 0               # tasks taken by the big-arena launch


# ---- index serialisation (lnr_index_save / lnr_index_load) -------------------------------------------------
@pytest.mark.parametrize("index_type", [1, 2])
def test_index_save_load_round_trip(lb, ctx, tmp_path, index_type):
    """an index written by lnr_index_save and read back by lnr_index_load (into another context) has the same arrays and maps
    the reads to the same cords; a truncated or altered file is refused"""
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, index_type, T)
    path = str(tmp_path / "genome.lnridx")
    index.save(path)
    ctx2 = lb.Context(0)
    back = lb.Index.load(ctx2, path)
    assert back.index_type == index_type
    if index_type == 1:
        d0, h0 = index.export_dindex()
        d1, h1 = back.export_dindex()
        assert np.array_equal(d0, d1) and np.array_equal(h0, h1)
        assert os.path.getsize(path) == 64 + d0.nbytes + h0.nbytes
    else:
        y0, e0, kv0, t0 = index.export_hindex()
        y1, e1, kv1, t1 = back.export_hindex()
        assert np.array_equal(y0, y1) and e0 == e1 and np.array_equal(kv0, kv1) and t0 == t1
    gen2 = lb.Genome(ctx2, g)
    feats2 = lb.create_features(ctx2, gen2, 2, T)
    c0, o0 = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    c1, o1 = lb.apx_map_batch(ctx2, back, feats2, bases, offs, preset=preset)
    assert np.array_equal(o0, o1) and np.array_equal(c0, c1) and len(c0) > 1000
    raw = bytearray(open(path, "rb").read())
    bad = str(tmp_path / "bad.lnridx")
    open(bad, "wb").write(raw[: len(raw) - 4096])
    with pytest.raises(lb.api.LnrError):
        lb.Index.load(ctx2, bad)
    raw[len(raw) // 2] ^= 0x40
    open(bad, "wb").write(raw)
    with pytest.raises(lb.api.LnrError):
        lb.Index.load(ctx2, bad)
    open(bad, "wb").write(b">chr1\nACGT\n" * 100)
    with pytest.raises(lb.api.LnrError):
        lb.Index.load(ctx2, bad)
