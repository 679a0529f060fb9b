"""CPU: the product's host logic. (1) the C-ABI library loads and exports every symbol include/lnr_b200.h declares,
(2) the pipeline headers (linear_b200/csrc/lnr_*.h), compiled for the host with a single-lane warp, agree with the
oracle, (3) lnr::gnu_sort reproduces libstdc++ std::sort's permutation with comparator ties."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cases import make_case
from cpu_checkers import ROOT, HostEmu, Oracle, build_emu

import linear_b200 as lb
from linear_b200 import api


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lnr_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(lnr_[a-z_0-9]+)\s*\(", hdr)))
    assert sorted(api.EXPORTS) == declared
    lib = lb.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(lb.LnrError):
        lb.Context(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "linear_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower().replace("the oracle", "").replace("against the oracle", ""), f


@pytest.mark.parametrize("name", ["clean_hifi", "repeat_ont", "repeat_t1_p0"])
def test_host_emulation_of_pipeline_matches_oracle(name):
    g, reads, bases, offs, T, preset = make_case(name)
    O = Oracle(g, threads=T, preset=preset)
    E = HostEmu(g, threads=T, preset=preset)
    d0, h0 = O.dindex()
    d1, h1 = E.dindex()
    assert np.array_equal(d0, d1) and np.array_equal(h0, h1)
    for i in range(len(g)):
        assert np.array_equal(O.genome_features(i), E.genome_features(i))
    sel = reads if name != "repeat_ont" else reads[:50]
    for i, r in enumerate(sel):
        if len(r) <= 200:
            continue
        for st in (0, 1):
            assert np.array_equal(O.read_features(r, st), E.read_features(r, st))
        assert np.array_equal(O.stage(r, 1), E.stage(r, 1)), f"anchors {i}"
        assert np.array_equal(O.stage(r, 1, len(r) // 3, len(r) - 100, 1), E.stage(r, 1, len(r) // 3, len(r) - 100, 1))
        assert np.array_equal(O.stage(r, 3), E.stage(r, 3)), f"hits {i}"
        assert np.array_equal(O.stage(r, 4), E.stage(r, 4)), f"cords1 {i}"
        assert np.array_equal(O.cords(r), E.cords(r)), f"cords {i}"


def test_gnu_sort_matches_std_sort_with_ties():
    lib = C.CDLL(build_emu())
    lib.emu_gnu_sort_check.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.c_int]
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 15, 16, 17, 33, 100, 1000, 5000, 40000):
        for nkeys in (1, 2, 5, 50, 10 ** 9):
            for desc in (0, 1):
                keys = rng.integers(0, nkeys, size=n, dtype=np.uint64)
                a = (keys << np.uint64(32)) | np.arange(n, dtype=np.uint64)
                assert lib.emu_gnu_sort_check(a.ctypes.data_as(C.POINTER(C.c_uint64)), n, desc) == 0, (n, nkeys, desc)
    # adversarial for quicksort: organ-pipe and sorted inputs push introsort into its heap-sort fallback
    for n in (1000, 20000):
        for arr in (np.arange(n), np.arange(n)[::-1], np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]])):
            a = (arr.astype(np.uint64) // np.uint64(3) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
            assert lib.emu_gnu_sort_check(a.ctypes.data_as(C.POINTER(C.c_uint64)), n, 0) == 0


def test_warp_cooperative_gnu_sort_matches_std_sort_with_ties():
    """gnu_sort_w (parallel partitions + stable radix pass, lnr_pipeline.h) == std::sort, one-lane emulation"""
    lib = C.CDLL(build_emu())
    lib.emu_gnu_sort_w_check.argtypes = [C.POINTER(C.c_uint64), C.c_int]
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 15, 16, 17, 18, 33, 64, 100, 1000, 5000, 40000):
        for nkeys in (1, 2, 5, 50, 1000, 2 ** 30):
            keys = rng.integers(0, nkeys, size=n, dtype=np.uint64)
            a = (keys << np.uint64(32)) | np.arange(n, dtype=np.uint64)
            assert lib.emu_gnu_sort_w_check(a.ctypes.data_as(C.POINTER(C.c_uint64)), n) == 0, (n, nkeys)
    for n in (1000, 20000):
        for arr in (np.arange(n), np.arange(n)[::-1], np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]])):
            a = (arr.astype(np.uint64) // np.uint64(3) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
            assert lib.emu_gnu_sort_w_check(a.ctypes.data_as(C.POINTER(C.c_uint64)), n) == 0


@pytest.mark.parametrize("seed", [0, 3, 5, 9, 12])
def test_host_emulation_on_random_cases(seed):
    """the randomised cases of tests/test_gpu_fuzz.py (N runs, repeat density, -t, preset all vary) through the product
    headers on the host: the same sweep the GPU runs, checkable without one"""
    from test_gpu_fuzz import make_fuzz_case
    g, bases, offs, T, preset = make_fuzz_case(seed)
    O = Oracle(g, threads=T, preset=preset)
    E = HostEmu(g, threads=T, preset=preset)
    d0, h0 = O.dindex()
    d1, h1 = E.dindex()
    assert np.array_equal(d0, d1) and np.array_equal(h0, h1)
    for r in range(0, len(offs) - 1, 3):
        read = bases[int(offs[r]):int(offs[r + 1])]
        if len(read) <= 200:
            continue
        assert np.array_equal(O.stage(read, 1), E.stage(read, 1)), (seed, "anchors", r)
        assert np.array_equal(O.stage(read, 3), E.stage(read, 3)), (seed, "hits", r)
        assert np.array_equal(O.cords(read), E.cords(read)), (seed, "cords", r)


@pytest.mark.parametrize("name", ["clean_hifi", "repeat_ont"])
def test_host_emulation_of_c0_path_matches_oracle(name):
    """-c 0: the product's phase_c0 / path_dst_1 / c0_finish (lnr_pipeline.h) with a one-lane warp against the oracle"""
    g, reads, bases, offs, T, preset = make_case(name)
    for ft in (2, 1):
        O = Oracle(g, threads=T, preset=preset, feature_type=ft)
        E = HostEmu(g, threads=T, preset=preset, feature_type=ft)
        for st in (0, 1):
            oc, oo = O.map_batch(bases, offs, map_threads=2, no_chain=True, gdl_state=st)
            ec, eo = E.map_batch(bases, offs, no_chain=True, gdl_state=st)
            assert np.array_equal(oo, eo) and np.array_equal(oc, ec), (ft, st)
