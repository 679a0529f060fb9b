"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libref_harness.so, built from
/root/reference by oracle/build_ref.sh). Run in the build container only:  python tests/golden/make_golden.py [base|f1|c0]  (no argument: all three files)
The reference ships no golden vectors of its own; these digests pin the oracle
restatement and the CUDA path to the reference's own outputs on the seeded cases of tests/cases.py."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
from cases import CASES, digest, make_case  # noqa: E402
from cpu_checkers import RefImpl  # noqa: E402

ONLY = sys.argv[1] if len(sys.argv) > 1 else "all"

if ONLY in ("all", "base"):
    out = {}
    for name in CASES:
        g, reads, bases, offs, T, preset = make_case(name)
        R = RefImpl(g, threads=T, preset=preset)
        d, hs = R.dindex()
        nz = np.flatnonzero(np.diff(d))
        e = {"threads": T, "preset": preset, "n_hs": int(len(hs)), "hs": digest(hs),
             "dir_nonzero": digest(nz.astype(np.int64)), "dir_counts": digest(np.diff(d)[nz].astype(np.int32)),
             "genome_features": [digest(R.genome_features(i)[:-1]) for i in range(len(g))]}
        # two passes of the reference over every read; reads whose output is not reproducible are excluded
        # (the reference reads unowned bytes in rare cases)
        empty = np.zeros(0, np.uint64)   # reads <= 200 bases are never mapped (mapper.cpp:440)
        cords_a = [R.cords(r) if len(r) > 200 else empty for r in reads]
        cords_b = [R.cords(r) if len(r) > 200 else empty for r in reads]
        stable = [bool(np.array_equal(a, b)) for a, b in zip(cords_a, cords_b)]
        e["stable"] = stable
        e["n_cords"] = [int(len(c)) for c in cords_a]
        e["cords"] = [digest(c) for c in cords_a]
        e["raw_anchors"] = [digest(R.stage(r, 1)[1:]) if len(r) > 200 else digest(np.zeros(0, np.uint64)) for r in reads]
        e["hits"] = [digest(R.stage(r, 3)) if len(r) > 200 else "" for r in reads]
        out[name] = e
        print(name, "reads", len(reads), "unstable", stable.count(False), "cords", sum(e["n_cords"]))
    json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=0)

if ONLY in ("all", "f1"):
    # -f 1 (1-mer / 32-base features): the reference under the canonical rule of oracle/ref_harness.cpp (unwritten and
    # out-of-range feature entries are 0); reads whose cords differ between two passes would be excluded and counted
    out1 = {}
    for name in CASES:
        g, reads, bases, offs, T, preset = make_case(name)
        R = RefImpl(g, threads=T, preset=preset, feature_type=1)
        e = {"threads": T, "preset": preset, "genome_features": [digest(R.genome_features(i)) for i in range(len(g))]}
        empty = np.zeros(0, np.uint64)
        cords_a = [R.cords(r) if len(r) > 200 else empty for r in reads]
        cords_b = [R.cords(r) if len(r) > 200 else empty for r in reads]
        e["stable"] = [bool(np.array_equal(a, b)) for a, b in zip(cords_a, cords_b)]
        e["n_cords"] = [int(len(c)) for c in cords_a]
        e["cords"] = [digest(c) for c in cords_a]
        e["read_features"] = [digest(np.concatenate([R.read_features(r, 0), R.read_features(r, 1)])) if len(r) > 200 else "" for r in reads[:16]]
        out1[name] = e
        print("-f 1", name, "reads", len(reads), "unstable", e["stable"].count(False), "cords", sum(e["n_cords"]))
    json.dump(out1, open(os.path.join(HERE, "golden_f1.json"), "w"), indent=0)

if ONLY in ("all", "c0"):
    # -c 0 (apxMap with f_chain = 0, alg_type 1) under the canonical parameter convention of oracle/ref_harness.cpp
    # (ref_map_batch_c0: a fresh PMPParms per read; gdl_state 0 / 1 = GetDHitListParms as constructed / after a toggle(0))
    outc = {}
    for name in ("clean_hifi", "repeat_ont", "repeat_t1_p0"):
        g, reads, bases, offs, T, preset = make_case(name)
        e = {}
        for ft in (2, 1):
            R = RefImpl(g, threads=T, preset=preset, feature_type=ft)
            for st in (0, 1):
                ca, oa = R.map_batch(bases, offs, map_threads=4, no_chain=True, gdl_state=st)
                cb, ob = R.map_batch(bases, offs, map_threads=2, no_chain=True, gdl_state=st)
                stable = [bool(np.array_equal(ca[int(oa[i]):int(oa[i + 1])], cb[int(ob[i]):int(ob[i + 1])])) for i in range(len(reads))]
                e[f"f{ft}_s{st}"] = {"stable": stable, "n_cords": [int(oa[i + 1] - oa[i]) for i in range(len(reads))],
                                     "cords": [digest(ca[int(oa[i]):int(oa[i + 1])]) for i in range(len(reads))]}
                print("-c 0", name, "-f", ft, "state", st, "unstable", stable.count(False), "cords", int(oa[-1]))
        outc[name] = e
    json.dump(outc, open(os.path.join(HERE, "golden_c0.json"), "w"), indent=0)
