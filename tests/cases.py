"""Seeded parity cases shared by the golden-vector generator, the CPU tests and the GPU tests."""
from __future__ import annotations

import hashlib
from typing import Dict, List

import numpy as np

from linear_b200 import datagen


def _junk_and_chimeras(rs, seed: int, n: int) -> List[np.ndarray]:
    """reads that do not map (or only partly): they drive the re-map path (pmpfinder.cpp:2749-2767)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        a = rs.read(i % rs.n)
        junk = rng.integers(0, 4, size=int(rng.integers(3000, 15000)), dtype=np.uint8)
        out.append(np.concatenate([junk, a[: len(a) // 3]]) if i % 2 else junk)
    return out


def _pack(reads: List[np.ndarray]):
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    for i, r in enumerate(reads):
        offs[i + 1] = offs[i] + len(r)
    bases = np.concatenate(reads) if reads else np.zeros(0, np.uint8)
    return np.ascontiguousarray(bases, dtype=np.uint8), offs


CASES: Dict[str, dict] = {
    # repeat-free genome, HiFi-like reads: any correct sort passes (SURVEY 7.2 stage 1)
    "clean_hifi": dict(total=2_000_000, n_contigs=2, gseed=1, fam=0, copies=0, tandem=0, rseed=11, n_reads=48,
                       mean=9000, sd=2500, err=0.01, mix=(1, 1, 1), sv=0.0, lognormal=False, junk=0, threads=4, preset=1),
    # repeat-rich genome, ONT-like reads with planted SVs + junk/chimeric reads: comparator ties, re-map
    "repeat_ont": dict(total=3_000_000, n_contigs=3, gseed=17, fam=6, copies=800, tandem=150, rseed=13, n_reads=64,
                       mean=10000, sd=3000, err=0.10, mix=(4, 3, 3), sv=0.4, lognormal=True, junk=12, threads=4, preset=1),
    # same data at other -t / -p: chunk seams and stop ratio are semantic
    "repeat_t1_p0": dict(total=3_000_000, n_contigs=3, gseed=17, fam=6, copies=800, tandem=150, rseed=13, n_reads=32,
                         mean=10000, sd=3000, err=0.10, mix=(4, 3, 3), sv=0.4, lognormal=True, junk=6, threads=1, preset=0),
    "repeat_t16": dict(total=3_000_000, n_contigs=3, gseed=17, fam=6, copies=800, tandem=150, rseed=13, n_reads=32,
                       mean=10000, sd=3000, err=0.10, mix=(4, 3, 3), sv=0.4, lognormal=True, junk=6, threads=16, preset=1),
}


def make_case(name: str):
    c = CASES[name]
    lens = datagen.contig_lengths(c["total"], c["n_contigs"], seed=c["gseed"])
    g = datagen.make_genome(c["gseed"], lens, n_families=c["fam"], copies=c["copies"], n_tandem=c["tandem"])
    rs = datagen.simulate_reads(c["rseed"], g, c["n_reads"], mean_len=c["mean"], sd_len=c["sd"], err=c["err"], mix=c["mix"],
                                sv_frac=c["sv"], lognormal=c["lognormal"])
    reads = [rs.read(i) for i in range(rs.n)] + _junk_and_chimeras(rs, c["rseed"] + 1, c["junk"])
    # edge cases the reference handles explicitly: reads <= 200 bases give no cords (mapper.cpp:440)
    reads.append(reads[0][:200].copy())
    reads.append(reads[0][:201].copy())
    reads.append(reads[1][:150].copy())
    bases, offs = _pack(reads)
    return g, reads, bases, offs, c["threads"], c["preset"]


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
