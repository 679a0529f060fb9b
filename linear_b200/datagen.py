"""Deterministic synthetic genomes and simulated long reads.

Bases use the reference's Dna5 ordinals: A,C,G,T,N = 0..4 (include/base.h:106-116 of the reference).
Everything is a pure function of the seed. numpy only; the large bench inputs are produced by
`bench.py` with the same recipe on the GPU.
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence, Tuple

import numpy as np

_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def contig_lengths(total: int, n: int, seed: int = 0, max_len: int = 298_000_000) -> List[int]:
    """n contig lengths summing to ~total, each < max_len (the reference's binningFilter is undefined for
    diagonals >= 300 M, pmpfinder.cpp:1993) and with (len-48)%16 != 0 (genome feature builder reads one
    byte past the end otherwise, pmpfinder.cpp:596-645)."""
    rng = np.random.default_rng(seed)
    w = rng.uniform(0.6, 1.4, size=n)
    lens = np.maximum((w / w.sum() * total).astype(np.int64), 1000)
    out = []
    for v in lens:
        v = int(min(v, max_len))
        while (v - 48) % 16 == 0:
            v -= 1
        out.append(v)
    return out


def make_genome(seed: int, lens: Sequence[int], n_families: int = 0, family_len: int = 300,
                copies: int = 0, divergence: float = 0.12, n_tandem: int = 0, n_runs: int = 0) -> List[np.ndarray]:
    """i.i.d. ACGT contigs with optional planted interspersed repeat families and tandem arrays."""
    rng = np.random.default_rng(seed)
    contigs = [rng.integers(0, 4, size=int(l), dtype=np.uint8) for l in lens]
    for _ in range(n_families):
        base = rng.integers(0, 4, size=family_len, dtype=np.uint8)
        for _ in range(copies):
            ci = int(rng.integers(len(contigs)))
            if len(contigs[ci]) <= family_len + 1:
                continue
            pos = int(rng.integers(0, len(contigs[ci]) - family_len))
            cp = base.copy()
            mut = rng.random(family_len) < divergence
            cp[mut] = rng.integers(0, 4, size=int(mut.sum()), dtype=np.uint8)
            contigs[ci][pos:pos + family_len] = cp
    for _ in range(n_tandem):
        ci = int(rng.integers(len(contigs)))
        unit = rng.integers(0, 4, size=int(rng.integers(20, 200)), dtype=np.uint8)
        n = int(rng.integers(5, 50))
        arr = np.tile(unit, n)
        if len(contigs[ci]) <= len(arr) + 1:
            continue
        pos = int(rng.integers(0, len(contigs[ci]) - len(arr)))
        contigs[ci][pos:pos + len(arr)] = arr
    for _ in range(n_runs):  # homopolymer / N runs, for the N regression pair only
        ci = int(rng.integers(len(contigs)))
        n = int(rng.integers(10, 200))
        if len(contigs[ci]) <= n + 1:
            continue
        pos = int(rng.integers(0, len(contigs[ci]) - n))
        contigs[ci][pos:pos + n] = 4
    return contigs


@dataclasses.dataclass
class ReadSet:
    bases: np.ndarray      # uint8, all reads back to back
    offsets: np.ndarray    # uint64, n+1
    truth: List[Tuple[int, int, int, str]]  # (contig, start, strand, sv kind)

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    def read(self, i: int) -> np.ndarray:
        return self.bases[int(self.offsets[i]):int(self.offsets[i + 1])]


def _apply_errors(rng, tpl: np.ndarray, err: float, mix=(1, 1, 1)) -> np.ndarray:
    if err <= 0:
        return tpl
    n = len(tpl)
    tot = float(sum(mix))
    p_sub, p_ins, p_del = (err * m / tot for m in mix)
    r = rng.random(n)
    is_sub = r < p_sub
    is_ins = (r >= p_sub) & (r < p_sub + p_ins)
    is_del = (r >= p_sub + p_ins) & (r < p_sub + p_ins + p_del)
    out = tpl.copy()
    ns = int(is_sub.sum())
    if ns:
        out[is_sub] = (out[is_sub] + rng.integers(1, 4, size=ns, dtype=np.uint8)) % 4
    reps = np.ones(n, dtype=np.int64)
    reps[is_del] = 0
    reps[is_ins] = 2
    res = np.repeat(out, reps)
    # the duplicated base of an insertion is replaced by a random base
    ins_pos = np.cumsum(reps)[is_ins] - 1
    if len(ins_pos):
        res[ins_pos] = rng.integers(0, 4, size=len(ins_pos), dtype=np.uint8)
    return res


def simulate_reads(seed: int, contigs: Sequence[np.ndarray], n_reads: int, mean_len: int = 15000, sd_len: int = 3000,
                   err: float = 0.01, mix=(1, 1, 1), rev_frac: float = 0.5, sv_frac: float = 0.0,
                   min_len: int = 2000, lognormal: bool = False, max_len: int = (1 << 20) - 1) -> ReadSet:
    """Reads sampled from the genome with sub/ins/del errors and optionally one planted SV
    (ins 100-2000 random, del 100-3000, inv 500-3000, tandem dup 300-2000)."""
    rng = np.random.default_rng(seed)
    clens = np.array([len(c) for c in contigs], dtype=np.float64)
    pc = clens / clens.sum()
    chunks, offs, truth = [], [0], []
    for _ in range(n_reads):
        if lognormal:
            sigma = 0.5
            L = int(rng.lognormal(np.log(mean_len) - sigma * sigma / 2, sigma))
        else:
            L = int(rng.normal(mean_len, sd_len))
        L = max(min_len, min(L, max_len))
        ci = int(rng.choice(len(contigs), p=pc))
        L = min(L, len(contigs[ci]) - 1)
        st = int(rng.integers(0, len(contigs[ci]) - L))
        tpl = contigs[ci][st:st + L]
        kind = "none"
        if sv_frac > 0 and rng.random() < sv_frac and L > 8000:
            kind = ("ins", "del", "inv", "dup")[int(rng.integers(4))]
            p = int(rng.integers(2000, L - 4000))
            if kind == "ins":
                n = int(rng.integers(100, 2000))
                tpl = np.concatenate([tpl[:p], rng.integers(0, 4, size=n, dtype=np.uint8), tpl[p:]])
            elif kind == "del":
                n = int(rng.integers(100, min(3000, L - p - 1000)))
                tpl = np.concatenate([tpl[:p], tpl[p + n:]])
            elif kind == "inv":
                n = int(rng.integers(500, min(3000, L - p - 500)))
                tpl = np.concatenate([tpl[:p], revcomp(tpl[p:p + n]), tpl[p + n:]])
            else:
                n = int(rng.integers(300, min(2000, L - p - 500)))
                tpl = np.concatenate([tpl[:p + n], tpl[p:p + n], tpl[p + n:]])
        rd = _apply_errors(rng, tpl, err, mix)
        strand = int(rng.random() < rev_frac)
        if strand:
            rd = revcomp(rd)
        rd = rd[:max_len]
        chunks.append(np.ascontiguousarray(rd, dtype=np.uint8))
        offs.append(offs[-1] + len(rd))
        truth.append((ci, st, strand, kind))
    bases = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return ReadSet(bases=bases, offsets=np.array(offs, dtype=np.uint64), truth=truth)


_ALPH = np.frombuffer(b"ACGTN", dtype=np.uint8)


def write_fasta(path: str, names: Sequence[str], seqs: Sequence[np.ndarray], width: int = 0) -> None:
    with open(path, "wb") as f:
        for nm, s in zip(names, seqs):
            f.write(b">" + nm.encode() + b"\n")
            f.write(_ALPH[s].tobytes())
            f.write(b"\n")
