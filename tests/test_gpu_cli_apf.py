"""GPU (-m gpu): whole-program parity. The C++ CLI mirror (linear_b200/csrc/host/linear_b200_filter) over the C ABI
must write the same APF as the unmodified reference binary `linear filter reads.fa genome.fa -ot 1 -g 0` (-g 0: cords
straight from apxMap).
  * -t 1 -b 0: byte-identical (one fixed 50 000-read block => deterministic blank lines)
  * -t 4 -b 1: identical after stripping blank lines. `-b 0` at -t > 1 is NOT reproducible in the reference itself:
    all threads share one PMPParms and any read that re-maps toggles it under the others (mapper.cpp:836,
    pmpfinder.cpp:2756-2761; 5 runs gave 5 different files on this very input), `-b 1` uses per-thread copies."""
import os
import subprocess

import pytest

from cases import make_case
from cpu_checkers import ROOT
from linear_b200 import datagen

pytestmark = pytest.mark.gpu
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "linear")
CLI = os.path.join(ROOT, "linear_b200", "csrc", "host", "linear_b200_filter")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/linear did not travel")
@pytest.mark.parametrize("name,threads,preset,bal", [("clean_hifi", 1, 0, 0), ("clean_hifi", 4, 1, 1), ("repeat_ont", 4, 1, 1)])
def test_cli_apf_identical_to_reference_binary(tmp_path, name, threads, preset, bal):
    assert os.path.exists(CLI), "run __graft_entry__.build()"
    g, reads, bases, offs, T, _ = make_case(name)
    gfa, rfa = str(tmp_path / "genome.fa"), str(tmp_path / "reads.fa")
    datagen.write_fasta(gfa, [f"chr{i + 1} synthetic" for i in range(len(g))], g)
    datagen.write_fasta(rfa, [f"read{i}" for i in range(len(reads))], reads)
    d_ref, d_new = tmp_path / "ref", tmp_path / "new"
    d_ref.mkdir(); d_new.mkdir()
    common = ["filter", rfa, gfa, "-ot", "1", "-t", str(threads), "-p", str(preset), "-g", "0", "-b", str(bal)]
    subprocess.run([REF_BIN] + common, cwd=d_ref, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
    subprocess.run([CLI] + common, cwd=d_new, check=True, timeout=900)
    a = open(d_ref / "reads.apf", "rb").read()
    b = open(d_new / "reads.apf", "rb").read()
    assert len(a) > 1000
    if bal:
        a = b"\n".join(l for l in a.split(b"\n") if l)
        b = b"\n".join(l for l in b.split(b"\n") if l)
    assert a == b
