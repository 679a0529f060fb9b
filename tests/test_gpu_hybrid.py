"""GPU (-m gpu): the drop-in, proven with the reference's OWN host stages. oracle/_ref/linear_hybrid is the unmodified
reference program (its main, option parser, FASTA reader, `-b 1` scheduler, Mapper::p_calRecords, mapGaps, reformCords,
cords2BamLink, APF writer -- the objects oracle/build_ref.sh compiles from /root/reference) in which two symbols,
createIndexDynamic and apxMap, resolve to integration/lnr_seqan_shim.cpp, i.e. to liblnr_b200.so. Its output must equal
the all-CPU reference binary's: the GPU cords are consumed unchanged by mapGaps (default -g) and by the writers."""
import os
import subprocess

import pytest

from cases import make_case
from cpu_checkers import ROOT
from linear_b200 import datagen

pytestmark = pytest.mark.gpu
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "linear")
HYBRID = os.path.join(ROOT, "oracle", "_ref", "linear_hybrid")


def _strip(b):
    return b"\n".join(l for l in b.split(b"\n") if l)


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(HYBRID)), reason="oracle/_ref (compiled reference + hybrid) did not travel")
@pytest.mark.parametrize("extra", [[], ["-g", "0"], ["-p", "0"], ["-i", "2"]])
def test_reference_host_stages_consume_gpu_cords(tmp_path, extra):
    g, reads, bases, offs, T, _ = make_case("clean_hifi")
    gfa, rfa = str(tmp_path / "genome.fa"), str(tmp_path / "reads.fa")
    datagen.write_fasta(gfa, [f"chr{i + 1} synthetic" for i in range(len(g))], g)
    datagen.write_fasta(rfa, [f"read{i}" for i in range(len(reads))], reads)
    d_ref, d_new = tmp_path / "ref", tmp_path / "new"
    d_ref.mkdir(); d_new.mkdir()
    common = ["filter", rfa, gfa, "-ot", "1", "-t", "4", "-b", "1"] + extra     # default -g: mapGaps runs on the GPU cords
    subprocess.run([REF_BIN] + common, cwd=d_ref, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
    env = dict(os.environ, LNR_ARENA_KB="1024")           # one small context per calling thread of the reference
    p = subprocess.run([HYBRID] + common, cwd=d_new, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, env=env)
    assert p.returncode == 0, p.stdout[-2000:]
    a = _strip(open(d_ref / "reads.apf", "rb").read())
    b = _strip(open(d_new / "reads.apf", "rb").read())
    assert len(a) > 1000 and a.count(b"\n@") + 1 >= 40
    assert a == b
