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


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(HYBRID)), reason="oracle/_ref (compiled reference + hybrid) did not travel")
def test_hybrid_c0_follows_the_reference_parameter_state(tmp_path):
    """`-c 0` (apxMap with f_chain = 0) reads and changes the thread's PMPParms: the first read that needs the second attempt
    leaves GetDHitListParms at (10, 999) for every later read. The shim hands the state to the library per call and keeps the
    caller's object in step, so with one mapping thread the hybrid writes the reference binary's APF, junk reads included."""
    g, reads, bases, offs, T, _ = make_case("repeat_ont")
    gfa, rfa = str(tmp_path / "genome.fa"), str(tmp_path / "reads.fa")
    datagen.write_fasta(gfa, [f"chr{i + 1} synthetic" for i in range(len(g))], g)
    datagen.write_fasta(rfa, [f"read{i}" for i in range(len(reads))], reads)
    d_ref, d_new = tmp_path / "ref", tmp_path / "new"
    d_ref.mkdir(); d_new.mkdir()
    common = ["filter", rfa, gfa, "-ot", "1", "-t", "1", "-b", "0", "-g", "0", "-c", "0"]
    subprocess.run([REF_BIN] + common, cwd=d_ref, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
    env = dict(os.environ, LNR_ARENA_KB="1024")
    p = subprocess.run([HYBRID] + common, cwd=d_new, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, env=env)
    assert p.returncode == 0, p.stdout[-2000:]
    a = _strip(open(d_ref / "reads.apf", "rb").read())
    b = _strip(open(d_new / "reads.apf", "rb").read())
    assert len(a) > 1000 and a.count(b"\n@") + 1 >= 40
    assert a == b


@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(HYBRID)), reason="oracle/_ref (compiled reference + hybrid) did not travel")
def test_samples_stream_against_replicated_index(tmp_path):
    """BASELINE configs[4] in miniature: several samples are streamed at the same time (`-b 1` scheduler, -ot 3 = APF + SAM,
    host gap mapping on), one hybrid process per sample, each with its own replica of the index on the GPU it is given
    (LNR_DEVICE; all of them on GPU 0 when the box has one). Every sample's APF and SAM equal the all-CPU reference's."""
    import torch
    n_gpu = max(torch.cuda.device_count(), 1)
    g, reads, bases, offs, T, _ = make_case("clean_hifi")
    gfa = str(tmp_path / "genome.fa")
    datagen.write_fasta(gfa, [f"chr{i + 1} synthetic" for i in range(len(g))], g)
    n_samples = 3
    procs = []
    for s in range(n_samples):
        sub = reads[s::n_samples]
        rfa = str(tmp_path / f"sample{s}.fa")
        datagen.write_fasta(rfa, [f"s{s}_read{i}" for i in range(len(sub))], sub)
        d_ref, d_new = tmp_path / f"ref{s}", tmp_path / f"new{s}"
        d_ref.mkdir(); d_new.mkdir()
        common = ["filter", rfa, gfa, "-ot", "3", "-t", "2", "-b", "1"]
        subprocess.run([REF_BIN] + common, cwd=d_ref, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
        env = dict(os.environ, LNR_ARENA_KB="1024", LNR_DEVICE=str(s % n_gpu))
        procs.append((s, subprocess.Popen([HYBRID] + common, cwd=d_new, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env)))
    for s, p in procs:
        out, _ = p.communicate(timeout=900)
        assert p.returncode == 0, out[-2000:]
        for ext in ("apf", "sam"):
            a = _strip(open(tmp_path / f"ref{s}" / f"sample{s}.{ext}", "rb").read())
            b = _strip(open(tmp_path / f"new{s}" / f"sample{s}.{ext}", "rb").read())
            assert a == b, (s, ext)
        assert os.path.getsize(tmp_path / f"new{s}" / f"sample{s}.apf") > 1000
