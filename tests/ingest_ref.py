"""Plain-Python statement of the read-ingest semantics (test infrastructure): the record model of seqan's readRecords +
Dna5 conversion as the reference uses it (loadRecords base.cpp:154), identical to the host reader of the CLI mirror
(linear_b200/csrc/host/lnr_filter_main.cpp), which tests/test_gpu_cli_apf.py pins against the reference binary."""
import numpy as np

_ORD = np.full(256, 4, np.uint8)
for ch, v in (("Aa", 0), ("Cc", 1), ("Gg", 2), ("TtUu", 3)):
    for c in ch:
        _ORD[ord(c)] = v


def parse_reads(text: bytes, cut_id_at_space: bool = False):
    """-> (bases uint8[], offsets uint64[n+1], ids list[str]) for FASTA (first byte '>') or 4-line FASTQ ('@')"""
    lines = text.split(b"\n")
    ids, seqs = [], []
    if text[:1] == b">":
        for ln in lines:
            if ln[:1] == b">":
                i = ln[1:]
                if i.endswith(b"\r"):
                    i = i[:-1]
                if cut_id_at_space:
                    i = i.split(b" ")[0]
                ids.append(i.decode("latin1"))
                seqs.append([])
            elif seqs:
                seqs[-1].append(ln.replace(b"\r", b"").replace(b" ", b""))
    elif text[:1] == b"@":
        for k in range(0, len(lines) - 1, 4):
            if not lines[k]:
                continue
            i = lines[k][1:]
            if i.endswith(b"\r"):
                i = i[:-1]
            if cut_id_at_space:
                i = i.split(b" ")[0]
            ids.append(i.decode("latin1"))
            seqs.append([lines[k + 1].replace(b"\r", b"")] if k + 1 < len(lines) else [])
    else:
        raise ValueError("not FASTA / FASTQ")
    off = [0]
    out = []
    for s in seqs:
        b = b"".join(s)
        out.append(_ORD[np.frombuffer(b, np.uint8)])
        off.append(off[-1] + len(b))
    bases = np.concatenate(out) if out else np.empty(0, np.uint8)
    return bases.astype(np.uint8), np.asarray(off, np.uint64), ids
