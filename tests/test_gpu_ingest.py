"""GPU (-m gpu): read ingest on the device (lnr_reads_parse) against the plain-Python statement of the reader semantics,
and lnr_apxmap_reads (map straight from the parsed device buffers) against the host-buffer call."""
import numpy as np
import pytest

from cases import make_case
from ingest_ref import parse_reads

pytestmark = pytest.mark.gpu
ALPHA = np.frombuffer(b"ACGTN", np.uint8)


@pytest.fixture(scope="module")
def lb():
    import linear_b200
    linear_b200.load_library()
    return linear_b200


@pytest.fixture(scope="module")
def ctx(lb):
    return lb.Context(0)


def fasta_text(reads, rng, wrap=None, crlf=False, lower=False, iupac=False, ids_with_blanks=False):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for k, r in enumerate(reads):
        s = ALPHA[r].tobytes()
        if lower:
            s = bytes(c | 0x20 if rng.random() < 0.3 else c for c in s)
        if iupac:
            b = bytearray(s)
            for p in rng.integers(0, max(len(b), 1), size=max(len(b) // 200, 1)):
                if len(b):
                    b[int(p)] = rng.choice(np.frombuffer(b"RYKMSWBDHVUu-.*", np.uint8))
            s = bytes(b)
        name = f"read_{k} len={len(r)} x" if ids_with_blanks else f"read_{k}"
        out.append(b">" + name.encode() + nl)
        if wrap:
            w = wrap if isinstance(wrap, int) else int(rng.integers(1, 120))
            for i in range(0, len(s), w):
                out.append(s[i:i + w] + nl)
            if not s:
                out.append(nl)
        else:
            out.append(s + nl)
    return b"".join(out)


@pytest.mark.parametrize("variant", ["plain", "wrapped60", "ragged_crlf_lower_iupac", "ids_cut", "no_trailing_newline", "fastq"])
def test_parse_matches_reader_semantics(lb, ctx, variant):
    rng = np.random.default_rng(11)
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    reads = list(reads) + [np.zeros(0, np.uint8), rng.integers(0, 5, size=33, dtype=np.uint8)]   # an empty and a tiny record
    cut = False
    if variant == "plain":
        text = fasta_text(reads, rng)
    elif variant == "wrapped60":
        text = fasta_text(reads, rng, wrap=60)
    elif variant == "ragged_crlf_lower_iupac":
        text = fasta_text(reads, rng, wrap="random", crlf=True, lower=True, iupac=True)
    elif variant == "ids_cut":
        text = fasta_text(reads, rng, wrap=80, ids_with_blanks=True)
        cut = True
    elif variant == "no_trailing_newline":
        text = fasta_text(reads, rng, wrap=70).rstrip(b"\n")
    else:
        text = b"".join(b"@q%d some text\n" % k + ALPHA[r].tobytes() + b"\n+\n" + b"I" * len(r) + b"\n" for k, r in enumerate(reads))
        cut = True
    want_b, want_o, want_ids = parse_reads(text, cut)
    R = lb.Reads(ctx, text, cut_id_at_space=cut)
    got_b, got_o, got_ids = R.download()
    assert R.n_reads == len(want_ids) and got_ids == want_ids
    assert np.array_equal(got_o, want_o)
    assert np.array_equal(got_b, want_b)
    R.close()


def test_map_from_parsed_reads_equals_host_buffer_call(lb, ctx):
    rng = np.random.default_rng(5)
    g, reads, bases, offs, T, preset = make_case("repeat_ont")
    text = fasta_text(reads, rng, wrap=60)
    gen = lb.Genome(ctx, g)
    feats = lb.create_features(ctx, gen, 2, T)
    index = lb.create_index(ctx, gen, 1, T)
    want_c, want_o = lb.apx_map_batch(ctx, index, feats, bases, offs, preset=preset)
    R = lb.Reads(ctx, text)
    got_c, got_o = R.apx_map(index, feats, preset=preset)
    assert np.array_equal(got_o, want_o) and np.array_equal(got_c, want_c)
    # a slice in the middle (arbitrary alignment of the first base)
    a, n = 3, 9
    sub_o = (offs[a:a + n + 1] - offs[a]).astype(np.uint64)
    sub_b = bases[int(offs[a]):int(offs[a + n])]
    wc, wo = lb.apx_map_batch(ctx, index, feats, sub_b, sub_o, preset=preset)
    gc, go = R.apx_map(index, feats, first=a, n=n, preset=preset)
    assert np.array_equal(go, wo) and np.array_equal(gc, wc)


def test_rejects_text_that_is_not_fasta_or_fastq(lb, ctx):
    with pytest.raises(lb.LnrError):
        lb.Reads(ctx, b"ACGT\nACGT\n")
    R = lb.Reads(ctx, b"")
    assert R.n_reads == 0 and R.total_bases == 0
