"""CPU: the Python statement of the reader semantics against hand-written expectations (it is the checker of
tests/test_gpu_ingest.py)."""
import numpy as np

from ingest_ref import parse_reads


def test_fasta_wrapping_crlf_case_iupac_and_ids():
    text = b">r1 first read\r\nACgt\r\nNNRY\r\n\r\n>r2\nuU-*\nA C\n>empty\n>last\nT"
    b, o, ids = parse_reads(text)
    assert ids == ["r1 first read", "r2", "empty", "last"]
    assert o.tolist() == [0, 8, 14, 14, 15]
    assert b.tolist() == [0, 1, 2, 3, 4, 4, 4, 4, 3, 3, 4, 4, 0, 1, 3]
    assert parse_reads(text, True)[2][0] == "r1"


def test_fastq_four_line_records():
    text = b"@q1 x\nACGN\n+\nIIII\n@q2\nTT\n+q2\n@@\n"
    b, o, ids = parse_reads(text, True)
    assert ids == ["q1", "q2"] and o.tolist() == [0, 4, 6] and b.tolist() == [0, 1, 2, 4, 3, 3]
