"""CPU: the host CLI's FASTA/FASTQ(.gz) reader (skani_b200/cli/fastx.hpp, through `skani-db-tool fastx`) follows the
needletail record rules skani relies on: compared with the tests' own Python reader on the committed
fixtures and on edge cases (CRLF, wrapped lines, blank trailing lines, lowercase / IUPAC bytes kept verbatim, FASTQ,
multi-member gzip, empty and non-FASTX files)."""
import gzip
import os
import subprocess

import pytest

from fasta_py import read_fastx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from conftest import db_tool
GOLD = os.path.join(ROOT, "tests", "golden")


def fnv(b):
    h = 0xcbf29ce484222325
    for ch in b:
        h = ((h ^ ch) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def tool(path):
    out = subprocess.run([db_tool(), "fastx", str(path)], check=True, capture_output=True).stdout.decode().split("\n")
    if out[0] == "ERR":
        return None
    n = int(out[0].split()[1])
    recs = []
    for ln in out[1:1 + n]:
        length, h, name = ln.split(" ", 2)
        recs.append((name, int(length), int(h)))
    return recs


def expect(path):
    return [(n, len(s), fnv(s)) for n, s in read_fastx(str(path))]


@pytest.mark.parametrize("name", ["viruses.fna", "o157_reads.fa.gz", "e.coli-K12.fasta.gz"])
def test_fixtures(name):
    p = os.path.join(GOLD, name)
    got = tool(p)
    assert got is not None and got == expect(p) and len(got) > 0


def test_edge_cases(tmp_path):
    cases = {
        "crlf.fa": b">a desc\r\nACGT\r\nacgtn\r\n>b\r\nNNRYK\r\n",
        "wrapped.fa": b">x\nAC\nGT\n\nAC\n>y\n\n>z\nT\n\n\n",                 # blank lines inside / empty record / trailing blanks
        "noeol.fa": b">only\nACGTACGT",
        "reads.fq": b"@r1 extra\nACGT\n+\nIIII\n@r2\nGG\n+r2\n@@\n",            # quality line starting with '@'
        "tab\tname.fa": b">id\twith\ttabs and spaces \nAAAA\n",
    }
    for name, data in cases.items():
        p = tmp_path / name
        p.write_bytes(data)
        assert tool(p) == expect(p), name
    assert tool(tmp_path / "wrapped.fa") == [("x", 6, fnv(b"ACGTAC")), ("y", 0, fnv(b"")), ("z", 1, fnv(b"T"))]
    assert tool(tmp_path / "reads.fq") == [("r1 extra", 4, fnv(b"ACGT")), ("r2", 2, fnv(b"GG"))]
    # gzip, two members concatenated (flate2 MultiGzDecoder semantics)
    p = tmp_path / "multi.fa.gz"
    p.write_bytes(gzip.compress(b">m1\nACGT\n") + gzip.compress(b"TTTT\n>m2\nCC\n"))
    assert tool(p) == [("m1", 8, fnv(b"ACGTTTTT")), ("m2", 2, fnv(b"CC"))]
    # errors: empty file, not FASTX, truncated FASTQ, missing file (the caller warns and skips, src/file_io.rs:159-166)
    for name, data in {"empty.fa": b"", "text.txt": b"this is not fasta\n", "trunc.fq": b"@r\nACGT\n+\n", "badq.fq": b"@r\nACGT\n+\nII\n"}.items():
        p = tmp_path / name
        p.write_bytes(data)
        assert tool(p) is None, name
    assert tool(tmp_path / "missing.fa") is None
