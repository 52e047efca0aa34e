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


def tool(path, threads=1, env=None):
    out = subprocess.run([db_tool(), "fastx", str(path), str(threads)], check=True, capture_output=True,
                         env=None if env is None else dict(os.environ, **env)).stdout.decode().split("\n")
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


def bgzf_compress(data, block=65280):
    """Block-gzip (BGZF, as bgzip / htslib write it): independent members with their compressed size in a 'BC' extra field."""
    import struct
    import zlib
    out = bytearray()
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + [b""]   # empty EOF member
    for ch in chunks:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        raw = co.compress(ch) + co.flush()
        bsize = 12 + 6 + len(raw) + 8 - 1
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + raw
        out += struct.pack("<II", zlib.crc32(ch) & 0xFFFFFFFF, len(ch))
    return bytes(out)


def test_bgzf_member_parallel_inflate(tmp_path):
    """Block-gzipped FASTA (many members, records spanning member boundaries) reads like the plain file."""
    import random
    rnd = random.Random(5)
    recs = []
    for i in range(40):
        seq = "".join(rnd.choice("ACGTN") for _ in range(rnd.randrange(0, 30000)))
        recs.append((">c%d some description" % i, seq))
    text = "".join(h + "\n" + "\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + "\n" for h, s in recs).encode()
    plain = tmp_path / "x.fa"
    plain.write_bytes(text)
    bg = tmp_path / "x.fa.gz"
    bg.write_bytes(bgzf_compress(text, block=4000))
    assert gzip.decompress(bg.read_bytes()) == text          # a valid multi-member gzip for every other reader too
    assert tool(bg) == tool(plain) == expect(plain)
    # a BGZF prefix followed by an ordinary member falls back to the sequential stream
    mixed = tmp_path / "m.fa.gz"
    mixed.write_bytes(bgzf_compress(b">a\nACGT\n")[:-28] + gzip.compress(b"GGGG\n>b\nTT\n"))
    assert tool(mixed) == [("a", 8, fnv(b"ACGTGGGG")), ("b", 2, fnv(b"TT"))]


def test_block_parallel_inflate_of_one_member(tmp_path):
    """ONE ordinary gzip member read with several threads (two-pass block-parallel decoding, here forced onto a small file with
    32 KB chunks) gives the records of the serial readers; a file that is not text, or has two members, silently takes the
    serial path."""
    import random
    rnd = random.Random(9)
    recs = []
    for i in range(30):
        seq = "".join(rnd.choice("ACGT") for _ in range(rnd.randrange(1000, 120000)))
        recs.append((">contig_%d len=%d" % (i, len(seq)), seq))
    text = "".join(h + "\n" + "\n".join(s[j:j + 80] for j in range(0, len(s), 80)) + "\n" for h, s in recs).encode()
    plain = tmp_path / "big.fa"
    plain.write_bytes(text)
    want = expect(plain)
    for level in (1, 6, 9):
        gz = tmp_path / ("big%d.fa.gz" % level)
        gz.write_bytes(gzip.compress(text, level))
        serial = tool(gz, 1)
        par = tool(gz, 6, env={"SK_INFLATE_MIN_CHUNK": "32768"})
        assert serial == par == want
        err = subprocess.run([db_tool(), "fastx", str(gz), "6"], capture_output=True, env=dict(os.environ, SK_INFLATE_MIN_CHUNK="32768", SK_TRACE="1")).stderr.decode()
        assert "[inflate] block-parallel:" in err, err                      # the parallel path really ran (it did not decline)
    two = tmp_path / "two.fa.gz"
    two.write_bytes(gzip.compress(text, 6) + gzip.compress(b">extra\nACGTACGT\n", 6))
    got = tool(two, 6, env={"SK_INFLATE_MIN_CHUNK": "32768"})
    assert got[:-1] == want and got[-1] == ("extra", 8, fnv(b"ACGTACGT"))
    err = subprocess.run([db_tool(), "fastx", str(two), "6"], capture_output=True, env=dict(os.environ, SK_INFLATE_MIN_CHUNK="32768", SK_TRACE="1")).stderr.decode()
    assert "declined" in err
