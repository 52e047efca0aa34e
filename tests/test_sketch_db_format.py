"""CPU: the skani v0.3.0 sketch database writer / reader of the CLI (skani_b200/cli/sketch_db.hpp) against an independent
Python decoder of the same format (tests/skani_db_py.py) on sketches produced by the oracle.  The reference itself pins
the v0.3 byte layout only through round trips (tests/integration_test.rs:598-701; SURVEY.md section 8c 'parity gaps'),
so this is the same level of evidence: layout derived from the serde struct definitions, writer and reader checked
against each other and against a second implementation."""
import os
import subprocess

import numpy as np

import oracle_py as O
import skani_db_py as D
from fasta_py import read_fastx as read_fasta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from conftest import db_tool
GOLD = os.path.join(ROOT, "tests", "golden")


def text_of(name, order, sk, contig_names):
    e = sk.export()
    lines = ["S %d %d %s" % (order, sk.total_len, name)]
    lines += ["C " + c for c in contig_names]
    lines.append("L %d " % len(e["contig_lengths"]) + " ".join(map(str, e["contig_lengths"].tolist())))
    rec = np.stack([e["kmer"], e["pos"], e["cc"]], 1).reshape(-1)
    lines.append("R %d " % len(e["kmer"]) + " ".join(map(str, rec.tolist())))
    lines.append("M %d " % len(e["markers"]) + " ".join(map(str, e["markers"].tolist())))
    lines.append("E")
    return "\n".join(lines) + "\n", e


def test_db_roundtrip_against_python_decoder(tmp_path):
    rng = np.random.default_rng(3)
    sketches = []
    # a multi-contig genome with repeated k-mers (multi-position entries), a one-contig genome, and a genome with no seeds
    a = rng.choice(np.frombuffer(b"ACGT", np.uint8), 60_000)
    contigs = [a[:30_000], np.concatenate([a[10_000:25_000], a[40_000:]]), a[5_000:6_000]]
    sketches.append(("dir/with space/a genome.fa", 0, O.sketch_from_contigs("a", contigs, c=30, k=15, marker_c=200),
                     ["ctg1 some description", "ctg2", "ctg3\tx"]))
    b = rng.choice(np.frombuffer(b"ACGT", np.uint8), 20_000)
    sketches.append(("b.fa", 7, O.sketch_from_contigs("b", [b], c=30, k=15, marker_c=200), ["only"]))
    sketches.append(("n.fa", 0, O.sketch_from_contigs("n", [np.full(700, ord("N"), np.uint8)], c=30, k=15, marker_c=200), ["all N"]))
    text, exports = "", []
    for name, order, sk, cn in sketches:
        t, e = text_of(name, order, sk, cn)
        text += t
        exports.append(e)
    d = str(tmp_path / "db")
    os.makedirs(d)
    subprocess.run([db_tool(), "write", d, "30", "15", "200"], input=text.encode(), check=True)
    par, sk, mk, index = D.read_db(d)
    # SketchParams blob: all 626 bytes as SketchParams::new(200, 30, 15, false, false) serialises them
    assert open(os.path.join(d, "markers.bin"), "rb").read(626) == D.expected_params_bytes(30, 15, 200)
    assert (par["c"], par["k"], par["marker_c"], par["orf_size"]) == (30, 15, 200, 30)
    assert len(sk) == len(mk) == len(index) == 3
    assert any(len(set(r[0] for r in s["records"])) < len(s["records"]) for s in sk)      # multi-position entries present
    for (name, order, osk, cn), e, s, m in zip(sketches, exports, sk, mk):
        assert s["file_name"] == m["file_name"] == name and s["contig_order"] == m["contig_order"] == order
        assert s["has_seeds"] and not m["has_seeds"] and m["records"] == [] and m["contig_lengths"] == []
        assert s["records"] == list(zip(e["kmer"].tolist(), e["pos"].tolist(), e["cc"].tolist()))
        assert s["n_keys"] == osk.n_kmers
        assert s["markers"] == m["markers"] == e["markers"].tolist()
        assert s["contigs"] == m["contigs"] == cn
        assert s["contig_lengths"] == e["contig_lengths"].tolist()
        assert s["total_len"] == m["total_len"] == osk.total_len
        assert (s["marker_c"], s["c"], s["k"]) == (30, 30, 15)           # marker_c field = c (src/types.rs:347)
        assert s["repetitive_kmers"] == 0 and s["individual_contig"] == 0 and s["amino_acid"] == 0
    # the C++ reader returns what was written
    dump = subprocess.run([db_tool(), "dump", d], check=True, capture_output=True).stdout.decode().splitlines()
    assert dump[0] == "PARAMS 30 15 200 0 0 30" and dump[1] == "N 3 3"
    want = [l for l in text.splitlines()]
    got_full = [l for l in dump[2:] if not l.startswith("P ")]
    n_full = len(want)
    assert [l.rstrip() for l in got_full[:n_full]] == [l.rstrip() for l in want]
    assert sum(1 for l in dump if l.startswith("K ")) == 3
    k_at = [i for i, l in enumerate(dump) if l.startswith("K ")]
    assert all(any(l.startswith("R 0") for l in dump[i:i + 8]) for i in k_at)     # marker sketches carry no records


def test_db_of_real_genome_roundtrips(tmp_path):
    """one of the reference's E. coli fixtures through writer -> reader (40 k records, ~10 % in multi-position groups)"""
    recs = read_fasta(os.path.join(GOLD, "e.coli-EC590.fasta.gz"))
    osk = O.sketch_from_contigs("EC590", [s for _n, s in recs if len(s) >= 500])
    text, e = text_of("refs/e.coli-EC590.fasta.gz", 0, osk, [n for n, s in recs if len(s) >= 500])
    d = str(tmp_path / "db")
    os.makedirs(d)
    subprocess.run([db_tool(), "write", d, "125", "15", "1000"], input=text.encode(), check=True)
    par, sk, mk, index = D.read_db(d)
    assert sk[0]["records"] == list(zip(e["kmer"].tolist(), e["pos"].tolist(), e["cc"].tolist()))
    assert sk[0]["markers"] == e["markers"].tolist() and sk[0]["n_keys"] == osk.n_kmers
    dump = subprocess.run([db_tool(), "dump", d], check=True, capture_output=True).stdout.decode().splitlines()
    assert [l.rstrip() for l in dump[2:] if not l.startswith("P ")][:len(text.splitlines())] == [l.rstrip() for l in text.splitlines()]
    # a corrupted length prefix is reported, not followed
    raw = bytearray(open(os.path.join(d, "sketches.db"), "rb").read())
    raw[626:634] = (2 ** 60).to_bytes(8, "little")
    open(os.path.join(d, "sketches.db"), "wb").write(bytes(raw))
    r = subprocess.run([db_tool(), "dump", d], capture_output=True)
    assert r.returncode == 1 and b"ERROR" in r.stderr
