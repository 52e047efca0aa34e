"""The C++ host CLI (skani-b200 triangle / dist) over the C ABI reproduces the reference's golden OUTPUT ROWS."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "skani_b200", "skani-b200")
GOLD = os.path.join(ROOT, "tests", "golden")


def run(args, cwd=None):
    p = subprocess.run([BIN] + args, capture_output=True, text=True, cwd=cwd, timeout=600)
    assert p.returncode == 0, p.stderr
    return p.stdout, p.stderr


def test_dist_config1_and_fast_goldens():
    ec, k12 = os.path.join(GOLD, "e.coli-EC590.fasta.gz"), os.path.join(GOLD, "e.coli-K12.fasta.gz")
    out, _ = run(["dist", ec, k12])            # BASELINE.json configs[0]; README.md:142 ANI, v0.3.0 AF semantics (G13)
    rows = [ln.split("\t") for ln in out.strip().split("\n")]
    assert rows[0][:5] == ["Ref_file", "Query_file", "ANI", "Align_fraction_ref", "Align_fraction_query"]
    assert rows[1][0] == k12 and rows[1][1] == ec and rows[1][2:5] == ["99.39", "91.89", "92.46"]
    assert rows[1][5].startswith("NC_007779.1 Escherichia coli str. K-12") and rows[1][6].startswith("NZ_CP016182.2 Escherichia coli strain EC590")
    out, _ = run(["dist", ec, k12, "--fast", "-n", "3"])   # test_results_versions/0.3.0:121 (G4)
    assert out.strip().split("\n")[1].split("\t")[2:5] == ["99.42", "94.47", "95.06"]


def test_dist_qi_robust_g8_rows():
    gold = {}
    for ln in open(os.path.join(GOLD, "g8_dist_qi_robust.tsv")):
        if not ln.startswith("#"):
            ani, afr, afq, name = ln.rstrip("\n").split("\t")
            gold[name] = (ani, afr, afq)
    out, _ = run(["dist", "-r", os.path.join(GOLD, "e.coli-EC590.fasta.gz"), "-q", os.path.join(GOLD, "o157_reads.fa.gz"), "--qi", "--robust"])
    rows = [ln.split("\t") for ln in out.strip().split("\n")[1:]]
    got = {r[6]: (r[2], r[3], r[4]) for r in rows}
    assert len(rows) == 269 and got == gold
    assert [r[6] for r in rows] == sorted(r[6] for r in rows)      # grouped by query name, sorted (src/file_io.rs:631-634)


def test_triangle_matrix_and_sparse(tmp_path):
    files = [os.path.join(GOLD, f) for f in ("e.coli-K12.fasta.gz", "e.coli-EC590.fasta.gz")]
    out, err = run(["triangle"] + files, cwd=str(tmp_path))
    lines = out.strip().split("\n")
    assert lines[0] == "2" and lines[1] == files[1] and lines[2].split("\t")[0] == files[0]   # sorted by file name
    assert lines[2].split("\t")[1] in ("99.39", "99.40", "99.38")
    af = open(os.path.join(str(tmp_path), "skani_matrix.af")).read().strip().split("\n")
    assert af[0] == "2" and af[1].split("\t")[1] == "100.00"
    out2, _ = run(["triangle", "-E", "--diagonal"] + files, cwd=str(tmp_path))
    rows = out2.strip().split("\n")
    assert rows[0].startswith("Ref_file\tQuery_file\tANI") and len(rows) == 4
    assert rows[1].split("\t")[2:5] == ["100.00", "100.00", "100.00"]
    out3, _ = run(["triangle", "-i", "-E", os.path.join(GOLD, "viruses.fna")], cwd=str(tmp_path))   # tests/int_test_new.rs:58-64
    anis = sorted(float(r.split("\t")[2]) for r in out3.strip().split("\n")[1:])
    assert len(anis) == 3 and 99.0 < anis[0] < 99.9 and anis[1] > 99.9


def test_degenerate_inputs(tmp_path):
    bad = tmp_path / "bad.fa"
    bad.write_text("this is not fasta\n")
    p = subprocess.run([BIN, "triangle", str(bad), str(tmp_path / "missing.fa")], capture_output=True, text=True)
    assert p.returncode == 1 and "WARN" in p.stderr and "ERROR" in p.stderr   # tests/int_test_new.rs:135-163


def _rowset(out):
    return sorted(ln for ln in out.strip().split("\n")[1:])


def test_intermediate_flushes_and_multi_gpu_flag(tmp_path):
    """Results are appended in blocks of INTERMEDIATE_WRITE_COUNT rows / queries (src/params.rs:9, src/triangle.rs:113-138,
    src/dist.rs:151-175): with a tiny block size the row SET is unchanged and the flush is announced.  `triangle --gpus N`
    (sk_triangle_multi; contexts share the device on a 1-GPU box) prints the same rows."""
    reads = os.path.join(GOLD, "o157_reads.fa.gz")
    vir = os.path.join(GOLD, "viruses.fna")
    base, _ = run(["triangle", "-i", "-E", reads], cwd=str(tmp_path))
    env = dict(os.environ, SK_INTERMEDIATE_WRITE_COUNT="37")
    p = subprocess.run([BIN, "triangle", "-i", "-E", reads], capture_output=True, text=True, cwd=str(tmp_path), timeout=600, env=env)
    assert p.returncode == 0, p.stderr
    assert _rowset(p.stdout) == _rowset(base) and len(_rowset(base)) == 270          # test_results_versions/0.3.0:60 (G10)
    assert "INFO Writing results for 37 query sequences." in p.stderr
    multi, err = run(["triangle", "-i", "-E", "--gpus", "3", reads], cwd=str(tmp_path))
    assert _rowset(multi) == _rowset(base)
    ec = os.path.join(GOLD, "e.coli-EC590.fasta.gz")
    d0, _ = run(["dist", "-r", ec, "-q", reads, "--qi"])
    p = subprocess.run([BIN, "dist", "-r", ec, "-q", reads, "--qi"], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and _rowset(p.stdout) == _rowset(d0) and len(_rowset(d0)) == 269
    assert "INFO Writing results for 37 query sequences." in p.stderr
    v0, _ = run(["triangle", "-E", "--gpus", "2", vir, ec, os.path.join(GOLD, "e.coli-K12.fasta.gz")], cwd=str(tmp_path))
    v1, _ = run(["triangle", "-E", vir, ec, os.path.join(GOLD, "e.coli-K12.fasta.gz")], cwd=str(tmp_path))
    assert _rowset(v0) == _rowset(v1) and len(_rowset(v1)) == 1
