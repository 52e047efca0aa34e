#!/usr/bin/env python3
"""Ingestion rate of the host CLI on FASTA inputs: writes N synthetic genomes as plain FASTA,
gzip and block-gzip (BGZF) files and times `skani-b200 ingest` (the reader + record rules + flat-buffer layout that
triangle / dist / sketch run before any GPU work) at several thread counts.  Usage: tools/bench_ingest.py [n_genomes] [genome_len]"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench_support import synth  # noqa: E402
from test_fastx_reader import bgzf_compress  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
BIN = os.path.join(ROOT, "skani_b200", "skani-b200")
bases, off, goc = synth.generate(0, n, L)
ncpu = len(os.sched_getaffinity(0))
with tempfile.TemporaryDirectory() as d:
    sets = {"plain": [], "gzip": [], "bgzf": []}
    for g in range(n):
        seq = bases[g * L:(g + 1) * L].tobytes()
        text = b">g%06d\n" % g + b"\n".join(seq[i:i + 80] for i in range(0, len(seq), 80)) + b"\n"
        for kind, data in (("plain", text), ("gzip", gzip.compress(text, 6)), ("bgzf", bgzf_compress(text))):
            p = os.path.join(d, "g%06d.%s.fa%s" % (g, kind, "" if kind == "plain" else ".gz"))
            open(p, "wb").write(data)
            sets[kind].append(p)
    rows = []
    for kind, files in sets.items():
        for t in sorted({1, 4, ncpu}):
            out = subprocess.run([BIN, "ingest", "-t", str(t)] + files, capture_output=True, text=True, check=True).stdout
            r = json.loads(out.strip().split("\n")[-1])
            r["kind"] = kind
            rows.append(r)
            print("%-6s t=%-3d  %8.1f MB/s of file bytes  %8.1f Mbases/s" % (kind, t, r["file_MB_per_s"], r["bases_MB_per_s"]))
    one = subprocess.run([BIN, "ingest", "-t", str(ncpu), sets["bgzf"][0]], capture_output=True, text=True, check=True).stdout
    r = json.loads(one.strip().split("\n")[-1]); r["kind"] = "bgzf, ONE file, member-parallel inflate"
    rows.append(r)
    print("%-6s t=%-3d  %8.1f MB/s of file bytes  %8.1f Mbases/s  (one file, member-parallel)" % ("bgzf", ncpu, r["file_MB_per_s"], r["bases_MB_per_s"]))
    # ONE ordinary gzip member holding many contigs (a metagenome assembly read with -i): block-parallel two-pass inflate
    k = min(n, 24)
    big = b"".join(b">c%06d\n" % g + b"\n".join(bases[g * L:(g + 1) * L].tobytes()[i:i + 80] for i in range(0, L, 80)) + b"\n" for g in range(k))
    bp = os.path.join(d, "big.fa.gz")
    open(bp, "wb").write(gzip.compress(big, 6))
    for t, env, label in ((1, {}, "serial"), (ncpu, {"SK_SERIAL_INFLATE": "1"}, "serial decoder, %d threads given" % ncpu), (ncpu, {}, "block-parallel")):
        one = subprocess.run([BIN, "ingest", "-i", "-t", str(t), bp], capture_output=True, text=True, check=True, env=dict(os.environ, **env)).stdout
        r = json.loads(one.strip().split("\n")[-1]); r["kind"] = "gzip, ONE file of %d contigs, %s" % (k, label)
        rows.append(r)
        print("%-6s t=%-3d  %8.1f MB/s of file bytes  %8.1f Mbases/s  (one file of %d contigs, %s)" % ("gzip", t, r["file_MB_per_s"], r["bases_MB_per_s"], k, label))
    print(json.dumps({"host_cpus": ncpu, "genomes": n, "genome_len": L, "rows": rows}))
