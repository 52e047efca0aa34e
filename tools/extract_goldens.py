#!/usr/bin/env python3
"""Extract the reference's golden OUTPUT rows (captured `cargo test -- --show-output` stdout,
/root/reference/test_results_versions/0.3.0) into small fixtures under tests/golden/.
Run only where /root/reference exists.  Data only (numbers printed by the reference binary).

  g8_dist_qi_robust.tsv  <- lines 153-421: `skani dist -r EC590.sketch markers.bin -q o157_reads.fastq --qi --robust`
                            (tests/integration_test.rs:158-171); columns: ANI, AF_ref, AF_query, Query_name
"""
import os
SRC = "/root/reference/test_results_versions/0.3.0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = open(SRC).read().split("\n")
rows = []
for ln in lines[151:425]:
    f = ln.split("\t")
    if len(f) == 7 and f[0].endswith("e.coli-EC590.fasta") and f[1].endswith("o157_reads.fastq"):
        rows.append((f[2], f[3], f[4], f[6]))
assert len(rows) == 269, len(rows)
with open(os.path.join(ROOT, "tests/golden/g8_dist_qi_robust.tsv"), "w") as o:
    o.write("# source: skani v0.3.0 test_results_versions/0.3.0:153-421 (dist --qi --robust, o157_reads.fastq vs EC590)\n")
    o.write("# ANI\tAF_ref\tAF_query\tQuery_name\n")
    for r in rows:
        o.write("\t".join(r) + "\n")
print("wrote", len(rows), "rows")
