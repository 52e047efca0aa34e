#!/usr/bin/env python3
"""e2e sk_triangle with and without the pipeline, with SK_TRACE timestamps. Usage: trace_triangle.py [n_genomes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import skani_b200 as sk
from bench_support import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = 5_000_000
pinned = torch.empty(n * L, dtype=torch.uint8, pin_memory=True)
host = pinned.numpy()
synth.generate(0, n, L, out=host)
off, goc = synth.layout(0, n, L)
ctx = sk.Context(0)
for mode in ("SK_NO_PIPELINE", None, None, "SK_NO_PIPELINE", None):
    os.environ.pop("SK_NO_PIPELINE", None)
    if mode:
        os.environ[mode] = "1"
    os.environ["SK_TRACE"] = "1"; os.environ["SK_PIPELINE"] = "1"
    t0 = time.perf_counter()
    res, st = sk.triangle(ctx, host, off, goc, n, as_array=True)
    print("mode=%s  %.1f ms  kept=%d" % (mode or "pipelined", (time.perf_counter() - t0) * 1e3, len(res)), flush=True)
    print(torch.cuda.mem_get_info(), flush=True)
