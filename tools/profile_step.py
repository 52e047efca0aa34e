#!/usr/bin/env python3
"""Stage / kernel breakdown of one triangle step (device-resident input): wall clock per stage + CUDA-event time per
kernel (sk_ctx_set_timing).  Usage: python tools/profile_step.py [n_genomes] [genome_len]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import skani_b200 as sk  # noqa: E402
from bench_support import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
L = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
pinned = torch.empty(n * L, dtype=torch.uint8, pin_memory=True)
host = pinned.numpy()
synth.generate(0, n, L, out=host)
off, goc = synth.layout(0, n, L)
ctx = sk.Context(0)
dev = pinned.to("cuda")
sp, mp = sk.sketch_params(), sk.map_params()


def run(timing):
    ctx.set_timing(timing)
    ctx.get_timing(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gs = sk.sketch_contigs(ctx, None, off, goc, n, sp, device_ptr=dev.data_ptr())
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pairs = sk.screen_triangle(ctx, gs, mp)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    res = sk.chain_pairs(ctx, gs, gs, pairs, mp, as_array=True)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    gs.free()
    t4 = time.perf_counter()
    tk = ctx.get_timing(reset=True)
    return (t1 - t0, t2 - t1, t3 - t2, t4 - t3), tk, len(pairs)


run(False)
for timing in (False, True):
    w, tk, npairs = run(timing)
    print("timing=%s n=%d pairs=%d  sketch %.1f ms  screen %.1f ms  chain %.1f ms  free %.1f ms" %
          (timing, n, npairs, w[0] * 1e3, w[1] * 1e3, w[2] * 1e3, w[3] * 1e3))
    for k, (ms, cnt) in sorted(tk.items(), key=lambda kv: -kv[1][0]):
        print("   %-20s %9.3f ms  %5d launches" % (k, ms, cnt))
t0 = time.perf_counter()
res, st = sk.triangle(ctx, host, off, goc, n, sp, mp, as_array=True)
print("e2e triangle %.1f ms (sketch %.1f screen %.1f chain %.1f)" % ((time.perf_counter() - t0) * 1e3, st.t_sketch * 1e3,
                                                                      st.t_screen * 1e3, st.t_chain * 1e3))
# raw H2D bandwidth of the same pinned buffer (reference point for the e2e leg)
torch.cuda.synchronize()
t0 = time.perf_counter()
dev.copy_(pinned, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("raw H2D %.1f GB/s (%.1f ms for %.2f GB)" % (n * L / dt / 1e9, dt * 1e3, n * L / 1e9))
