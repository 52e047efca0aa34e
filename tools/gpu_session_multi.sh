#!/bin/bash
# Multi-GPU session (run under gpurun --gpus N): 2-rank parity tests, torchrun bench, one-process sk_triangle_multi timing.
TAG=${1:-m}; N=${2:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/${TAG}_env.txt; cat /sys/fs/cgroup/cpu.max >> $O/${TAG}_env.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_multi_onebox.py -q 2>&1 | tail -15 > $O/${TAG}_tests.log; tail -4 $O/${TAG}_tests.log
SK_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -c 2200 $O/${TAG}_bench.json; grep "multi_gpu rank0" $O/${TAG}_bench.err | tail -6 | cut -c1-250
timeout 900 python - > $O/${TAG}_multi_c.txt 2>&1 <<PY
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import skani_b200 as sk
from bench_support import synth
n, L, N = 2000, 5_000_000, $N
pinned = torch.empty(n * L, dtype=torch.uint8, pin_memory=True)
host = pinned.numpy()
synth.generate(0, n, L, out=host)
off, goc = synth.layout(0, n, L)
ctxs = [sk.Context(d) for d in range(N)]
for rep in range(3):
    t0 = time.perf_counter()
    res, st = sk.triangle_multi(ctxs, host, off, goc, n)
    t1 = time.perf_counter()
    print("sk_triangle_multi %d GPUs, %d genomes: %.1f ms, %d kept" % (N, n, (t1 - t0) * 1e3, len(res)))
t0 = time.perf_counter(); r1, _ = sk.triangle(ctxs[0], host, off, goc, n, as_array=True); print("sk_triangle 1 GPU: %.1f ms, %d kept" % ((time.perf_counter() - t0) * 1e3, len(r1)))
t0 = time.perf_counter(); r1, _ = sk.triangle(ctxs[0], host, off, goc, n, as_array=True); print("sk_triangle 1 GPU: %.1f ms, %d kept" % ((time.perf_counter() - t0) * 1e3, len(r1)))
a = np.sort(res, order=["ref_id", "query_id"]); b = np.sort(r1, order=["ref_id", "query_id"])
print("same result set:", a.tobytes() == b.tobytes())
PY
cat $O/${TAG}_multi_c.txt | tail -8
