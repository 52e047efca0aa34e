#!/bin/bash
# Short GPU-box session: parity tests, the default bench, A/B bench runs for the environment switches given as arguments
# ("NAME=VALUE" each; "P:NAME=VALUE" = A/B of the per-kernel event times instead), per-kernel event times.
# Usage: tools/gpu_session_ab.sh <tag> [ENV=VAL | P:ENV=VAL ...]
TAG=${1:-s}; shift
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/${TAG}_env.txt; nproc >> $O/${TAG}_env.txt; cat /sys/fs/cgroup/cpu.max >> $O/${TAG}_env.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/${TAG}_tests.log
tail -5 $O/${TAG}_tests.log
SK_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 2 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -c 2500 $O/${TAG}_bench.json
grep "sk_triangle\] worker" $O/${TAG}_bench.err | tail -12
for kv in "$@"; do
  case $kv in P:*) continue;; esac
  n=$(echo $kv | tr '=' '_')
  env $kv timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --spot-check 0 > $O/${TAG}_bench_${n}.json 2> $O/${TAG}_bench_${n}.err
  echo "== $kv"; python - <<PY
import json
l=json.loads(open("$O/${TAG}_bench_${n}.json").read().strip().splitlines()[-1])
print("value ms", l["ms_per_step"], "e2e", l["e2e"])
PY
done
timeout 600 python tools/profile_step.py 400 > $O/${TAG}_profile_step.txt 2>&1
head -24 $O/${TAG}_profile_step.txt
for kv in "$@"; do          # "P:NAME=VALUE": per-kernel event times with that switch
  case $kv in P:*) ;; *) continue;; esac
  kv=${kv#P:}; n=$(echo $kv | tr '=' '_')
  env $kv timeout 600 python tools/profile_step.py 400 > $O/${TAG}_profile_step_${n}.txt 2>&1
  echo "== $kv"; grep "timing=True\|probe_kernel\|anchor_kernel\|chunk_fast\|dp_kernel" $O/${TAG}_profile_step_${n}.txt
done
