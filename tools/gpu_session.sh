#!/bin/bash
# One GPU-box session (run under gpurun): parity tests, a short bench, per-kernel event times incl. A/B variants, ncu launch
# list + one full capture.  Everything lands in gpurun_out/<tag>_*.  Usage: tools/gpu_session.sh <tag> [quick]
TAG=${1:-s}
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/${TAG}_env.txt; nproc >> $O/${TAG}_env.txt; cat /sys/fs/cgroup/cpu.max >> $O/${TAG}_env.txt; free -g | head -2 >> $O/${TAG}_env.txt
lscpu | grep -i "model name\|^CPU(s)\|flags" | cut -c1-400 >> $O/${TAG}_env.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/${TAG}_tests.log
tail -5 $O/${TAG}_tests.log
SK_TRACE=1 timeout 900 python bench.py --steps 3 --warmup 2 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -c 2500 $O/${TAG}_bench.json
grep "sk_sketch_batch\|sk_triangle\]" $O/${TAG}_bench.err | tail -8
timeout 600 python tools/profile_step.py 400 > $O/${TAG}_profile_step.txt 2>&1
SK_DP_GL=8 timeout 600 python tools/profile_step.py 400 > $O/${TAG}_profile_step_gl8.txt 2>&1
head -24 $O/${TAG}_profile_step.txt; grep "dp_kernel" $O/${TAG}_profile_step_gl8.txt | head -3
if [ "$2" != "quick" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline --spot-check 0 > $O/${TAG}_bench_under_ncu.log 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"pack_kernel|hashpass|expand_kernel|hash_build|probe_kernel|chunk_fast|anchor_kernel|dp_group|select_kernel|chunkstat|final_kernel" -c 21 -f -o $O/${TAG}_full python tools/profile_step.py 200 > $O/${TAG}_ncu_full.log 2>&1
  ls -la $O/${TAG}_full.ncu-rep
fi
if [ "$2" == "configs" ]; then
  timeout 900 python bench.py --config c2 --steps 3 --warmup 2 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; tail -c 600 $O/${TAG}_bench_c2.json
  timeout 1200 python bench.py --config c5 --steps 2 --warmup 1 --cpu-sample 2000 > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err; tail -c 600 $O/${TAG}_bench_c5.json; tail -3 $O/${TAG}_bench_c5.err
  SK_PROBE_TMA=0 timeout 1200 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline --spot-check 0 > $O/${TAG}_bench_c5_notma.json 2> $O/${TAG}_bench_c5_notma.err
  timeout 1200 python bench.py --config dense --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG}_bench_dense.json 2> $O/${TAG}_bench_dense.err; tail -c 600 $O/${TAG}_bench_dense.json
  timeout 1500 python tools/bench_search.py --refs 6500 --queries 1000 > $O/${TAG}_bench_search.json 2> $O/${TAG}_bench_search.err; tail -c 900 $O/${TAG}_bench_search.json; tail -3 $O/${TAG}_bench_search.err
fi
