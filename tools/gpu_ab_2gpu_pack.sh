#!/bin/bash
# 2-GPU end-to-end A/B of the host packing variants (same box, same session)
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift; env "$@" SK_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-shuffle --spot-check 0 > $O/r2ab_$tag.json 2> $O/r2ab_$tag.err
  python - <<PY
import json
l=json.loads(open("$O/r2ab_$tag.json").read().strip().splitlines()[-1])
print("$tag", "value", round(l["ms_per_step"],1), "e2e", round(l["e2e"]["ms_per_step"],1), l["e2e"].get("step_ms_rank0"), l["e2e"]["host_pack_share"])
PY
  grep "packed on the host" $O/r2ab_$tag.err | tail -2 | cut -c1-170; }
run default X=1
run novbmi SK_PACK_NO_VBMI=1
run ascii SK_HOST_PACK=0
run f25 SK_HOST_PACK=0.25
run f50 SK_HOST_PACK=0.5
