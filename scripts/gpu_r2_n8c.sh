#!/bin/bash
# Round 2, third 8-GPU call: split21 data-parallel backward at N=8 (single-launch scheme: 3.229 ms, wait 0.366)
mkdir -p gpurun_out
TAG=${TAG:-r02n8c}
run() { n=$1; name=$2; shift; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-render --no-cpu-baseline --no-breakdown "$@" > gpurun_out/${TAG}_$name.log 2>&1
  tail -1 gpurun_out/${TAG}_$name.log > gpurun_out/${TAG}_$name.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$name.json")); st=d.get('step_ms_rank0') or [0]
    print("$name", 'ms', round(d['ms_per_step'],3), 'first', st[0], 'steady', sorted(st)[len(st)//2], (d.get('kernel_ms_per_step_data_parallel_rank0') or {}))
except Exception as e:
    print("$name failed", e)
PY
}
run 8 split21 --sync-split21
run 8 single
