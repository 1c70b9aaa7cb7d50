#!/bin/bash
# Round 2, call L (2 GPUs): split21 data-parallel backward (planes 0+1 | plane 2 | density) vs the single-launch scheme.
mkdir -p gpurun_out
TAG=${TAG:-r02l}
run() { name=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-render --no-cpu-baseline --no-breakdown "$@" > gpurun_out/${TAG}_n2_$name.log 2>&1
  tail -1 gpurun_out/${TAG}_n2_$name.log > gpurun_out/${TAG}_n2_$name.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_n2_$name.json")); st=d['step_ms_rank0']
    print("$name", round(d['ms_per_step'],3), 'first', st[0], 'steady', sorted(st)[len(st)//2], d.get('kernel_ms_per_step_data_parallel_rank0'))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/${TAG}_n2_$name.log").read()[-1200:])
PY
}
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3
run split21
run single --sync-single
run cfg4_split21 --workload cfg4
