#!/bin/bash
# Round 2, call O (2 GPUs): the data-parallel step replayed from a CUDA graph (NCCL all-reduces captured).
mkdir -p gpurun_out
TAG=${TAG:-r02o}
run() { n=$1; name=$2; shift; shift
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-render --no-cpu-baseline --no-breakdown "$@" > gpurun_out/${TAG}_$name.log 2>&1
  grep '^{"metric"' gpurun_out/${TAG}_$name.log | tail -1 > gpurun_out/${TAG}_$name.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$name.json")); st=d.get('step_ms_rank0') or [0]
    print("$name", 'ms', round(d['ms_per_step'],3), 'first', st[0], 'steady', sorted(st)[len(st)//2], 'host', d['host_enqueue_ms_per_step'], 'value', round(d['value']))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/${TAG}_$name.log").read()[-800:])
PY
}
N=${NG:-2}
run $N strong_graph_n$N --scaling strong --cuda-graph
run $N weak_graph_n$N --cuda-graph
