#!/bin/bash
# Round 2, call P (4 GPUs): the driver's N=4 default bench (split backward engages at 4+ ranks).
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --no-render --no-cpu-baseline > gpurun_out/r02p_n4.log 2>&1
grep '^{"metric"' gpurun_out/r02p_n4.log | tail -1 > gpurun_out/r02p_n4.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02p_n4.json")); st=d['step_ms_rank0']
    print("n4", round(d['ms_per_step'],3), 'first', st[0], 'steady', sorted(st)[len(st)//2], 'value', round(d['value']), d.get('kernel_ms_per_step_data_parallel_rank0'))
except Exception as e:
    print("n4 failed", e); print(open("gpurun_out/r02p_n4.log").read()[-1500:])
PY
