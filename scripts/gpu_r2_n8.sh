#!/bin/bash
# Round 2, 8-GPU call: the BASELINE configs that round 1 left unmeasured at N > 1 (VERDICT "+2", "+3"):
#   configs[1] cfg2_sh weak + strong scaling, configs[2] cfg2 + per-step blur, configs[3] cfg4 (LLFF) at 2/4/8,
#   configs[4] 200 views at 800x800 image-sharded over 8 GPUs.
mkdir -p gpurun_out
TAG=${TAG:-r02n8}
run() { # n name args...
  n=$1; name=$2; shift; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-render --no-cpu-baseline "$@" > gpurun_out/${TAG}_$name.log 2>&1
  tail -1 gpurun_out/${TAG}_$name.log > gpurun_out/${TAG}_$name.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$name.json"))
    st=d.get('step_ms_rank0') or [0]
    print("$name", d.get('n_gpus'), d.get('scaling'), 'value', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],3), 'first', st[0], 'steady', sorted(st)[len(st)//2], 'wait', (d.get('kernel_ms_per_step_data_parallel_rank0') or {}).get('grad_sync_wait'))
except Exception as e:
    print("$name failed", e)
PY
}
run 8 weak_cfg2sh_n8
run 8 strong_cfg2sh_n8 --scaling strong
run 8 blur_cfg2_n8 --workload cfg2 --blur 0.15
run 8 cfg4_n8 --workload cfg4
run 4 cfg4_n4 --workload cfg4
run 2 cfg4_n2 --workload cfg4
run 8 render_cfg2_n8 --mode render --frames 200 --workload cfg2
run 8 render_cfg2sh_n8 --mode render --frames 200 --workload cfg2_sh
