#!/bin/bash
# Round 2, call D (2 GPUs): cfg4 gradient diagnostic, NCCL parity tests, data-parallel bench variants at N=2.
mkdir -p gpurun_out
TAG=${TAG:-r02d}
timeout 600 python scripts/debug_cfg4_grad.py 0.4 > gpurun_out/${TAG}_debug_cfg4.log 2>&1; tail -12 gpurun_out/${TAG}_debug_cfg4.log
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_dist_tests.log 2>&1; tail -3 gpurun_out/${TAG}_dist_tests.log
run() { # name, args...
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-render --no-cpu-baseline "$@" > gpurun_out/${TAG}_n2_$name.log 2>&1
  tail -1 gpurun_out/${TAG}_n2_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('$name', round(d['ms_per_step'],3), 'steady', sorted(d['step_ms_rank0'])[len(d['step_ms_rank0'])//2], d.get('kernel_ms_per_step_data_parallel_rank0'))
except Exception as e: print('$name failed', e)"
}
timeout 300 python bench.py --steps 20 --warmup 5 --no-render --no-cpu-baseline --no-also > gpurun_out/${TAG}_n1.log 2>&1; tail -1 gpurun_out/${TAG}_n1.log | cut -c1-160
run default
run ctas8 --nccl-max-ctas 8
run ctas4 --nccl-max-ctas 4
run strong --scaling strong
run cfg4 --workload cfg4
run blur --workload cfg2 --blur 0.15
