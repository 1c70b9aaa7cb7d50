#!/bin/bash
# Round 2, call I (2 GPUs): data-parallel variants at N=2 (every run bounded to 150 s).
mkdir -p gpurun_out
TAG=${TAG:-r02i}
run() { name=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-render --no-cpu-baseline "$@" > gpurun_out/${TAG}_n2_$name.log 2>&1
  tail -1 gpurun_out/${TAG}_n2_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('$name', round(d['ms_per_step'],3), 'first', d['step_ms_rank0'][0], 'steady', sorted(d['step_ms_rank0'])[len(d['step_ms_rank0'])//2], d.get('kernel_ms_per_step_data_parallel_rank0'))
except Exception as e: print('$name failed', e)"
}
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3
run default
run reserve16 --sync-reserve-sms 16 --nccl-max-ctas 16
run reserve8 --sync-reserve-sms 8 --nccl-max-ctas 8
run reserve24 --sync-reserve-sms 24
