#!/bin/bash
# Round 2, call G (2 GPUs): double-precision compositing scans -> parity; data-parallel variants at N=2.
mkdir -p gpurun_out
TAG=${TAG:-r02g}
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -12 gpurun_out/${TAG}_gpu_tests.log | cut -c1-700
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
timeout 600 python scripts/debug_cfg4_grad.py 0.4 > gpurun_out/${TAG}_debug_cfg4.log 2>&1; grep -E "density_plane|per-sample|coherence|my weights" gpurun_out/${TAG}_debug_cfg4.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-render --no-cpu-baseline --no-also > gpurun_out/${TAG}_n1.log 2>&1; tail -1 gpurun_out/${TAG}_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('n1', round(d['ms_per_step'],3), d['kernel_ms_per_step'])"
run() { name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-render --no-cpu-baseline "$@" > gpurun_out/${TAG}_n2_$name.log 2>&1
  tail -1 gpurun_out/${TAG}_n2_$name.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.readline()); print('$name', round(d['ms_per_step'],3), 'first', d['step_ms_rank0'][0], 'steady', sorted(d['step_ms_rank0'])[len(d['step_ms_rank0'])//2], d.get('kernel_ms_per_step_data_parallel_rank0'))
except Exception as e: print('$name failed', e)"
}
run default
run reserve16 --sync-reserve-sms 16 --nccl-max-ctas 16
run reserve8 --sync-reserve-sms 8 --nccl-max-ctas 8
run reserve24 --sync-reserve-sms 24
