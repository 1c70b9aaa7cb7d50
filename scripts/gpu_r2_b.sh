#!/bin/bash
# Round 2, call B: GPU tests after the exp / bf16-storage changes + fp32 vs bf16 storage bench lines.
mkdir -p gpurun_out
TAG=${TAG:-r02b}
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -30 gpurun_out/${TAG}_gpu_tests.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
for st in fp32 bf16; do
timeout 300 python bench.py --steps 10 --warmup 3 --storage $st --no-cpu-baseline --no-render > gpurun_out/${TAG}_bench_$st.log 2>&1
tail -1 gpurun_out/${TAG}_bench_$st.log | cut -c1-200
done
timeout 300 python bench.py --workload cfg4 --storage bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/${TAG}_bench_cfg4_bf16.log 2>&1
tail -1 gpurun_out/${TAG}_bench_cfg4_bf16.log | cut -c1-200
