#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r01b}
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-4500
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-render --workload cfg4 > gpurun_out/${TAG}_bench_cfg4.log 2>&1
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_cfg4.log
tail -8 gpurun_out/${TAG}_bench_cfg4.log | cut -c1-3000
