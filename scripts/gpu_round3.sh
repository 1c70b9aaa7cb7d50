#!/bin/bash
# One GPU-box round trip: parity tests, default bench (cfg2_sh + cfg2 "also"), ncu launch list of the same command.
mkdir -p gpurun_out
TAG=${TAG:-r01}
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -2 gpurun_out/${TAG}_bench.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-300} -c 120 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-breakdown --no-render --no-also > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_launches.csv | cut -c1-200
