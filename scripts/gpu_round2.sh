#!/bin/bash
# One GPU-box round trip: parity tests, bench on every workload (cfg2 default, cfg3 = blur, cfg1, cfg4), ncu launch list.
mkdir -p gpurun_out
TAG=${TAG:-r01}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -5 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-600
for W in "cfg2 --blur 0.15" "cfg1" "cfg4" "cfg2_sh"; do
  N=$(echo $W | tr -d ' .-')
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-render --workload $W > gpurun_out/${TAG}_bench_${N}.log 2>&1
  echo "bench rc=$?" >> gpurun_out/${TAG}_bench_${N}.log
  tail -2 gpurun_out/${TAG}_bench_${N}.log | cut -c1-400
done
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-300} -c 120 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-breakdown --no-render > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
fi
