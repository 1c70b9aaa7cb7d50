#!/bin/bash
# One GPU-box round trip: parity tests, sanitizer on the smoke case, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
if [ "${SANITIZE:-1}" = "1" ]; then
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1
  echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
  tail -3 gpurun_out/sanitizer.log
fi
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --cpu-rays ${CPU_RAYS:-128} > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
