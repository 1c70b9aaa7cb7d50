#!/bin/bash
# Round 2, call C: WeakView tensor-core head + parity after the test-helper fix; cfg4 bench on the tc head.
mkdir -p gpurun_out
TAG=${TAG:-r02c}
rm -f gpurun_out/parity_errors.jsonl
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "weakview_on_reference" > gpurun_out/${TAG}_wv.log 2>&1
rc=$?; tail -15 gpurun_out/${TAG}_wv.log
if [ $rc -ne 0 ]; then
  timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "weakview_on_reference and not blur" > gpurun_out/${TAG}_wv_memcheck.log 2>&1
  grep -E "Invalid|Error|at 0x|by thread" gpurun_out/${TAG}_wv_memcheck.log | head -20
fi
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -30 gpurun_out/${TAG}_gpu_tests.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
for hd in fp32 auto; do
timeout 300 python bench.py --workload cfg4 --head $hd --steps 10 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/${TAG}_bench_cfg4_$hd.log 2>&1
tail -1 gpurun_out/${TAG}_bench_cfg4_$hd.log | cut -c1-200
done
