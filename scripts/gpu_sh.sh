#!/bin/bash
# SH tensor-core path: parity tests + cfg2_sh bench (short)
mkdir -p gpurun_out
TAG=${TAG:-r01sh}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "sh or tensor_core" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -25 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload cfg2_sh > gpurun_out/${TAG}_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -3 gpurun_out/${TAG}_bench.log | cut -c1-3000
