#!/bin/bash
# Round 2, call H (1 GPU): parity after the double-precision / 4-samples-per-lane compositing kernels + bench lines.
mkdir -p gpurun_out
TAG=${TAG:-r02k}
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -12 gpurun_out/${TAG}_gpu_tests.log | cut -c1-900
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
timeout 300 python scripts/debug_cfg4_grad.py 0.4 > gpurun_out/${TAG}_debug_cfg4.log 2>&1; grep -E "density_plane|density_line|per-sample|coherence|my weights" gpurun_out/${TAG}_debug_cfg4.log
for wl in cfg2_sh cfg4; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-render --no-also > gpurun_out/${TAG}_bench_$wl.log 2>&1
tail -1 gpurun_out/${TAG}_bench_$wl.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$wl', round(d['ms_per_step'],3), d['kernel_ms_per_step'])"
done
