#!/bin/bash
# Round 2, call F: tests after the capacity tracker / ABI changes, cfg4 gradient diagnostic, bench lines.
mkdir -p gpurun_out
TAG=${TAG:-r02f}
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -22 gpurun_out/${TAG}_gpu_tests.log | cut -c1-500
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
timeout 600 python scripts/debug_cfg4_grad.py 0.4 > gpurun_out/${TAG}_debug_cfg4.log 2>&1; tail -28 gpurun_out/${TAG}_debug_cfg4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for wl in cfg2_sh cfg4; do
timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_$wl.log 2>&1
tail -1 gpurun_out/${TAG}_bench_$wl.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$wl', round(d['ms_per_step'],3), d['kernel_ms_per_step'], d.get('render_800x800'), (d.get('also') or {}).get('ms_per_step'))"
done
