#!/bin/bash
# A/B a tuning environment variable: VAR=name VALS="0 1 2" bash scripts/gpu_ab.sh
mkdir -p gpurun_out
for v in $VALS; do
  echo "== $VAR=$v" >> gpurun_out/${TAG:-ab}.log
  env $VAR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(round(d['ms_per_step'],3), d.get('kernel_ms_per_step'))
" >> gpurun_out/${TAG:-ab}.log
done
cat gpurun_out/${TAG:-ab}.log
