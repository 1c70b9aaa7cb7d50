#!/bin/bash
# A/B over two env vars: VAR1/VALS1 x VAR2/VALS2
mkdir -p gpurun_out
for a in $VALS1; do for b in $VALS2; do
  echo "== $VAR1=$a $VAR2=$b" >> gpurun_out/${TAG:-ab2}.log
  env $VAR1=$a $VAR2=$b timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    k=d.get('kernel_ms_per_step') or {}
    print(round(d['ms_per_step'],3), k.get('vm_app_bwd'), k.get('vm_density_bwd'))
" >> gpurun_out/${TAG:-ab2}.log
done; done
cat gpurun_out/${TAG:-ab2}.log
