#!/bin/bash
# Round 2, final evidence: default bench line, reference arm, ncu launch list + `--set full` capture of the same code.
mkdir -p gpurun_out
TAG=${TAG:-r02final}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -2 gpurun_out/${TAG}_bench.log | head -1 > gpurun_out/${TAG}_bench.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print('bench', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'also', (d.get('also') or {}).get('ms_per_step'), 'strict', (d.get('strict_fp32') or {}).get('ms_per_step'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('render_800x800',{}).get('ms_per_frame'))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_ref.log 2>&1; tail -1 gpurun_out/${TAG}_ref.log | cut -c1-300
timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg4.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg4.log > gpurun_out/${TAG}_bench_cfg4.json; tail -1 gpurun_out/${TAG}_bench_cfg4.log | cut -c1-200
timeout 300 python bench.py --workload cfg2 --blur 0.15 --steps 20 --warmup 5 --no-cpu-baseline --no-render --no-also > gpurun_out/${TAG}_bench_cfg3.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg3.log > gpurun_out/${TAG}_bench_cfg3.json; tail -1 gpurun_out/${TAG}_bench_cfg3.log | cut -c1-200
timeout 300 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline --no-render --no-also > gpurun_out/${TAG}_bench_cfg1.log 2>&1; tail -1 gpurun_out/${TAG}_bench_cfg1.log > gpurun_out/${TAG}_bench_cfg1.json; tail -1 gpurun_out/${TAG}_bench_cfg1.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-breakdown --no-render --no-also > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_launches.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vm_scatter_walk|app_basis_fwd|vm_fwd_kernel|sh_bwd_data|head_bwd_wgrad|alpha_fwd|app_fill|composite_fwd|render_bwd|march" -s 52 -c 13 \
   -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-breakdown --no-render --no-also > gpurun_out/${TAG}_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wv_head|vm_scatter_walk|vm_fwd_kernel" -s 28 -c 7 \
   -o gpurun_out/${TAG}_prof_cfg4 -f python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline --no-breakdown --no-render --no-also > gpurun_out/${TAG}_prof_cfg4.log 2>&1
ls -la gpurun_out | tail -6
