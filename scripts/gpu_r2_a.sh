#!/bin/bash
# Round 2, call A: all GPU tests (new full-size cfg3/cfg4 parity), smoke, headline bench with the new roofline fields,
# cfg4 / blur baselines, and an `ncu --set full` capture of one cfg2_sh step for profiles/ncu_kernels.json.
mkdir -p gpurun_out
TAG=${TAG:-r02a}
rm -f gpurun_out/parity_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -25 gpurun_out/${TAG}_gpu_tests.log
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log; tail -4 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/${TAG}_bench.log
tail -2 gpurun_out/${TAG}_bench.log | cut -c1-600
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/${TAG}_bench_cfg4.log 2>&1
tail -1 gpurun_out/${TAG}_bench_cfg4.log | cut -c1-300
timeout 300 python bench.py --workload cfg2 --blur 0.15 --steps 10 --warmup 3 --no-cpu-baseline --no-render --no-also > gpurun_out/${TAG}_bench_cfg3.log 2>&1
tail -1 gpurun_out/${TAG}_bench_cfg3.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vm_scatter_walk|app_basis_fwd|vm_fwd_kernel|sh_bwd_data|head_bwd_wgrad|alpha_fwd|composite_fwd|render_bwd|march" -s 44 -c 12 \
   -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-breakdown --no-render --no-also > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out | tail -5
