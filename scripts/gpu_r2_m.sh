#!/bin/bash
# Round 2, call M (2 GPUs): final-HEAD safety run: full GPU suite (incl. NCCL parity), smoke(), default N=2 bench.
mkdir -p gpurun_out
TAG=${TAG:-r02m}
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -5 gpurun_out/${TAG}_gpu_tests.log | cut -c1-300
cp gpurun_out/parity_errors.jsonl gpurun_out/${TAG}_parity_errors.jsonl 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_n2.log 2>&1
tail -1 gpurun_out/${TAG}_n2.log > gpurun_out/${TAG}_n2.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_n2.json")); st=d['step_ms_rank0']
    print("n2 default", round(d['ms_per_step'],3), 'first', st[0], 'steady', sorted(st)[len(st)//2], 'e2e', round(d['e2e']['ms_per_step'],3), 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline'].get('ncu',{}).get('commit'))
except Exception as e:
    print("n2 failed", e); print(open("gpurun_out/${TAG}_n2.log").read()[-1500:])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
