#!/bin/bash
# Round 2, call J (2 GPUs): CUDA-graph replay of the step -- test, launch-bound regimes (512 rays, cfg1), N=2 strong.
mkdir -p gpurun_out
TAG=${TAG:-r02j}
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "cuda_graph or capacity or weakview_on_reference" 2>&1 | tail -8 | cut -c1-400
b1() { name=$1; shift
  timeout 200 python bench.py --steps 20 --warmup 5 --no-render --no-cpu-baseline --no-also --no-breakdown "$@" > gpurun_out/${TAG}_$name.log 2>&1
  tail -1 gpurun_out/${TAG}_$name.log > gpurun_out/${TAG}_$name.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$name.json")); print("$name", 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'host', d['host_enqueue_ms_per_step'], 'launches', d['gpu_launches'], 'value', round(d['value']))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/${TAG}_$name.log").read()[-1500:])
PY
}
b1 n1_eager
b1 n1_graph --cuda-graph
b1 n1_512_eager --rays 512
b1 n1_512_graph --rays 512 --cuda-graph
b1 cfg1_eager --workload cfg1
b1 cfg1_graph --workload cfg1 --cuda-graph
b2() { name=$1; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-render --no-cpu-baseline --no-breakdown "$@" > gpurun_out/${TAG}_$name.log 2>&1
  tail -1 gpurun_out/${TAG}_$name.log > gpurun_out/${TAG}_$name.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$name.json")); print("$name", 'ms', round(d['ms_per_step'],3), 'host', d['host_enqueue_ms_per_step'], 'value', round(d['value']))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/${TAG}_$name.log").read()[-1500:])
PY
}
b2 n2_strong_eager --scaling strong
b2 n2_strong_graph --scaling strong --cuda-graph
b2 n2_weak_graph --cuda-graph
