#!/bin/bash
# run one bench command: ARGS="..." TAG=...
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 $ARGS > gpurun_out/${TAG:-one}.log 2>&1
echo "bench rc=$?" >> gpurun_out/${TAG:-one}.log
tail -8 gpurun_out/${TAG:-one}.log | cut -c1-3500
