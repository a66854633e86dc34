#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_c20.log 2>&1; echo "pytest rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_c20.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/gputests_c20.log; tail -1 gpurun_out/bench_c20.log | cut -c1-400
