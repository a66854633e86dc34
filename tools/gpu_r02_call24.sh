#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-infer4k > gpurun_out/bench_c24.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_c24.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"
