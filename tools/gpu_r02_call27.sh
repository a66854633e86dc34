#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --no-gpu-baseline --no-cpu-baseline --no-infer4k > gpurun_out/bench_c27.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --no-gpu-baseline --no-cpu-baseline --no-infer4k --no-streams > gpurun_out/bench_c27_nostreams.log 2>&1; echo "bench rc=$?"
for f in bench_c27 bench_c27_nostreams; do tail -1 gpurun_out/$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"; done
