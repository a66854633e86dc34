#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_c17.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_c17.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_c17.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/gputests_c17.log; tail -1 gpurun_out/smoke_c17.log
