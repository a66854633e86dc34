#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_head.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_head.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/gputests_head.log; tail -1 gpurun_out/smoke_head.log
