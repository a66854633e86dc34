#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_c22.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_c22.log 2>&1; echo "smoke rc=$?"
timeout 300 python tools/model_train_times.py --batch 4 --res 256 --steps 3 > gpurun_out/model_train_times_c22.json 2>/dev/null; echo rc=$?
tail -2 gpurun_out/gputests_c22.log; cat gpurun_out/model_train_times_c22.json
