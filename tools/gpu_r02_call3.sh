#!/bin/bash
# Round-2 GPU call 3: edge-weight walkers - parity tests, the whole GPU suite, stage times and a bench line
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_weights.py -x -q > gpurun_out/gpu_weights_tests.log 2>&1; echo "weights rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/gputests_c3.log 2>&1; echo "pytest rc=$?"
python tools/fwd_stage_times.py --reps 5 --bwd > gpurun_out/stage_times_c3.json 2> gpurun_out/stage_times_c3.err; echo rc=$?
python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench rc=$?"
python tools/bench_infer4k.py > gpurun_out/infer4k_c3.json 2> gpurun_out/infer4k_c3.err; echo rc=$?
B1="python bench.py --steps 1 --warmup 1 --no-gpu-baseline --no-cpu-baseline --no-infer4k"
ncu --set full --clock-control none --import-source on -k regex:"k_weights_walk" -s 0 -c 4 -o /tmp/r02_ww -f $B1 > gpurun_out/ncu_ww.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_ww.ncu-rep --page raw --csv > gpurun_out/r02_ww_raw.csv 2>/dev/null
tail -3 gpurun_out/gpu_weights_tests.log gpurun_out/gputests_c3.log
