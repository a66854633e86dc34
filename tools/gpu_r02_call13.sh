#!/bin/bash
# forward quad-walker occupancy variants (128 registers: 16 warps per SM instead of 12) and source-level stall samples of the
# weight-gradient GEMM and a forward stage kernel
set -x
mkdir -p gpurun_out
python tools/fwd_stage_times.py --reps 5 > gpurun_out/fwd_var_default.json 2> gpurun_out/fwd_var.err; echo rc=$?
GLRGTV_LIB=variants/fwd_t256.so python tools/fwd_stage_times.py --reps 5 > gpurun_out/fwd_var_t256.json 2>> gpurun_out/fwd_var.err; echo rc=$?
GLRGTV_LIB=variants/fwd_t128.so python tools/fwd_stage_times.py --reps 5 > gpurun_out/fwd_var_t128.json 2>> gpurun_out/fwd_var.err; echo rc=$?
python tools/bench_infer4k.py --streams > gpurun_out/infer4k_var_default.json 2>> gpurun_out/fwd_var.err; echo rc=$?
GLRGTV_LIB=variants/fwd_t256.so python tools/bench_infer4k.py --streams > gpurun_out/infer4k_var_t256.json 2>> gpurun_out/fwd_var.err; echo rc=$?
GLRGTV_LIB=variants/fwd_t128.so python tools/bench_infer4k.py --streams > gpurun_out/infer4k_var_t128.json 2>> gpurun_out/fwd_var.err; echo rc=$?
B1="python bench.py --steps 1 --warmup 1 --no-gpu-baseline --no-cpu-baseline --no-infer4k --no-streams"
ncu --set full --clock-control none --import-source on -k regex:"k_proj_tc<\(int\)1>|k_proj_tcILi1" -s 0 -c 1 -o /tmp/r02_wgrad -f $B1 > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_wgrad.ncu-rep --page source --csv > gpurun_out/r02_wgrad_source.csv 2>/dev/null
ncu -i /tmp/r02_wgrad.ncu-rep --page raw --csv > gpurun_out/r02_wgrad_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_stream_fwd" -s 1 -c 1 -o /tmp/r02_sf -f $B1 > gpurun_out/ncu_sf.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_sf.ncu-rep --page source --csv > gpurun_out/r02_sf_source.csv 2>/dev/null
ls -la gpurun_out | tail -8
