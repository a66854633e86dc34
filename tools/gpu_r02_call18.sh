#!/bin/bash
# source-level stall samples of the weight-gradient GEMM (k_proj_tc<1>) on the scale-0 shape
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_proj_tc -s 2 -c 1 -o /tmp/r02_wgrad -f python tools/proj_times.py one 32,96,48,65536 > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_wgrad.ncu-rep --page source --csv > gpurun_out/r02_wgrad_source.csv 2>/dev/null
ncu -i /tmp/r02_wgrad.ncu-rep --page raw --csv > gpurun_out/r02_wgrad_raw.csv 2>/dev/null
ls -la gpurun_out/r02_wgrad*; tail -3 gpurun_out/ncu_wgrad.log
