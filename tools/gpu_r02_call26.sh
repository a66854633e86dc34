#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_proj_tc -s 1 -c 1 -o /tmp/r02_dgrad -f python tools/proj_times.py one 32,96,48,65536 > gpurun_out/ncu_dgrad.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/r02_dgrad.ncu-rep --page source --csv > gpurun_out/r02_dgrad_source.csv 2>/dev/null
ncu -i /tmp/r02_dgrad.ncu-rep --page raw --csv > gpurun_out/r02_dgrad_raw.csv 2>/dev/null
ls -la gpurun_out/r02_dgrad*
