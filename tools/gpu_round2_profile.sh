#!/bin/bash
# Round-2 GPU evidence run (one gpurun call): GPU tests, a bench line, the ncu launch list of one bench step, and `ncu --set full`
# captures of the stage / weights / projection kernels, exported to CSV on the box (the .ncu-rep files are too large to bring back).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputests19.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench19.log 2>&1; echo "bench rc=$?"
python tools/bench_config1.py > gpurun_out/config1_r02.log 2>&1; echo "config1 rc=$?"
B="python bench.py --steps 1 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-infer4k"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"^k_" -s 492 -c 164 --csv --log-file gpurun_out/r02_step.csv $B > gpurun_out/ncu_a.log 2>&1; echo "ncuA rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_all.csv $B > gpurun_out/ncu_a2.log 2>&1; echo "ncuA2 rc=$?"
B1="python bench.py --steps 1 --warmup 1 --no-gpu-baseline --no-cpu-baseline --no-infer4k"
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/$1 -f $B1 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  rm -f /tmp/$1.ncu-rep
}
cap r02_bw2 k_bw2 30 10
cap r02_fwd "k_stream_fwd|k_block_weights|k_gtv_coeffs" 0 10
cap r02_proj k_proj_tc 0 7
ls -la gpurun_out/
