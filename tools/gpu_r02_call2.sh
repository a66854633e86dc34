#!/bin/bash
# Round-2 GPU call 2: forward generations side by side, 4K inference, one correctly windowed ncu step, configs 3 and 5, the default bench line.
set -x
mkdir -p gpurun_out
python tools/fwd_stage_times.py --reps 5 > gpurun_out/fwd_times_gen1.json 2> gpurun_out/fwd_times_gen1.err; echo rc=$?
python tools/fwd_stage_times.py --reps 5 --fw2 > gpurun_out/fwd_times_fw2.json 2> gpurun_out/fwd_times_fw2.err; echo rc=$?
python tools/bench_infer4k.py > gpurun_out/infer4k_gen1.json 2> gpurun_out/infer4k_gen1.err; echo rc=$?
python tools/bench_infer4k.py --fw2 > gpurun_out/infer4k_fw2.json 2> gpurun_out/infer4k_fw2.err; echo rc=$?
B="python bench.py --steps 1 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-infer4k"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -s 444 -c 148 --csv --log-file gpurun_out/r02_step.csv $B > gpurun_out/ncu_a.log 2>&1; echo "ncuA rc=$?"
timeout 300 python tools/bench_config3.py > gpurun_out/config3_v7_1gpu.json 2> gpurun_out/config3_v7.err; echo rc=$?
timeout 300 python tools/bench_config3.py --v1 > gpurun_out/config3_v1_1gpu.json 2> gpurun_out/config3_v1.err; echo rc=$?
timeout 600 python tools/bench_config5.py > gpurun_out/config5.jsonl 2> gpurun_out/config5.err; echo rc=$?
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time; echo rc=$?
( time python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2> gpurun_out/bench_reference.time; echo rc=$?
ls -la gpurun_out/
