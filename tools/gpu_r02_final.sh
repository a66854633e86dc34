#!/bin/bash
# Round-2 evidence run (one GPU): the whole GPU suite, smoke, the default bench line with its baselines, the ncu launch list and
# per-slot DRAM bytes of one step, `ncu --set full` captures (exported to CSV on the box), source-level stall samples of k_proj_tc.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputests_final.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"
( time python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2> gpurun_out/bench_final.time; echo "bench rc=$?"
B="python bench.py --steps 1 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-infer4k"
NK=${NK:-158}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -s $((3*NK)) -c $NK --csv --log-file gpurun_out/r02_step.csv $B > gpurun_out/ncu_a.log 2>&1; echo "ncuA rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_all.csv $B > gpurun_out/ncu_a2.log 2>&1; echo "ncuA2 rc=$?"
B1="python bench.py --steps 1 --warmup 1 --no-gpu-baseline --no-cpu-baseline --no-infer4k"
cap() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/$1 -f $B1 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
}
cap r02_bw2 k_bw2 30 10
cap r02_fwd "k_stream_fwd|k_fw2|k_gtv_coeffs" 0 8
cap r02_ww "k_weights_walk" 0 4
cap r02_proj k_proj_tc 0 7
ncu -i /tmp/r02_proj.ncu-rep --page source --csv --kernel-name regex:k_proj_tc --launch-skip 0 --launch-count 1 > gpurun_out/r02_proj_source.csv 2>/dev/null
ncu -i /tmp/r02_bw2.ncu-rep --page source --csv --launch-skip 1 --launch-count 1 > gpurun_out/r02_bw2_source.csv 2>/dev/null
ls -la gpurun_out/ | tail -20
tail -3 gpurun_out/gputests_final.log
