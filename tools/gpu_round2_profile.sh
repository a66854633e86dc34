set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gputests19.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench19.log 2>&1; echo "bench rc=$?"
B="python bench.py --steps 1 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-infer4k"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"^k_" -s 492 -c 164 --csv --log-file gpurun_out/r02_step.csv $B > gpurun_out/ncu_a.log 2>&1; echo "ncuA rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 900 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_a2.log 2>&1; echo "ncuA2 rc=$?"
B1="python bench.py --steps 1 --warmup 1 --no-gpu-baseline --no-cpu-baseline --no-infer4k"
ncu --set full --clock-control none --import-source on -k regex:k_bw2 -s 30 -c 10 -o gpurun_out/r02_bw2 -f $B1 > gpurun_out/ncu_b.log 2>&1; echo "ncuB rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_stream_fwd|k_block_weights|k_gtv_coeffs" -c 10 -o gpurun_out/r02_fwd -f $B1 > gpurun_out/ncu_c.log 2>&1; echo "ncuC rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_proj_tc" -c 7 -o gpurun_out/r02_proj -f $B1 > gpurun_out/ncu_d.log 2>&1; echo "ncuD rc=$?"
ls -la gpurun_out/
