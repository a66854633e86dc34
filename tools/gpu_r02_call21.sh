#!/bin/bash
set -x
mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err ) 2> gpurun_out/bench_final2.time; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final2.json 2>/dev/null; echo "ref rc=$?"
tail -1 gpurun_out/bench_final2.json | cut -c1-300; tail -3 gpurun_out/bench_final2.time
