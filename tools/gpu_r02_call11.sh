#!/bin/bash
# lean epilogue of the projection GEMM: accuracy, timings of both pipelines, bench, full GPU suite
set -x
mkdir -p gpurun_out
PROJ_PIPELINE=0 timeout 600 python tools/proj_times.py check > gpurun_out/proj_check_c.log 2>&1; echo "check rc=$?"
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k projection > gpurun_out/gputests_proj.log 2>&1; echo "pytest rc=$?"
PROJ_PIPELINE=0 timeout 300 python tools/proj_times.py > gpurun_out/proj_times_inplace_c.jsonl 2> gpurun_out/proj_times.err; echo rc=$?
PROJ_PIPELINE=1 timeout 300 python tools/proj_times.py > gpurun_out/proj_times_ring_c.jsonl 2>> gpurun_out/proj_times.err; echo rc=$?
timeout 300 python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/bench_c11.log 2>&1; echo "bench rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_c11.log 2>&1; echo "pytest all rc=$?"
timeout 300 python tools/host_cnn_times.py > gpurun_out/host_cnn_times_c11_s0.json 2> /dev/null; echo rc=$?
cat gpurun_out/proj_check_c.log | cut -c1-200
