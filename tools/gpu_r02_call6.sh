#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_2gpu.log 2>&1; echo "check rc=$?"
$TR --master-port 29512 tools/bench_infer4k.py --batched > gpurun_out/infer4k_2gpu_v2.json 2> gpurun_out/infer4k_2gpu_v2.err; echo rc=$?
$TR --master-port 29513 tools/bench_infer4k.py --batched --no-streams > gpurun_out/infer4k_2gpu_v2_nostreams.json 2>> gpurun_out/infer4k_2gpu_v2.err; echo rc=$?
$TR --master-port 29514 tools/bench_infer4k.py --batched --low-priority-nccl > gpurun_out/infer4k_2gpu_v2_lowprio.json 2>> gpurun_out/infer4k_2gpu_v2.err; echo rc=$?
$TR --master-port 29515 tools/bench_infer4k.py --batched --no-overlap > gpurun_out/infer4k_2gpu_v2_nooverlap.json 2>> gpurun_out/infer4k_2gpu_v2.err; echo rc=$?
$TR --master-port 29516 tools/bench_infer4k.py --batched --trace > gpurun_out/infer4k_2gpu_trace.json 2> gpurun_out/infer4k_2gpu_trace.err; echo rc=$?
python tools/strip_host_overhead.py --ranks 8 > gpurun_out/host_overhead_r8_v2.json 2> gpurun_out/host_overhead.err; echo rc=$?
$TR --master-port 29517 tools/bench_infer4k.py --model > gpurun_out/infer4k_model_2gpu.json 2> gpurun_out/infer4k_model.err; echo rc=$?
python tools/bench_infer4k.py --model > gpurun_out/infer4k_model_1gpu.json 2>> gpurun_out/infer4k_model.err; echo rc=$?
tail -4 gpurun_out/multi_gpu_check_2gpu.log
