#!/bin/bash
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29512 tools/bench_infer4k.py --batched --trace > gpurun_out/infer4k_2gpu_trace.json 2> gpurun_out/infer4k_2gpu_trace.err; echo rc=$?
python tools/host_cnn_times.py > gpurun_out/host_cnn_times_r02_s0.json 2> gpurun_out/host_cnn_times.err; echo rc=$?
python tools/host_cnn_times.py --dim 384 --hidden 768 --rows 270 --cols 480 > gpurun_out/host_cnn_times_r02_s3.json 2>> gpurun_out/host_cnn_times.err; echo rc=$?
python -m pytest tests/test_gpu_host_cnn.py tests/test_gpu_model.py -x -q > gpurun_out/gputests_c5.log 2>&1; echo "pytest rc=$?"
ls -la gpurun_out | tail -5
