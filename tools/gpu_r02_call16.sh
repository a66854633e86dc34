#!/bin/bash
# host CNN training path on the tensor-core GEMMs: parity tests, whole-network step times (PyTorch modules vs kernels)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host_cnn_train.py tests/test_gpu_host_cnn.py tests/test_gpu_train.py -x -q > gpurun_out/gputests_c16.log 2>&1; echo "pytest rc=$?"
timeout 400 python tools/model_train_times.py --batch 4 --res 128 --steps 3 > gpurun_out/model_train_times_b4_128.json 2> gpurun_out/model_train_times.err; echo rc=$?
timeout 400 python tools/model_train_times.py --batch 4 --res 256 --steps 3 > gpurun_out/model_train_times_b4_256.json 2>> gpurun_out/model_train_times.err; echo rc=$?
tail -3 gpurun_out/gputests_c16.log; cat gpurun_out/model_train_times_b4_128.json gpurun_out/model_train_times_b4_256.json
