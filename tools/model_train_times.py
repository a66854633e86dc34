"""Whole-network step times with the host CNN on PyTorch ops vs on libglrgtv's kernels (set_host_cnn_kernels), CUDA-event timed:
training step (reference loss, forward + backward, no optimiser) and no_grad inference, v13 configuration.
    python tools/model_train_times.py [--batch 4 --res 128 --steps 3]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M, train as T  # noqa: E402


def flag(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, R, n = flag("--batch", 4), flag("--res", 128), flag("--steps", 3)
    torch.manual_seed(0)
    model = T.build_model({"type": "AbtractMultiScaleGraphFilter"}).cuda().train()
    noisy, clean = torch.rand(B, R, R, 3, device="cuda"), torch.rand(B, R, R, 3, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(1)

    def step():
        model.zero_grad(set_to_none=True)
        loss, _ = T.reference_loss(model, noisy, clean, generator=gen)
        loss.backward()

    def infer():
        with torch.no_grad():
            model(noisy.permute(0, 3, 1, 2))

    res = {"batch": B, "res": R, "Mpix": B * R * R / 1e6}
    for name, on in (("torch", False), ("kernels", True)):
        M.set_host_cnn_kernels(on)
        res[f"train_ms_{name}"] = timed(step, n)
        res[f"infer_ms_{name}"] = timed(infer, n)
        res[f"peak_GB_{name}"] = torch.cuda.max_memory_allocated() / 2 ** 30
        torch.cuda.reset_peak_memory_stats()
    M.set_host_cnn_kernels(False)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
