"""BASELINE config 3: the older family (MultiScaleSequenceDenoiser, model_GLR_GTV_deep_v7) trained on 64x64 patches, batch 4 per
GPU, sigma 15 (experiment_conf/example.yaml:15-23), batch-sharded over N GPUs with one gradient all-reduce per step.

    python tools/bench_config3.py [--steps 10 --warmup 3]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tools/bench_config3.py

One step = forward + L1 loss + backward + all-reduce + Adam step on synthetic patches already on the device.  Prints one JSON
line: whole-job Mpix/s (max over ranks of the CUDA-event time)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagerestoration_development_unrolling_b200 import shard, train as T  # noqa: E402


def flag(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    steps, warmup, B, R = flag("--steps", 10), flag("--warmup", 3), flag("--batch", 4), flag("--res", 64)
    torch.manual_seed(0)
    family = "v1" if "--v1" in sys.argv else "v7"                  # --v1: the three-block chain
    model = T.build_model({"type": "MultiScaleSequenceDenoiser" + ("_v1" if family == "v1" else "")}, dev).to(dev).train()
    opt, sched = T.build_optimizer(model)
    params = [p for p in model.parameters() if p.requires_grad]
    data = T.SyntheticNoisyPatches(patch_size=R, lambda_noise=15.0, max_num_patchs=B * world * 4)
    batches = []
    for k in range(4):                                            # four resident batches, cycled (64x64 patches live in L2 by design)
        items = [data[(k * world + rank) * B + i] for i in range(B)]
        batches.append((torch.stack([a for a, _ in items]).to(dev), torch.stack([b for _, b in items]).to(dev)))
    flat = None

    def step(k):
        nonlocal flat
        noisy, clean = batches[k % len(batches)]
        opt.zero_grad(set_to_none=True)
        loss, _ = T.reference_loss(model, noisy, clean)
        loss.backward()
        if world > 1:
            flat = shard.allreduce_gradients(params, average=True, flat=flat)
        opt.step()
        sched.step()
        return loss

    for k in range(warmup):
        step(k)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        loss = step(k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        dist.barrier()
    if rank == 0:
        print(json.dumps({"metric": "train_Mpix_per_s", "value": world * B * R * R / ms / 1e3, "unit": "Mpix/s", "n_gpus": world, "steps": steps,
                          "warmup": warmup, "ms_per_step": ms, "scaling": "weak", "dtype": "f32", "data": "synthetic", "loss": float(loss),
                          "config": {"workload": f"{family} MultiScaleSequenceDenoiser, {B} x 3 x {R} x {R} per GPU, sigma 15, fwd + bwd + all-reduce + Adam",
                                     "params": sum(p.numel() for p in params)}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
