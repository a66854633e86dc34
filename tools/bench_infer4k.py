"""BASELINE config 4: 4K (3840x2160) full-image inference of the hot path, spatially sharded.

    python tools/bench_infer4k.py                                   # one GPU, whole image
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/bench_infer4k.py

The four LocalLowpassFilteringBlock of the v13 model run (no_grad) on the feature maps of one 3840x2160 image:
[1,48,2160,3840], [1,96,1080,1920], [1,192,540,960], [1,384,270,480].  With N ranks every map is cut into N row
strips (boundaries at even rows) and each block does one 26-row NCCL halo exchange (shard.sharded_block_forward).
`--staged`: one 8-row exchange per solver stage instead; `--batched`: the same with the four scales in lock-step and one
batched exchange per round (shard.sharded_filtering_staged).  `--model`: the WHOLE v13 network (host CNN on strips with one
exchanged row per 3x3 convolution + the staged filter blocks, shard.ShardedMultiScaleFilter) on the 3-channel image.
`--torch-cnn`: with --model, keep the LocalNonLinearBlocks on the PyTorch modules instead of host_cnn.py's kernels.
`--steps K --warmup W` (default 5 / 3).
Prints one JSON line: Mpix/s of the network input (8.29 Mpix per image), max over ranks of the CUDA-event time."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M  # noqa: E402
from imagerestoration_development_unrolling_b200 import shard  # noqa: E402

DIMS, NG, H0, W0 = [48, 96, 192, 384], [8, 16, 16, 32], 2160, 3840


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream="--low-priority-nccl" not in sys.argv)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    if "--fw2" in sys.argv:               # forward stages on the pair walkers (csrc/fw2.cuh)
        from imagerestoration_development_unrolling_b200 import _lib as L
        L.load().glrgtv_set_fwd_kernels(2)
    blocks = [M.LocalLowpassFilteringBlock(d, 1, g).to(dev) for d, g in zip(DIMS, NG)]
    strips = []
    for s, d in enumerate(DIMS if "--model" not in sys.argv else []):
        H, W = H0 >> s, W0 >> s
        a, b = shard.strip_bounds(H, world, align=2)[rank]
        x = shard.strip_with_halo_room((1, d, b - a, W), rank, world, device=dev) if "--batched" in sys.argv else torch.empty(1, d, b - a, W, device=dev)
        x.copy_(torch.randn(1, d, b - a, W, device=dev, generator=torch.Generator(device=dev).manual_seed(s)))
        strips.append(x)

    staged = "--staged" in sys.argv       # one 8-row exchange per solver stage instead of one 26-row exchange per block
    whole = "--model" in sys.argv

    def flag(name, default):
        return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default

    steps, warmup = flag("--steps", 5), flag("--warmup", 3)
    if whole:
        torch.manual_seed(0)
        net = M.AbtractMultiScaleGraphFilter(dims=DIMS, hidden_dims=[96, 192, 384, 768], nsubnets=[1, 1, 1, 1], ngraphs=NG,
                                             num_blocks=[4, 6, 6, 8], num_blocks_out=4).to(dev).eval()
        a, b = shard.strip_bounds(H0, world, align=16)[rank]
        img = torch.rand(1, 3, b - a, W0, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        ex = shard.ShardedMultiScaleFilter(net, rank, world, cnn_kernels=None if "--torch-cnn" in sys.argv else "auto")

    def run():
        with torch.no_grad():
            if whole:
                return ex(img)
            if "--batched" in sys.argv:     # lock-step stages, one batched exchange per round for the four scales
                return shard.sharded_filtering_staged(blocks, strips, rank, world, overlap="--no-overlap" not in sys.argv, streams=False if "--no-streams" in sys.argv else None)
            if staged:
                return [shard.sharded_block_forward_staged(blk, x, rank, world) for blk, x in zip(blocks, strips)]
            if world == 1 and "--streams" in sys.argv:       # one GPU: the four blocks on four streams
                return shard.sharded_filtering_staged(blocks, strips, rank, world)
            return [shard.sharded_block_forward(blk, x, rank, world) for blk, x in zip(blocks, strips)]

    for _ in range(warmup):
        run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    if "--trace" in sys.argv:
        # GPU timeline of two images on this rank (CUPTI through torch.profiler; nsys is not installed): kernel name, start, duration
        # in microseconds, in launch order - enough to see where a round's time goes (stage kernels, NCCL SendRecv, copies, gaps)
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(2):
                run()
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/infer4k_trace_rank{rank}of{world}.csv", "w") as f:
            f.write("start_us,dur_us,name\n")
            t0 = evs[0].time_range.start if evs else 0
            for e in evs:
                f.write(f"{e.time_range.start - t0:.1f},{e.time_range.end - e.time_range.start:.1f},{e.name[:80].replace(',', ';')}\n")
    if rank == 0:
        print(json.dumps({"metric": "infer_Mpix_per_s", "value": H0 * W0 / ms / 1e3, "unit": "Mpix/s", "n_gpus": world,
                          "ms_per_image": ms, "scaling": "strong", "dtype": "f32", "data": "synthetic",
                          "steps": steps, "warmup": warmup, "fw2": "--fw2" in sys.argv, "overlap": "--no-overlap" not in sys.argv,
                          "config": {"workload": ("whole v13 network, one 3840x2160 image, row strips, one row per 3x3 convolution + 8-row halo "
                                                  "exchange per solver stage, LocalNonLinearBlocks on "
                                                  + ("the PyTorch modules" if "--torch-cnn" in sys.argv else "libglrgtv kernels + cuBLAS")) if whole else
                                     "v13 four filter blocks, forward, feature maps of one 3840x2160 image, row strips, "
                                     + ("8-row halo exchange per solver stage, batched over the scales" if "--batched" in sys.argv else
                                        "8-row halo exchange per solver stage" if staged else "26-row halo exchange per block")}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
