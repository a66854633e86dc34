"""Multi-GPU check, launched by hand on a box with >= 2 GPUs:

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

(1) spatially sharded inference of a filter block (row strips + NCCL halo exchange) equals the single-GPU result;
(2) batch-sharded training: all-reduced parameter gradients equal the gradients of the concatenated batch;
(3) the whole v13 network on row strips (shard.ShardedMultiScaleFilter) equals the network on the whole image."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M  # noqa: E402
from imagerestoration_development_unrolling_b200 import shard  # noqa: E402
from oracle.glr_gtv_oracle import randomize_block_state  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    blk = M.LocalLowpassFilteringBlock(48, 1, 8)
    blk.load_state_dict(randomize_block_state({k: v.clone() for k, v in blk.state_dict().items()}, seed=1))
    blk = blk.to(dev)
    gen = torch.Generator().manual_seed(5)

    # (1) sharded inference on a 1 x 48 x 272 x 480 feature map (a 4K image at 1/8 resolution has 270 x 480)
    H, W = 16 * 2 * world * 4 + 16, 480
    x = torch.randn(1, 48, H, W, generator=gen).to(dev)
    a, b = shard.strip_bounds(H, world, align=2)[rank]
    with torch.no_grad():
        full = blk(x)
        mine = shard.sharded_block_forward(blk, x[:, :, a:b].contiguous(), rank, world)
    err = float((mine - full[:, :, a:b]).abs().max() / full.abs().max())
    t = torch.tensor([err], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"sharded inference, {world} strips of a {H}x{W} map: max rel err vs single GPU = {float(t):.2e}")
    assert float(t) < 1e-6
    # (1b) the same with one 8-row exchange per solver stage (no redundant rows)
    with torch.no_grad():
        mine = shard.sharded_block_forward_staged(blk, x[:, :, a:b].contiguous(), rank, world)
    err = float((mine - full[:, :, a:b]).abs().max() / full.abs().max())
    t = torch.tensor([err], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"staged sharded inference (8-row exchange per stage), {world} strips: max rel err vs single GPU = {float(t):.2e}")
    assert float(t) < 1e-6
    # (1c) several blocks in lock-step with one batched exchange per round (here: the same block on two different maps)
    x2 = torch.randn(1, 48, H, W, generator=gen).to(dev)
    with torch.no_grad():
        full2 = blk(x2)
        mine = shard.sharded_filtering_staged([blk, blk], [x[:, :, a:b].contiguous(), x2[:, :, a:b].contiguous()], rank, world)
    err = max(float((mine[0] - full[:, :, a:b]).abs().max() / full.abs().max()), float((mine[1] - full2[:, :, a:b]).abs().max() / full2.abs().max()))
    t = torch.tensor([err], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"batched staged sharded inference (two maps, one exchange per round), {world} strips: max rel err vs single GPU = {float(t):.2e}")
    assert float(t) < 1e-6

    # (2) batch-sharded gradients
    xb = torch.randn(2 * world, 48, 64, 64, generator=gen).to(dev)
    gb = torch.randn(2 * world, 48, 64, 64, generator=gen).to(dev)
    params = list(blk.parameters())
    out = blk(xb)
    ref = torch.autograd.grad(out, params, gb)
    out = blk(xb[2 * rank:2 * rank + 2].contiguous())
    for p, g in zip(params, torch.autograd.grad(out, params, gb[2 * rank:2 * rank + 2].contiguous())):
        p.grad = g
    shard.allreduce_gradients(params, average=False)
    worst = max(float((p.grad - r).norm() / r.norm().clamp_min(1e-20)) for p, r in zip(params, ref) if float(r.abs().max()) > 0)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"batch-sharded training, {world} ranks: worst relative gradient error vs the full batch = {float(t):.2e}")
    assert float(t) < 1e-4

    # (3) whole network, spatially sharded: 128 input rows per rank (16 rows at 1/8 resolution: the thinnest strip the
    #     per-stage exchange takes), width 256
    torch.manual_seed(2)
    net = M.AbtractMultiScaleGraphFilter(dims=[48, 96, 192, 384], hidden_dims=[96, 192, 384, 768], nsubnets=[1, 1, 1, 1],
                                         ngraphs=[8, 16, 16, 32], num_blocks=[4, 6, 6, 8], num_blocks_out=4).to(dev).eval()
    for i in range(4):
        fb = getattr(net, f"localfilter_scale_0{i}")
        fb.load_state_dict(randomize_block_state({k: v.cpu().clone() for k, v in fb.state_dict().items()}, seed=30 + i))
    img = torch.rand(1, 3, 128 * world, 256, generator=gen).to(dev)
    a, b = shard.strip_bounds(img.shape[-2], world, align=16)[rank]
    with torch.no_grad():
        full = net(img)
        mine = shard.sharded_restore(net, img, rank, world)
    err = float((mine - full[:, :, a:b]).abs().max() / full.abs().max())
    t = torch.tensor([err], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"whole v13 network on {world} row strips of a {img.shape[-2]}x256 image: max rel err vs single GPU = {float(t):.2e}")
    assert float(t) < 1e-5
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
