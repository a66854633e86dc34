"""Baseline harness over the UNMODIFIED reference module vendored in oracle/_ref/ (oracle/vendor_ref.sh).

Test / baseline infrastructure: only bench.py's baseline legs (`--impl reference`, `cpu_baseline`, `gpu_baseline`), tests/ and
tools/ call this; the product package never imports oracle/.

The workload is bench.py's: forward + backward of the four LocalLowpassFilteringBlock of the shipped v13 configuration
(deep_multiscale_GGLR_GGTV_v1x0.py:967-988, built exactly as AbtractMultiScaleGraphFilter does at :1073-1090) on feature maps
[B,48,R,R], [B,96,R/2,R/2], [B,192,R/4,R/4], [B,384,R/8,R/8].

    python oracle/ref_runner.py --device cuda --mode eager|compile --tf32 0|1 --batch 8 --res 256 --reps 3     # one JSON line
"""
import argparse
import importlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
DIMS, NGRAPHS = [48, 96, 192, 384], [8, 16, 16, 32]


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "deep_multiscale_GGLR_GGTV_v1x0.py"))


def load_reference(name="deep_multiscale_GGLR_GGTV_v1x0"):
    """import the vendored reference file as it is (no edits, no monkey patching)"""
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return importlib.import_module(name)


def build_blocks(device, seed=0):
    """the four filter blocks with the reference's default initialisation.  The reference creates its stencil constants on the
    DEFAULT device (scripts_v2/run_abtract_lightformer_GGTV_GGLR_sigma25.py:25 sets it before construction), so do the same."""
    import torch
    ref = load_reference()
    prev = torch.get_default_device()
    torch.set_default_device(device)
    try:
        torch.manual_seed(seed)
        blocks = [ref.LocalLowpassFilteringBlock(d, 1, g).to(device) for d, g in zip(DIMS, NGRAPHS)]
    finally:
        torch.set_default_device(prev)
    return blocks


def make_inputs(batch, res, device, seed=1):
    import torch
    gen = torch.Generator(device="cpu").manual_seed(seed)
    xs = [torch.randn(batch, d, res >> s, res >> s, generator=gen).to(device) for s, d in enumerate(DIMS)]
    gs = [torch.randn(batch, d, res >> s, res >> s, generator=gen).to(device) for s, d in enumerate(DIMS)]
    return xs, gs


def step(blocks, xs, gs):
    """forward + backward of the four blocks, one block at a time (the autograd tape of one block is freed before the next)"""
    import torch
    for blk, x, g in zip(blocks, xs, gs):
        xx = x.detach().requires_grad_(True)
        out = blk(xx)
        torch.autograd.backward([out], [g], inputs=[xx] + list(blk.parameters()))
        for p in blk.parameters():
            p.grad = None


def time_cpu(batch, res, steps, warmup, threads=None):
    import torch
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device("cpu")
    blocks = build_blocks(dev)
    xs, gs = make_inputs(batch, res, dev)
    for _ in range(warmup):
        step(blocks, xs, gs)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(blocks, xs, gs)
    dt = (time.perf_counter() - t0) / steps
    return batch * res * res / dt / 1e6, dt, cores


def time_gpu(batch, res, reps, warmup, mode, tf32):
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    torch.set_float32_matmul_precision("high" if tf32 else "highest")     # 'high' is what the reference's scripts set (:23)
    blocks = build_blocks(dev)
    if mode == "compile":
        for b in blocks:
            b.compile()                                                   # nn.Module.compile(), as scripts_v2/...sigma25.py:130
    xs, gs = make_inputs(batch, res, dev)
    t_first = time.perf_counter()
    step(blocks, xs, gs)
    torch.cuda.synchronize()
    first_s = time.perf_counter() - t_first
    for _ in range(max(warmup - 1, 0)):
        step(blocks, xs, gs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step(blocks, xs, gs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"mode": mode, "tf32": bool(tf32), "batch": batch, "res": res, "ms_per_step": ms, "Mpix_per_s": batch * res * res / ms / 1e3,
            "first_step_s": round(first_s, 2), "reps": reps, "warmup": warmup,
            "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--mode", default="eager", choices=["eager", "compile"])
    ap.add_argument("--tf32", type=int, default=1)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gpu", type=int, default=0)
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "oracle/_ref is empty: run oracle/vendor_ref.sh where /root/reference exists"}))
        sys.exit(0)
    if a.device == "cpu":
        mpix, dt, cores = time_cpu(a.batch, a.res, a.reps, a.warmup)
        print(json.dumps({"device": "cpu", "Mpix_per_s": mpix, "ms_per_step": dt * 1e3, "cores": cores, "batch": a.batch, "res": a.res}))
    else:
        import torch
        torch.cuda.set_device(a.gpu)
        print(json.dumps(time_gpu(a.batch, a.res, a.reps, a.warmup, a.mode, a.tf32)))
