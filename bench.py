#!/usr/bin/env python
"""bench.py - throughput of the GLR/GTV hot path (BASELINE.json metric: Mpix/s).

Workload (config[1], "model_multiscale_mixture_GLR denoising training, 256x256 patches, batch 32"):
one training pass (forward + backward, all parameter gradients) of the FOUR LocalLowpassFilteringBlock
of the shipped v13 model on the feature maps a 32 x 3 x 256 x 256 batch produces
    [32,48,256,256] G=8   [32,96,128,128] G=16   [32,192,64,64] G=16   [32,384,32,32] G=32
A "step" = that pass; pixels = B*256*256 network-input pixels per rank.  Synthetic N(0,1) feature maps,
default-init (random projection) weights.  The host CNN around the blocks is out of the hot path
 and is not run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N>1: launched under torchrun, one rank per GPU; batch-sharded (every rank its own 32-image batch, weak
scaling) with ONE NCCL all-reduce of the flattened block-parameter gradients per step inside the timed region.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS, NGRAPHS = [48, 96, 192, 384], [8, 16, 16, 32]
BATCH, RES = 32, 256
GPU_BASELINE_BATCH = 32         # the bench's own batch: reference autograd peaks at 65 GB for the largest block (one block at a time)
METRIC, UNIT = "train_Mpix_per_s", "Mpix/s"
WORKLOAD = ("v13 four LocalLowpassFilteringBlock fwd+bwd on feature maps of a 32x3x256x256 batch "
            "([32,48,256,256],[32,96,128,128],[32,192,64,64],[32,384,32,32])")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="per-rank batch (default = the config's 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-infer4k", action="store_true")
    ap.add_argument("--no-streams", action="store_true", help="run the four blocks of a step one after the other on one stream")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm
def run_cpu(steps, warmup, sample_batch=1):
    """forward+backward of the four blocks on the host cores: the UNMODIFIED reference module (oracle/_ref, vendored by
    oracle/vendor_ref.sh; kind "reference") when it travelled with the snapshot, else the oracle port (kind "port")."""
    from oracle import ref_runner as R
    if R.available():
        mpix, dt, cores = R.time_cpu(sample_batch, RES, steps, warmup)
        return mpix, dt, cores, "reference"
    import torch
    from oracle import glr_gtv_oracle as O
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    states = [{k: v.detach().clone() for k, v in M.LocalLowpassFilteringBlock(d, 1, g).state_dict().items()} for d, g in zip(DIMS, NGRAPHS)]
    xs, gs = R.make_inputs(sample_batch, RES, torch.device("cpu"))

    def port_step():
        for sd, x, g in zip(states, xs, gs):
            O.lowpass_block_fwd_bwd(sd, x, g)

    for _ in range(warmup):
        port_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        port_step()
    dt = (time.perf_counter() - t0) / steps
    return sample_batch * RES * RES / dt / 1e6, dt, cores, "port"


def cpu_sample_text(kind, cores, reps):
    what = ("the reference's own deep_multiscale_GGLR_GGTV_v1x0.LocalLowpassFilteringBlock (oracle/_ref, unmodified), eager"
            if kind == "reference" else "oracle port (oracle/glr_gtv_oracle.py)")
    return (f"batch 1 of the {BATCH} (65,536 px per step), forward+backward of the four blocks, {what}, torch CPU fp32, "
            f"{cores} threads, {reps} timed repetitions")


def reference_arm(a):
    """`--impl reference`: the reference's own implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mpix, dt, cores, kind = run_cpu(a.steps, a.warmup)
    sample = cpu_sample_text(kind, cores, a.steps)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def gpu_baseline(local, budget_s=420.0):
    """The reference module itself on the same B200 (SURVEY 8d "GPU baseline"): eager and nn.Module.compile() (as
    scripts_v2/run_abtract_lightformer_GGTV_GGLR_sigma25.py:130 runs it), TF32 as shipped (:23) and off, forward+backward of the
    four blocks.  One subprocess per variant (its own CUDA context, a hard time limit); runs after the timed regions."""
    from oracle import ref_runner as R
    if not R.available():
        return {"unavailable": "oracle/_ref not shipped"}
    out, t_start = {}, time.perf_counter()
    # measured on this pool (profiles/r02_gpu_baseline.json): eager 1.45 s / step; compile() spends ~160 s in Dynamo (a graph break at
    # every tensor-valued slice bound, ~820 per block) before its first step and then runs no faster than eager
    for name, mode, tf32, limit in (("eager", "eager", 1, 90), ("eager_tf32_off", "eager", 0, 90), ("compile", "compile", 1, 300)):
        left = budget_s - (time.perf_counter() - t_start)
        if left < 20:
            out[name] = {"skipped": "time budget of the default bench run"}
            continue
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "--device", "cuda", "--mode", mode, "--tf32", str(tf32),
               "--batch", str(GPU_BASELINE_BATCH), "--res", str(RES), "--reps", "3", "--warmup", "3", "--gpu", str(local)]
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=min(limit, left))
            lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            out[name] = json.loads(lines[-1]) if lines else {"failed": (p.stderr or "no output").strip()[-300:]}
        except subprocess.TimeoutExpired:
            out[name] = {"failed": f"no result within {int(min(limit, left))} s (graph breaks at the reference's ~820 .item() calls per block)"}
    out["note"] = (f"reference LocalLowpassFilteringBlock x4, forward+backward, batch {GPU_BASELINE_BATCH} x {RES}x{RES}, CUDA events, "
                   "3 warm-ups + 3 repetitions")
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                if t0 - 0.05 <= t <= t1 + 0.15:
                    sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if t0 - 0.05 <= t <= t1 + 0.15 and v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ roofline bookkeeping
SLOTS = ["fwd_weights", "fwd_BA", "fwd_X1", "fwd_X2", "fwd_X3", "bwd_X3", "bwd_X2", "bwd_X1", "bwd_BA", "bwd_weights",
         "proj_fwd", "proj_dgrad", "proj_wgrad", "gw"]


def algorithmic_bytes(slot, B):
    """fp32 bytes one step moves through the kernels of `slot`, summed over the four scales: every operand tensor read
    once + every result written once; halo re-reads and L2 hits are not algorithmic.
    C = channels, GE = 4G edge-weight planes of one set at full resolution; 1.25 = fine + quarter-size coarse set;
    the symmetric GTV coefficients cT are half a set (0.625 GE with its coarse part).
    Backward stages (csrc/bw2.cu, a half-resolution and a full-resolution launch each): the half-resolution launch reads z and g
    at full resolution (pooled on the fly, 2C) and writes its result vc (C/4), which the full-resolution launch reads back; the
    raw weight sets are read (GE each, x1.25) and the edge-weight gradients accumulated with red.global.add (read + write:
    2 GE per set, x1.25)."""
    total = 0
    for s, (C, G) in enumerate(zip(DIMS, NGRAPHS)):
        N = B * (RES >> s) * (RES >> s)
        GE = 4 * G
        per_px = {
            "fwd_weights": 1.25 * (2 * C + 2 * GE) + 1.875 * GE,        # feat -> wT, wL; wT -> cT
            "fwd_BA": 2 * C + 0.625 * GE,                               # y, cT -> bA
            "fwd_X1": 2 * C + 1.875 * GE,                               # bA, wL, cT -> x1
            "fwd_X2": 5 * C + 3.125 * GE,                               # x1, y, wL, cT, wT -> x2, bB, r1
            "fwd_X3": 5 * C + 1.875 * GE,                               # x2, bB, r1, x, wL, cT -> out
            # x2, gout, r1, bB, x -> gx2, gA, gB (8C) + coarse re-read 2C + vc 0.5C; wT, wL 2.5 GE; gwT, gwL += 5 GE
            "bwd_X3": 10.5 * C + 7.5 * GE,
            # A: x1, gA, r1, gx2 -> gx1 (5C + 2.5C); B: x1, gB, gx1 -> gx1 (4C + 2.5C), GTV only: wT 1.25 GE, gwT += 2.5 GE
            "bwd_X2": 14 * C + 11.25 * GE,
            "bwd_X1": 5.5 * C + 7.5 * GE,                               # bA, gx1 -> gbA
            "bwd_BA": 7.5 * C + 3.75 * GE,                              # x, gbA, gB, gout -> gx; GTV only
            "bwd_weights": 2.5 * (2 * C + 2 * GE),
            "proj_fwd": 7 * C,      # x -> feat0 (3C); space-to-depth (2C); 2x2-s2 conv (1.25C); 1x1 at half resolution (0.75C)
            "proj_dgrad": 7 * C,
            "proj_wgrad": 5 * C,    # (gfeat0, x) 3C; (gxd, s2d x) 1.25C; (gfeat1, xd) 0.75C
            "gw": 0.0,              # the round-1 gradient pass: not launched by the default backward
        }[slot]
        total += per_px * N * 4
    return total


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ 4K inference (config 4)
def infer4k(blocks, dev, rank, world, steps=3, warmup=2):
    """BASELINE config[3]: the four filter blocks, forward only, on the feature maps of ONE 3840x2160 image
    ([1,48,2160,3840] ... [1,384,270,480]); with N ranks every map is cut into N row strips and the blocks run stage by stage
    with one 8-row NCCL halo exchange per solver stage, batched over the scales (shard.sharded_filtering_staged): STRONG scaling,
    results identical to the single-GPU run (tests/test_shard_cpu.py, tests/multi_gpu_check.py)."""
    import torch
    import torch.distributed as dist
    from imagerestoration_development_unrolling_b200 import shard
    H0, W0 = 2160, 3840
    strips = []
    for s, d in enumerate(DIMS):
        a, b = shard.strip_bounds(H0 >> s, world, align=2)[rank]
        x = shard.strip_with_halo_room((1, d, b - a, W0 >> s), rank, world, device=dev)     # the rank's strip, with room for the halo rows
        x.copy_(torch.randn(1, d, b - a, W0 >> s, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + s)))
        strips.append(x)

    def run():
        with torch.no_grad():
            return shard.sharded_filtering_staged(blocks, strips, rank, world)

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
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    del strips
    torch.cuda.empty_cache()
    return {"metric": "infer_Mpix_per_s", "value": H0 * W0 / ms / 1e3, "unit": "Mpix/s", "ms_per_image": ms, "n_gpus": world,
            "scaling": "strong", "steps": steps, "warmup": warmup,
            "workload": "v13 four LocalLowpassFilteringBlock, forward, feature maps of one 3840x2160 image, row strips, one 8-row "
                        "halo exchange per solver stage batched over the scales",
            "compulsory_GBs": sum(8 * C * (H0 >> s) * (W0 >> s) for s, C in enumerate(DIMS)) / (ms / 1e3) / 1e9}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    a = parse()
    if a.impl == "reference":
        return reference_arm(a)

    import torch
    import torch.distributed as dist
    from imagerestoration_development_unrolling_b200 import _lib as L
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    from imagerestoration_development_unrolling_b200 import shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's stream at high priority: a halo exchange started while the interior rows of a stage are still being computed gets
        # SMs as CTAs retire instead of waiting behind the whole stage kernel (profiles/r02_scaling.md)
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    torch.backends.cudnn.allow_tf32 = False           # fp32 projections: the parity setting is the measured one
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = L.load()
    L.check(lib.glrgtv_check_device(), lib, "check_device")

    B = a.batch
    torch.manual_seed(0)
    blocks = [M.LocalLowpassFilteringBlock(d, 1, g).to(dev) for d, g in zip(DIMS, NGRAPHS)]
    params = [p for b in blocks for p in b.parameters()]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    shapes = [(B, d, RES >> s, RES >> s) for s, d in enumerate(DIMS)]
    xs = [torch.randn(sh, device=dev, generator=gen).requires_grad_(True) for sh in shapes]
    gs = [torch.randn(sh, device=dev, generator=gen) for sh in shapes]
    flat_grad = torch.zeros(sum(p.numel() for p in params), device=dev)

    side = None if a.no_streams else [torch.cuda.Stream(device=dev) for _ in blocks]

    def step(inputs, serial=False, ready=None):
        """ready: one event per input (its host -> device copy); a block waits only for its OWN input"""
        if side is None or serial:
            outs = []
            for i, (blk, x) in enumerate(zip(blocks, inputs)):
                if ready is not None:
                    torch.cuda.current_stream().wait_event(ready[i])
                outs.append(blk(x))
        else:                                           # V1X0:1117-1131: the four blocks are independent (autograd runs each block's
            cur = torch.cuda.current_stream()           # backward on the stream its forward ran on and joins them at the end)
            outs = []
            for i, (st, blk, x) in enumerate(zip(side, blocks, inputs)):
                st.wait_stream(cur)
                if ready is not None:
                    st.wait_event(ready[i])
                with torch.cuda.stream(st):
                    outs.append(blk(x))
            for st in side:
                cur.wait_stream(st)
        torch.autograd.backward(outs, gs, inputs=list(inputs) + params)
        if world > 1:                                   # data-parallel training: ONE gradient all-reduce over NVLink
            shard.allreduce_gradients(params, average=False, flat=flat_grad)
        for p in params:
            p.grad = None
        for x in inputs:
            x.grad = None
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step(xs)
    barrier()

    # ---- timed region (device-resident inputs), clocks sampled.  The four blocks are independent (V1X0:1117-1131) and run on four
    # streams, so kernels of different blocks overlap here; the per-kernel CUDA-event times behind `roofline` are therefore taken in
    # a second pass of the same K steps with the blocks one after the other (below), where a kernel has the GPU to itself
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = lib.glrgtv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        step(xs)
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = lib.glrgtv_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop(t0, t1) if clocks else None
    # ---- the same K steps serially, per-kernel events on: the roofline pass
    step(xs, serial=True)
    barrier()
    lib.glrgtv_profile_enable(1)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(a.steps):
        step(xs, serial=True)
    s1.record()
    barrier()
    lib.glrgtv_profile_enable(0)
    ms_serial = s0.elapsed_time(s1)
    slot_ms = (ctypes.c_float * 16)()
    slot_n = (ctypes.c_int * 16)()
    lib.glrgtv_profile_read(slot_ms, slot_n, 16)

    # ---- end-to-end: pinned host inputs -> device -> fwd+bwd -> loss back on the host, every step.
    # The H2D copy of step i+1 runs on a side stream while step i computes (double-buffered device inputs); every copy
    # and every loss read-back is inside the timed region.
    hx = [[torch.randn(sh).pin_memory() for sh in shapes] for _ in range(2)]
    dbuf = [[torch.empty(sh, device=dev) for sh in shapes] for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [[torch.cuda.Event() for _ in shapes] for _ in range(2)]     # one event per input tensor: a block starts when ITS map is in
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def start_copy(i):
        k = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])          # the step that last read this buffer has finished
            for h, d, ev in zip(hx[k], dbuf[k], copied[k]):      # largest map first: its block is more than half of the step
                d.copy_(h, non_blocking=True)
                ev.record(copy_stream)

    # the step's result (its loss) goes to pinned host memory every step; the host consumes the value of step i while step i+1 is
    # already enqueued (one event per buffer), so the read-back does not drain the GPU between steps - what an asynchronous training
    # log does.  Every value is read inside the timed region; the last one after the last step.
    loss_host = torch.empty(2, pin_memory=True)
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    seen = []

    def e2e_step(i, n):
        k = i & 1
        if i + 1 < n:
            start_copy(i + 1)
        ins = [d.detach().requires_grad_(True) for d in dbuf[k]]
        outs = step(ins, ready=copied[k])
        loss = sum(o.mean() for o in outs)
        consumed[k].record(torch.cuda.current_stream())
        loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)       # D2H read of the step's result
        loss_ready[k].record(torch.cuda.current_stream())
        if i > 0:
            loss_ready[1 - k].synchronize()
            seen.append(float(loss_host[1 - k]))

    def e2e_run(n):
        for ev in consumed:
            ev.record(torch.cuda.current_stream())
        start_copy(0)
        for i in range(n):
            e2e_step(i, n)
        loss_ready[(n - 1) & 1].synchronize()
        seen.append(float(loss_host[(n - 1) & 1]))

    e2e_run(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_run(a.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())

    # what the host -> device link delivers for these inputs on their own (explains e2e when the step is shorter than the copy)
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        h0.record(copy_stream)
        for _ in range(3):
            for h, d in zip(hx[0], dbuf[0]):
                d.copy_(h, non_blocking=True)
        h1.record(copy_stream)
    torch.cuda.synchronize()
    h2d_ms = h0.elapsed_time(h1) / 3

    i4k = None if a.no_infer4k else infer4k(blocks, dev, rank, world)

    if rank == 0:
        pix = B * RES * RES * world
        value = pix * a.steps / (ms / 1e3) / 1e6
        e2e_value = pix * a.steps / (ms_e2e / 1e3) / 1e6
        # dominant kernel of the step and its roofline
        per_slot = {SLOTS[i]: (slot_ms[i], slot_n[i]) for i in range(len(SLOTS)) if slot_n[i] > 0}
        top = max(per_slot, key=lambda k: per_slot[k][0])
        top_ms_per_step = per_slot[top][0] / a.steps
        peak, peak_src = measured_peaks()
        achieved = algorithmic_bytes(top, B) / (top_ms_per_step / 1e3) / 1e9
        whole_ms = sum(v[0] for v in per_slot.values()) / a.steps
        # DRAM bytes of the same kernels from the committed ncu capture of one step at these sizes (profiles/, tools/ncu_slots.py)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r02_step_slots.json")
        if os.path.exists(tp) and B == BATCH:
            traffic = json.load(open(tp))["slots"].get(top, {}).get("dram_bytes")
        per_kernel = {}
        for k, v in per_slot.items():
            gbs = algorithmic_bytes(k, B) / (v[0] / a.steps / 1e3) / 1e9
            per_kernel[k] = {"ms": round(v[0] / a.steps, 4), "launches": int(v[1] // a.steps), "GBs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        roofline = {
            "bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src, "traffic": traffic,
            "note": "a slot = the launches of one solver stage over the four scales (backward stages: a half- and a full-resolution "
                    "launch each, edge-weight gradients included; bwd_X2 = parts A and B); achieved = algorithmic bytes of the "
                    "slot / its summed CUDA-event time over K steps run with the four blocks one after the other "
                    "(serial_ms_per_step), right after the timed region - in the timed region itself the blocks run on four "
                    "streams and kernels of different blocks overlap",
            "per_kernel": per_kernel,
            "serial_ms_per_step": round(ms_serial / a.steps, 4),
            "kernel_share_of_step": round(top_ms_per_step / (ms_serial / a.steps), 4),
            "own_kernels_share_of_step": round(whole_ms / (ms_serial / a.steps), 4),
            "whole_block_compulsory_GBs": sum(20 * C * B * (RES >> s) ** 2 for s, C in enumerate(DIMS)) / (ms / a.steps / 1e3) / 1e9,
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_rank_batch": B, "parallelism": f"dp{world}",
                       "l2": "inputs larger than L2 (755 MB of block inputs per step)", "tf32": False, "streams": 1 if a.no_streams else 4},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": int(sum(h.numel() for h in hx[0]) * 4), "d2h_bytes_per_step": 4,
                    "h2d_copy_alone_ms": round(h2d_ms, 3), "h2d_copy_alone_GBs": round(sum(h.numel() for h in hx[0]) * 4 / h2d_ms / 1e6, 1),
                    "note": "pinned-host inputs of step i+1 are copied on a side stream while step i computes; the loss of step i is copied "
                            "to pinned host memory and consumed by the host while step i+1 is enqueued; at N > 1 the ranks' copies "
                            "(755 MB per rank and step: feature maps, not images) share the host's memory channels and PCIe root complexes"},
            "roofline": roofline,
        }
        if i4k is not None:
            line["infer4k"] = i4k
        if world == 1 and not a.no_cpu_baseline:
            mpix, dt, cores, kind = run_cpu(steps=3, warmup=1)
            line["cpu_baseline"] = {"value": mpix, "unit": UNIT, "cores": cores, "kind": kind, "sample": cpu_sample_text(kind, cores, 3)}
        if world == 1 and not a.no_gpu_baseline:
            line["gpu_baseline"] = gpu_baseline(local)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
