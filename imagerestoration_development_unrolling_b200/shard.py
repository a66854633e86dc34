"""Multi-GPU sharding of the hot path.  The reference has no distributed code at all; this is
the B200-side addition, one process per GPU over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

* training: batch-sharded data parallelism.  Every op of the block is per-sample, so ranks never exchange
  activations; `allreduce_gradients` does the ONE collective of a step (flattened parameter gradients).
* full-resolution inference: row-strip partition of the image with a halo exchange in front of each filter block.
  One output pixel of a LocalLowpassFilteringBlock depends on inputs at most 25 pixels away at the block's own
  resolution (3 for bA, 7 for each of the three A(.) applications through the half-resolution branch, 7 for the
  thresholded right-hand side - SURVEY 8e measured 24-25), so a 26-row halo (even, to keep the 2x2 pooling grid
  aligned) makes the strip result exact: rows contaminated by the artificial strip border are cropped away, true
  image borders keep the reference's padding rules because no halo is added there.
"""
import contextlib
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

BLOCK_HALO_ROWS = 26
STAGE_HALO_ROWS = 8     # reach of ONE solver stage: 3 rows at full resolution, 3 coarse rows = 6 through the half-resolution branch, even


# ----------------------------------------------------------------------------------------------- training
def allreduce_gradients(params: Sequence[torch.nn.Parameter], group=None, average: bool = True,
                        flat: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Sum (or average) the gradients of `params` across ranks with ONE all-reduce of a flat buffer; afterwards every `.grad`
    is a VIEW of that buffer (no copy back).  Packing is one fused multi-tensor copy (torch._foreach_copy_), not one launch per
    parameter: the four filter blocks have 128 small parameter tensors and the step is only ~35 ms.  Parameters without a
    gradient contribute zeros.  Returns the flat buffer (pass it back in as `flat` to reuse the allocation)."""
    params = [p for p in params if p.requires_grad]
    n = sum(p.numel() for p in params)
    if flat is None or flat.numel() != n:
        flat = torch.empty(n, dtype=params[0].dtype, device=params[0].device)
    views, o = [], 0
    for p in params:
        views.append(flat[o:o + p.numel()].view_as(p))
        o += p.numel()
    have = [(v, p.grad) for v, p in zip(views, params) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
    if len(have) < len(params):
        for v, p in zip(views, params):
            if p.grad is None:
                v.zero_()
    if have:
        torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
        if average:
            flat.div_(dist.get_world_size(group))
    for v, p in zip(views, params):
        p.grad = v
    return flat


# ----------------------------------------------------------------------------------------------- inference
def strip_bounds(height: int, world: int, align: int = 2) -> List[Tuple[int, int]]:
    """[start, stop) rows of every rank: contiguous strips whose boundaries are multiples of `align`
    (16 at the network input of the 4-scale model: 8x down-sampling times the block's own 2x pooling)."""
    if height % align:
        raise ValueError(f"height {height} is not a multiple of the strip alignment {align}")
    units = height // align
    base, extra = divmod(units, world)
    bounds, start = [], 0
    for r in range(world):
        n = (base + (1 if r < extra else 0)) * align
        bounds.append((start, start + n))
        start += n
    return bounds


def exchange_row_halos(strip: torch.Tensor, halo: int, rank: int, world: int, group=None) -> Tuple[torch.Tensor, int, int]:
    """Extend a [B,C,h,W] row strip with `halo` rows of each neighbour (none at the true image border).
    Returns (extended strip, rows added on top, rows added at the bottom)."""
    if world == 1:
        return strip, 0, 0
    if strip.shape[-2] < halo:
        raise ValueError(f"strip of {strip.shape[-2]} rows is thinner than the {halo}-row halo")
    up, down = rank - 1, rank + 1
    ops, top, bot = [], None, None
    send_top = strip[..., :halo, :].contiguous()
    send_bot = strip[..., -halo:, :].contiguous()
    if up >= 0:
        top = torch.empty_like(send_top)
        ops += [dist.P2POp(dist.isend, send_top, up, group), dist.P2POp(dist.irecv, top, up, group)]
    if down < world:
        bot = torch.empty_like(send_bot)
        ops += [dist.P2POp(dist.isend, send_bot, down, group), dist.P2POp(dist.irecv, bot, down, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = ([top] if top is not None else []) + [strip] + ([bot] if bot is not None else [])
    return torch.cat(parts, dim=-2), (halo if top is not None else 0), (halo if bot is not None else 0)


def exchange_rows(first: torch.Tensor, last: torch.Tensor, rank: int, world: int, group=None):
    """Send this strip's `first` row (any shape) to the rank above and its `last` row to the rank below; returns
    (the upper neighbour's last row, the lower neighbour's first row), None at the true image border."""
    ops, top, bot = [], None, None
    if rank > 0:
        top = torch.empty_like(first)
        ops += [dist.P2POp(dist.isend, first, rank - 1, group), dist.P2POp(dist.irecv, top, rank - 1, group)]
    if rank < world - 1:
        bot = torch.empty_like(last)
        ops += [dist.P2POp(dist.isend, last, rank + 1, group), dist.P2POp(dist.irecv, bot, rank + 1, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return top, bot


def sharded_block_forward(block: Callable[[torch.Tensor], torch.Tensor], strip: torch.Tensor, rank: int, world: int,
                          halo: int = BLOCK_HALO_ROWS, group=None) -> torch.Tensor:
    """Run one filter block on this rank's row strip of a spatially sharded feature map: halo exchange with the two
    neighbours, the block on the extended strip, crop.  Exact (not approximate) for halo >= 25, see module docstring."""
    ext, t, b = exchange_row_halos(strip, halo, rank, world, group)
    out = block(ext)
    return out[..., t:out.shape[-2] - b, :].contiguous() if (t or b) else out


def exchange_row_halos_inplace(buf: torch.Tensor, top: int, bot: int, rank: int, world: int, group=None) -> None:
    """buf [B,C,top+h+bot,W] holds this rank's h rows between halo rows that belong to the neighbours: send the first `top`
    / last `bot` OWN rows up / down and receive the neighbours' rows into the halo rows (top / bot are 0 at the image border)."""
    ops, keep = [], []
    H = buf.shape[-2]
    if top:
        send = buf[..., top:2 * top, :].contiguous()
        recv = torch.empty_like(send)
        keep.append((recv, slice(0, top)))
        ops += [dist.P2POp(dist.isend, send, rank - 1, group), dist.P2POp(dist.irecv, recv, rank - 1, group)]
    if bot:
        send = buf[..., H - 2 * bot:H - bot, :].contiguous()
        recv = torch.empty_like(send)
        keep.append((recv, slice(H - bot, H)))
        ops += [dist.P2POp(dist.isend, send, rank + 1, group), dist.P2POp(dist.irecv, recv, rank + 1, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for recv, rows in keep:
            buf[..., rows, :] = recv


class CudaStageRunner:
    """The pieces of one LocalLowpassFilteringBlock forward on a local plane, through glrgtv_block_fwd_stage.
    (The CPU tests substitute a runner that drives the emulation build of the same kernels.)"""

    def __init__(self, blk):
        self.blk, self.lf = blk, blk.local_filter

    def prepare(self, ext: torch.Tensor):
        from . import ops
        lf = self.lf
        feat0, feat1 = lf._projections(ext)                   # per-pixel / 2x2-aligned: valid on the halo rows too
        params = lf._block_params() + [self.blk.skip_weight]
        saved = ops.alloc_block_saved(ext, lf.n_graphs)
        out = torch.empty_like(ext)
        calls = ops.PreparedBlockStages(ext, params, lf.n_graphs, saved, out)      # structs marshalled once per plane
        st = dict(ext=ext, calls=calls, names=dict(zip(ops._SAVED, saved)), out=out)
        calls.run(0, feat0.contiguous(), feat1.contiguous(), 0, ext.shape[-2])
        return st

    def stage(self, st, k: int, row0: int, row1: int) -> None:
        st["calls"].run(k, None, None, row0, row1)

    def buffer(self, st, name: str) -> torch.Tensor:
        return st["names"][name].view(st["ext"].shape)        # [B,G,F,H,W] -> [B,C,H,W]

    def output(self, st) -> torch.Tensor:
        return st["out"]


@torch.no_grad()
def sharded_block_forward_staged(blk, strip: torch.Tensor, rank: int, world: int, group=None, runner=None) -> torch.Tensor:
    """LocalLowpassFilteringBlock on this rank's row strip with ONE 8-row halo exchange PER SOLVER STAGE instead of one
    26-row exchange per block: x, bA, x1 and x2 are exchanged, every stage computes exactly the rank's own rows
    (glrgtv_block_fwd_stage), so nothing but the edge weights of the 8 halo rows is computed twice.  Exact: a stage reaches
    at most 7 rows beyond the rows it produces.  `blk` is the drop-in module (its streaming kernels: W % 8 == 0)."""
    if world == 1 and runner is None:
        return blk(strip)
    if strip.shape[-2] < 2 * STAGE_HALO_ROWS:
        raise ValueError(f"strip of {strip.shape[-2]} rows is too thin for two {STAGE_HALO_ROWS}-row halos")
    runner = runner or CudaStageRunner(blk)
    ext, t, b = exchange_row_halos(strip, STAGE_HALO_ROWS, rank, world, group)
    st = runner.prepare(ext.contiguous())
    H = ext.shape[-2]
    r0, r1 = t, H - b
    for k, produced in ((1, "bA"), (2, "x1"), (3, "x2")):
        runner.stage(st, k, r0, r1)
        exchange_row_halos_inplace(runner.buffer(st, produced), t, b, rank, world, group)
    runner.stage(st, 4, r0, r1)
    return runner.output(st)[..., r0:r1, :].contiguous()


def _exchange_many_start(sends_up: List[Optional[torch.Tensor]], sends_down: List[Optional[torch.Tensor]], rank: int, world: int, group=None):
    """Start one batched neighbour exchange for several tensors at once (one NCCL group = one latency instead of one per tensor):
    sends_up[i] goes to rank-1 and is answered by that rank's sends_down[i], and vice versa.  Every rank must pass lists of the
    same length and order.  Returns a handle for `_exchange_many_finish`; work enqueued in between overlaps the transfer (the
    NCCL stream waits only for what was enqueued BEFORE this call)."""
    ops, from_up, from_down = [], [None] * len(sends_up), [None] * len(sends_down)
    for i, (su, sd) in enumerate(zip(sends_up, sends_down)):
        if rank > 0:
            from_up[i] = torch.empty_like(su)
            ops += [dist.P2POp(dist.isend, su, rank - 1, group), dist.P2POp(dist.irecv, from_up[i], rank - 1, group)]
        if rank < world - 1:
            from_down[i] = torch.empty_like(sd)
            ops += [dist.P2POp(dist.isend, sd, rank + 1, group), dist.P2POp(dist.irecv, from_down[i], rank + 1, group)]
    reqs = dist.batch_isend_irecv(ops) if ops else []
    return reqs, from_up, from_down, (sends_up, sends_down)       # (the send buffers stay referenced until the exchange is over)


def _exchange_many_finish(handle):
    reqs, from_up, from_down, _ = handle
    for req in reqs:
        req.wait()
    return from_up, from_down


def _exchange_many(sends_up, sends_down, rank: int, world: int, group=None):
    return _exchange_many_finish(_exchange_many_start(sends_up, sends_down, rank, world, group))


def strip_with_halo_room(shape: Sequence[int], rank: int, world: int, halo: int = STAGE_HALO_ROWS, **kw) -> torch.Tensor:
    """An uninitialised [B,C,rows,W] row strip allocated INSIDE a buffer that has room for the neighbours' halo rows above and
    below it (none at the true image border).  A producer that writes its strip into this tensor (an `out=` argument, `copy_`)
    saves `sharded_filtering_staged` the copy of the whole strip into an extended plane - 1 ms per 4K image and rank on two GPUs
    (profiles/r02_scaling.md); any other tensor works too and is copied."""
    B, C, rows, W = shape
    t, b = (halo if rank > 0 else 0), (halo if rank < world - 1 else 0)
    ext = torch.empty(B, C, t + rows + b, W, **kw)
    ext._glrgtv_halo_room = (t, b)                                  # only buffers made here are ever written outside the strip
    return ext[:, :, t:t + rows, :]


def _extended_plane(x: torch.Tensor, t: int, b: int) -> Optional[torch.Tensor]:
    """the buffer around a strip made by strip_with_halo_room (same halo geometry), else None"""
    base = getattr(x, "_base", None)
    if base is None or getattr(base, "_glrgtv_halo_room", None) != (t, b) or base.dim() != 4 or not base.is_contiguous():
        return None
    B, C, rows, W = x.shape
    if tuple(base.shape) != (B, C, t + rows + b, W) or x.storage_offset() != base.storage_offset() + t * W or x.stride() != base.stride():
        return None
    return base


class _ScaleStreams:
    """one side stream per scale, forked from / joined to the caller's stream (cached per device)"""
    _cache = {}

    def __init__(self, ref: torch.Tensor, n: int):
        key = (ref.device, n)
        if key not in self._cache:
            self._cache[key] = [torch.cuda.Stream(device=ref.device) for _ in range(n)]
        self.side, self.dev = self._cache[key], ref.device

    def fork(self):
        cur = torch.cuda.current_stream(self.dev)
        for s in self.side:
            s.wait_stream(cur)

    def on(self, i: int):
        return torch.cuda.stream(self.side[i])

    def join(self):
        cur = torch.cuda.current_stream(self.dev)
        for s in self.side:
            cur.wait_stream(s)


OVERLAP_MIN_ROWS = 64    # strips at least this tall compute their boundary rows first and overlap the halo exchange with the interior


@torch.no_grad()
def sharded_filtering_staged(blocks: Sequence, strips: Sequence[torch.Tensor], rank: int, world: int, group=None,
                             runners: Optional[Sequence] = None, overlap: bool = True, streams: Optional[bool] = None) -> List[torch.Tensor]:
    """`sharded_block_forward_staged` for several independent filter blocks at once (the four scales of
    AbtractMultiScaleGraphFilter.filtering, V1X0:1117-1131): the blocks advance through the solver stages in lock-step and
    each round's halo rows of ALL blocks travel in ONE batched exchange - 5 exchange rounds per image instead of 5 per block.
    `overlap`: a stage first produces the 8 rows next to each neighbour (the stage kernels take any even row range), the
    exchange of those rows starts, and the interior rows are computed while it is in flight; strips shorter than OVERLAP_MIN_ROWS
    are computed in one piece (two extra launches with a 7-row pipeline fill each would cost more than the exchange hides).
    Same arithmetic, same results as the per-block form either way.  Returns row VIEWS of the extended output planes; strips made
    by `strip_with_halo_room` are used in place, others are copied into an extended plane first."""
    n = len(strips)
    if streams is None:
        streams = strips[0].is_cuda and n > 1
    if world == 1 and runners is None:                             # one rank, whole maps: the blocks as they are, one stream per scale
        if not streams:
            return [blk(x) for blk, x in zip(blocks, strips)]
        lanes = _ScaleStreams(strips[0], n)
        lanes.fork()
        outs = []
        for i, (blk, x) in enumerate(zip(blocks, strips)):
            with lanes.on(i):
                outs.append(blk(x))
        lanes.join()
        return outs
    hr = STAGE_HALO_ROWS
    for x in strips:
        if x.shape[-2] < 2 * hr:
            raise ValueError(f"strip of {x.shape[-2]} rows is too thin for two {hr}-row halos")
    runners = list(runners) if runners is not None else [CudaStageRunner(blk) for blk in blocks]
    t, b = (hr if rank > 0 else 0), (hr if rank < world - 1 else 0)
    tops, bots = _exchange_many([x[..., :hr, :].contiguous() for x in strips], [x[..., -hr:, :].contiguous() for x in strips], rank, world, group)
    # the blocks are independent (V1X0:1117-1131): each scale's kernels go to its own CUDA stream, so the launch-latency-bound
    # deep scales (a 1/8 strip of the 1/8-resolution map is 34 rows) run beside the large ones instead of after them; every
    # round joins the streams before the exchange
    lanes = _ScaleStreams(strips[0], n) if streams else None
    states = [None] * n
    for i, x in enumerate(strips):
        ext = _extended_plane(x, t, b)                             # a strip that already sits inside its extended plane: no copy
        if ext is None:
            ext = x.new_empty(x.shape[:-2] + (t + x.shape[-2] + b, x.shape[-1]))
            ext[..., t:t + x.shape[-2], :] = x
        if t:
            ext[..., :t, :] = tops[i]
        if b:
            ext[..., t + x.shape[-2]:, :] = bots[i]
        states[i] = ext
    if lanes:
        lanes.fork()
    for i in range(n):
        with (lanes.on(i) if lanes else contextlib.nullcontext()):
            states[i] = runners[i].prepare(states[i])
    ends = [x.shape[-2] + t for x in strips]                       # [t, ends[i]) = this rank's own rows inside the extended plane
    split = [overlap and world > 1 and x.shape[-2] >= OVERLAP_MIN_ROWS for x in strips]
    for k, produced in ((1, "bA"), (2, "x1"), (3, "x2")):
        bufs = [runners[i].buffer(states[i], produced) for i in range(n)]
        for i in range(n):                                         # the rows a neighbour waits for (everything, for short strips)
            with (lanes.on(i) if lanes else contextlib.nullcontext()):
                if not split[i]:
                    runners[i].stage(states[i], k, t, ends[i])
                else:
                    if t:
                        runners[i].stage(states[i], k, t, t + hr)
                    if b:
                        runners[i].stage(states[i], k, ends[i] - hr, ends[i])
        if lanes:
            lanes.join()
        handle = _exchange_many_start([bf[..., t:t + hr, :].contiguous() for bf in bufs],
                                      [bf[..., e - hr:e, :].contiguous() for bf, e in zip(bufs, ends)], rank, world, group)
        if lanes:
            lanes.fork()
        for i in range(n):                                         # interior rows, while the halo rows travel
            if split[i]:
                with (lanes.on(i) if lanes else contextlib.nullcontext()):
                    runners[i].stage(states[i], k, t + (hr if t else 0), ends[i] - (hr if b else 0))
        if lanes:
            lanes.join()
        ups, downs = _exchange_many_finish(handle)
        for bf, e, u, d in zip(bufs, ends, ups, downs):
            if u is not None:
                bf[..., :t, :] = u
            if d is not None:
                bf[..., e:, :] = d
        if lanes:
            lanes.fork()
    outs = []
    for i in range(n):
        with (lanes.on(i) if lanes else contextlib.nullcontext()):
            runners[i].stage(states[i], 4, t, ends[i])
    if lanes:
        lanes.join()
    for i in range(n):                                             # row views of the extended outputs (no copy; .contiguous() them if needed)
        outs.append(runners[i].output(states[i])[..., t:ends[i], :])
    return outs


# ----------------------------------------------------------------------------------------------- whole-model inference
def conv3x3_on_strip(conv: torch.nn.Conv2d, strip: torch.Tensor, rank: int, world: int, group=None) -> torch.Tensor:
    """A replicate-padded 3x3 convolution (the host CNN's only spatial operator besides the aligned 2x2 re-sampling,
    V1X0:929-948, 992-1005) on a row strip: one row from each neighbour inside the image, replicate padding at the true
    image border and at the left / right edges - the rows the whole-image convolution would have seen."""
    if conv.kernel_size != (3, 3) or conv.stride != (1, 1) or conv.dilation != (1, 1) or conv.padding_mode != "replicate":
        raise ValueError("conv3x3_on_strip expects the host CNN's 3x3 / stride 1 / replicate-padded convolution")
    ext, t, b = exchange_row_halos(strip, 1, rank, world, group)
    ext = torch.nn.functional.pad(ext, (1, 1, 1 - t, 1 - b), mode="replicate")
    return torch.nn.functional.conv2d(ext, conv.weight, conv.bias, 1, 0, 1, conv.groups)


class ShardedMultiScaleFilter:
    """AbtractMultiScaleGraphFilter (V1X0:1028-1173) on ONE rank's row strip of a spatially sharded image: the same
    `encode / filtering / decode / enc_dec / forward` surface, every tensor a [B,C,rows,W] strip.

    What crosses ranks: one row per 3x3 convolution (45 of them in the shipped v13 model: the embedding and one depthwise
    convolution in each of the 44 LocalNonLinearBlocks) and the filter blocks' per-stage halos
    (`sharded_block_forward_staged`).  Everything else in the host CNN is per pixel (variance norm, 1x1 convolutions, gate,
    skips, channel concat) or an aligned 2x2 stride-2 down / up-sampling, which never straddles a strip boundary because
    strips start at multiples of 16 input rows (`strip_bounds(H, world, align=16)`).

    The four filter blocks run through `sharded_filtering_staged` (lock-step solver stages, one batched halo exchange per
    round for all scales).  `stage_runner(blk)` supplies the per-block stage runner (default: CudaStageRunner on the module's
    CUDA kernels; the CPU tests plug the emulation build in here); `block_forward(blk, strip, scale)`, if given, replaces
    that with an independent call per block."""

    ALIGN = 16

    def __init__(self, model, rank: int, world: int, group=None, block_forward=None, cnn_kernels="auto", stage_runner=None):
        """cnn_kernels: "auto" = libglrgtv's LocalNonLinearBlock kernels (host_cnn.py) for CUDA strips under no_grad, the
        PyTorch modules otherwise; None = always the modules; or an object with pixel_rstd / dwconv_gate (the CPU tests)."""
        self.model, self.rank, self.world, self.group = model, rank, world, group
        self.cnn_kernels = cnn_kernels
        self.block_forward, self.stage_runner = block_forward, stage_runner

    # -- pieces of the host CNN
    def nonlinear_block(self, blk, x: torch.Tensor) -> torch.Tensor:
        """LocalNonLinearBlock (V1X0:951-964) on a strip; only its depthwise 3x3 needs neighbour rows."""
        kern = self.cnn_kernels
        if isinstance(kern, str):
            from . import host_cnn
            # libglrgtv's pixel_rstd / dwconv_gate take rows of whole 16-byte pieces (W % 4 == 0), the guard LocalNonLinearBlock.forward
            # applies too: e.g. a 336x496 CBSD68 image is 42 columns wide at the 1/8 scale and stays on the module path there
            kern = host_cnn.CudaCnnKernels() if (x.is_cuda and not torch.is_grad_enabled() and x.shape[-1] % 4 == 0) else None
        if kern is not None:
            from . import host_cnn
            ex = (lambda first, last: exchange_rows(first, last, self.rank, self.world, self.group)) if self.world > 1 else None
            return host_cnn.nonlinear_block_forward(blk, x, kern, ex)
        ll = blk.local_linear
        h = ll.channels_linear_op(blk.norm(x))
        gate, val = conv3x3_on_strip(ll.channels_local_linear_op, h, self.rank, self.world, self.group).chunk(2, dim=1)
        return blk.skip_weight[0] * x + blk.skip_weight[1] * ll.project_out(torch.sigmoid(gate) * gate * val)

    def stack(self, blocks, x: torch.Tensor) -> torch.Tensor:
        for blk in blocks:
            x = self.nonlinear_block(blk, x)
        return x

    def _check(self, img: torch.Tensor) -> None:
        if img.shape[-2] % self.ALIGN or img.shape[-1] % self.ALIGN:
            raise ValueError(f"strip of {img.shape[-2]}x{img.shape[-1]}: rows and width must be multiples of {self.ALIGN} "
                             "(pad the image first, evalpipe.pad_to_factor; cut strips with strip_bounds(H, world, align=16))")

    # -- the reference's surface
    def encode(self, img: torch.Tensor):
        self._check(img)
        m = self.model
        x = conv3x3_on_strip(m.patch_3x3_embeding.channels_local_linear_op01, img, self.rank, self.world, self.group)
        x = self.stack(m.encoder_scale_00, x)
        outs = [x]
        for i in (1, 2, 3):
            x = self.stack(getattr(m, f"encoder_scale_0{i}"), getattr(m, f"down_sample_0{i - 1}_0{i}")(x))
            outs.append(x)
        return tuple(outs)

    def filtering(self, coefs):
        blocks = [getattr(self.model, f"localfilter_scale_0{i}") for i in range(len(coefs))]
        strips = [c.contiguous() for c in coefs]
        if self.block_forward is not None:
            return tuple(self.block_forward(blk, c, i) for i, (blk, c) in enumerate(zip(blocks, strips)))
        runners = [self.stage_runner(blk) for blk in blocks] if self.stage_runner is not None else None
        # the per-stage exchange runs on the streaming stage kernels (W % 8 == 0 at every scale, i.e. W % 64 == 0 at the input) and
        # needs strips of at least two 8-row halos; other geometries - the same on every rank, so no rank is left waiting in an
        # exchange - take one 26-row halo exchange per block, which works with any block implementation
        staged = runners is not None or all(c.shape[-1] % 8 == 0 for c in strips)
        if self.world > 1:
            # strip heights differ between ranks: agree on the path (and on failing) BEFORE any exchange, so that no rank raises
            # alone and leaves its neighbours waiting
            thin = torch.tensor([int(any(c.shape[-2] < 2 * STAGE_HALO_ROWS for c in strips)),
                                 int(any(c.shape[-2] < BLOCK_HALO_ROWS for c in strips))], device=strips[0].device)
            dist.all_reduce(thin, op=dist.ReduceOp.MAX, group=self.group)
            staged = staged and int(thin[0].item()) == 0
            if not staged and int(thin[1].item()):
                raise ValueError(f"some rank's strip is thinner than the {BLOCK_HALO_ROWS}-row halo of a filter block: use fewer ranks")
        with torch.no_grad():
            if staged:
                return tuple(sharded_filtering_staged(blocks, strips, self.rank, self.world, self.group, runners))
            return tuple(sharded_block_forward(blk, c, self.rank, self.world, group=self.group) for blk, c in zip(blocks, strips))

    def decode(self, coefs):
        m = self.model
        x = coefs[3]
        for i in (2, 1, 0):
            up = getattr(m, f"up_sample_0{i + 1}_0{i}")(x)
            x = getattr(m, f"combine_channels_0{i}")(torch.cat([up, coefs[i]], 1))
            x = self.stack(getattr(m, f"decoder_scale_0{i}"), x)
        return m.linear_output(self.stack(m.refining_block, x))

    def enc_dec(self, img: torch.Tensor) -> torch.Tensor:
        return self.decode(self.encode(img))

    @torch.no_grad()
    def forward(self, img: torch.Tensor) -> torch.Tensor:
        return self.decode(self.filtering(self.encode(img)))

    __call__ = forward


@torch.no_grad()
def sharded_restore(model, image: torch.Tensor, rank: int, world: int, group=None, block_forward=None) -> torch.Tensor:
    """Config 4 end to end: every rank holds the whole padded [B,3,H,W] input (H, W multiples of 16), restores its own row
    strip with ShardedMultiScaleFilter and returns that strip ([B,3,rows,W]); `strip_bounds(H, world, 16)[rank]` says which rows."""
    a, b = strip_bounds(image.shape[-2], world, ShardedMultiScaleFilter.ALIGN)[rank]
    return ShardedMultiScaleFilter(model, rank, world, group, block_forward)(image[..., a:b, :].contiguous())
