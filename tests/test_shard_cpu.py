"""Host-side multi-GPU logic on CPU with the gloo backend, world_size 2 and 3:
gradient all-reduce for batch-sharded training and row-strip halo exchange for spatially sharded inference.
The 'block' run on each strip is the CPU oracle, so this also pins the 26-row halo as sufficient for exactness."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import glr_gtv_oracle as O
from imagerestoration_development_unrolling_b200 import shard
from tests.util import random_block_state


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)


def _grad_worker(rank, world, port, q):
    _init(rank, world, port)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5.0) * (rank + 1)
    # params[2] has no gradient on any rank: must come back as zeros
    shard.allreduce_gradients(params, average=False)
    q.put((rank, [p.grad.numpy().copy() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_allreduce_gradients(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in procs]
    tot = sum(range(1, world + 1))
    for r in range(world):
        g = [torch.from_numpy(t) for t in res[r]]
        assert torch.equal(g[0], torch.full((3, 4), float(tot)))
        assert torch.equal(g[1], torch.arange(5.0) * tot)
        assert torch.equal(g[2], torch.zeros(2, 2))


def test_strip_bounds():
    assert shard.strip_bounds(2160, 8, align=16) == [(0, 272), (272, 544), (544, 816), (816, 1088), (1088, 1360),
                                                     (1360, 1632), (1632, 1904), (1904, 2160)]
    assert shard.strip_bounds(64, 3, align=2) == [(0, 22), (22, 44), (44, 64)]
    with pytest.raises(ValueError):
        shard.strip_bounds(30, 2, align=4)


def _strip_worker(rank, world, port, q, sd, x, bounds):
    _init(rank, world, port)
    a, b = bounds[rank]
    strip = x[:, :, a:b].contiguous()
    block = lambda t: O.lowpass_block_forward(sd, t)
    out = shard.sharded_block_forward(block, strip, rank, world)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_block_equals_full_image(world):
    dim, G, H, W = 12, 2, 96, 24
    sd = {k: v.double() for k, v in random_block_state(dim, G, seed=9).items()}
    x = torch.randn(1, dim, H, W, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    full = O.lowpass_block_forward(sd, x)
    bounds = shard.strip_bounds(H, world, align=2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_strip_worker, args=(r, world, port, q, sd, x, bounds)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(world))
    [p.join(60) for p in procs]
    got = torch.cat([torch.from_numpy(res[r]) for r in range(world)], dim=2)
    assert got.shape == full.shape
    assert float((got - full).abs().max()) < 1e-12      # exact: the 26-row halo covers the block's receptive radius


def test_halo_must_cover_the_receptive_radius():
    """a thin halo changes interior strip rows; the 26-row halo (structural radius 25) does not"""
    dim, G, H, W = 6, 1, 80, 8
    sd = {k: v.double() for k, v in random_block_state(dim, G, seed=2).items()}
    x = torch.randn(1, dim, H, W, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    full = O.lowpass_block_forward(sd, x)
    errs = {}
    for halo in (8, 26):
        out = O.lowpass_block_forward(sd, x[:, :, :40 + halo])[:, :, :40]
        errs[halo] = float((out - full[:, :, :40]).abs().max())
    assert errs[8] > 1e-9 and errs[26] < 1e-14, errs


def _inplace_worker(rank, world, port, q):
    _init(rank, world, port)
    h, halo = 6, 2
    top, bot = (halo if rank > 0 else 0), (halo if rank < world - 1 else 0)
    buf = torch.full((1, 2, top + h + bot, 3), float("nan"))
    # own rows hold their global row index (rank * h + local row)
    buf[..., top:top + h, :] = (rank * h + torch.arange(h, dtype=torch.float32)).view(1, 1, h, 1)
    shard.exchange_row_halos_inplace(buf, top, bot, rank, world)
    q.put((rank, buf[0, 0, :, 0].numpy().copy(), top))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_inplace_halo_exchange(world):
    """after the exchange every buffer row (own or halo) holds its global row index"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_inplace_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(60) for p in procs]
    for rank, rows, top in res:
        first = rank * 6 - top
        assert rows.tolist() == [float(first + i) for i in range(len(rows))], (rank, rows)


class EmuStageRunner:
    """shard.CudaStageRunner's interface on the g++ EMULATION build of the same kernels (CPU tensors)"""

    def __init__(self, sd, G):
        self.sd, self.G = sd, G

    def prepare(self, ext):
        from imagerestoration_development_unrolling_b200 import _lib as L
        from tests import emu_harness as E
        from tests.util import block_structs, alloc_saved, oracle_features
        B, C, H, W = ext.shape
        F = C // self.G
        f0, f1 = oracle_features(self.sd, ext)
        p, keep = block_structs(self.sd)
        sv, saved = alloc_saved(B, self.G, F, H, W)
        for v in saved.values():
            v.fill_(float("nan"))
        st = dict(ext=ext, p=p, keep=keep, sv=sv, saved=saved, out=torch.full_like(ext, float("nan")), shp=L.make_shape(B, self.G, F, H, W), E=E)
        E.call("glrgtv_block_fwd_stage", 0, st["shp"], p, ext, f0, f1, st["out"], sv, 0, H, None)
        return st

    def stage(self, st, k, row0, row1):
        st["E"].call("glrgtv_block_fwd_stage", k, st["shp"], st["p"], st["ext"], None, None, st["out"], st["sv"], row0, row1, None)

    def buffer(self, st, name):
        return st["saved"][name].view(st["ext"].shape)

    def output(self, st):
        return st["out"]


def _staged_worker(rank, world, port, q, sd, x, bounds, G):
    _init(rank, world, port)
    a, b = bounds[rank]
    out = shard.sharded_block_forward_staged(None, x[:, :, a:b].contiguous(), rank, world, runner=EmuStageRunner(sd, G))
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_staged_sharded_block_equals_full_image(world):
    """the per-stage halo exchange (one 8-row exchange per solver stage, gloo) with the emulated CUDA kernels on every rank
    reproduces the whole-image oracle; NaN-filled buffers prove no stage read a row that nobody produced or exchanged"""
    dim, G, H, W = 12, 2, 72, 16
    sd = random_block_state(dim, G, seed=13)
    x = torch.randn(1, dim, H, W, generator=torch.Generator().manual_seed(6))
    full = O.lowpass_block_forward({k: v.double() for k, v in sd.items()}, x.double())
    bounds = shard.strip_bounds(H, world, align=2)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_staged_worker, args=(r, world, port, q, sd, x, bounds, G)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(world))
    [p.join(60) for p in procs]
    got = torch.cat([torch.from_numpy(res[r]) for r in range(world)], dim=2)
    assert got.shape == full.shape and torch.isfinite(got).all()
    assert float((got.double() - full).norm() / full.norm()) < 1e-5


def _lockstep_worker(rank, world, port, q, sds, xs, G, in_place):
    _init(rank, world, port)
    shard.OVERLAP_MIN_ROWS = 24            # strips of 32+ rows: boundary rows first, exchange, interior rows (the overlapped schedule)
    strips = []
    for x in xs:
        a, b = shard.strip_bounds(x.shape[-2], world, align=2)[rank]
        if in_place:                       # the strip already sits inside a buffer with room for the halo rows: used without a copy
            s = shard.strip_with_halo_room((x.shape[0], x.shape[1], b - a, x.shape[-1]), rank, world)
            s._base.fill_(float("nan"))
            s.copy_(x[:, :, a:b])
        else:
            s = x[:, :, a:b].contiguous()
        strips.append(s)
    outs = shard.sharded_filtering_staged([None] * len(xs), strips, rank, world, runners=[EmuStageRunner(sd, G) for sd in sds])
    q.put((rank, [o.contiguous().numpy() for o in outs]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,in_place", [(2, True), (3, True), (3, False)])
def test_lockstep_filtering_with_overlapped_exchange(world, in_place):
    """shard.sharded_filtering_staged on two maps of different height: boundary rows first, halo exchange in flight during the
    interior rows (tall strips) or after the whole stage (short strips), strips used in place inside their extended planes -
    against the whole-image oracle; NaN-filled buffers prove no stage read a row nobody produced or exchanged"""
    dim, G = 12, 2
    sds = [random_block_state(dim, G, seed=17), random_block_state(dim, G, seed=18)]
    xs = [torch.randn(1, dim, 96, 16, generator=torch.Generator().manual_seed(3)), torch.randn(1, dim, 48, 16, generator=torch.Generator().manual_seed(4))]
    full = [O.lowpass_block_forward({k: v.double() for k, v in sd.items()}, x.double()) for sd, x in zip(sds, xs)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lockstep_worker, args=(r, world, port, q, sds, xs, G, in_place)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(world))
    [p.join(60) for p in procs]
    for i, ref in enumerate(full):
        got = torch.cat([torch.from_numpy(res[r][i]) for r in range(world)], dim=2)
        assert got.shape == ref.shape and torch.isfinite(got).all()
        assert float((got.double() - ref).norm() / ref.norm()) < 1e-5


# ----------------------------------------------------------------------------------------------- whole model on strips
def _tiny_model(seed=3):
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    torch.manual_seed(seed)
    m = M.AbtractMultiScaleGraphFilter(dims=[8, 8, 12, 8], hidden_dims=[8, 12, 8, 8], ngraphs=[2, 2, 2, 2],
                                       num_blocks=[2, 1, 1, 1], num_blocks_out=1).eval()
    with torch.no_grad():          # move the filter blocks and the skips off their init, as every parity test does
        for i in range(4):
            blk = getattr(m, f"localfilter_scale_0{i}")
            blk.load_state_dict(O.randomize_block_state({k: v.clone() for k, v in blk.state_dict().items()}, 20 + i))
        for n, p in m.named_parameters():
            if n.endswith("skip_weight") and "localfilter" not in n:
                p.copy_(torch.tensor([0.8, 0.6]))
    return m


def _oracle_filtering(m, coefs):
    outs = []
    for i, c in enumerate(coefs):
        sd = {k: v.double() for k, v in getattr(m, f"localfilter_scale_0{i}").state_dict().items()}
        outs.append(O.lowpass_block_forward(sd, c.double()).float())
    return tuple(outs)


def _model_worker(rank, world, port, q, img, with_blocks, cnn="torch", batched=True):
    _init(rank, world, port)
    m = _tiny_model()
    kernels = None
    if cnn == "emu":
        from tests.test_emu_host_cnn import EmuCnnKernels
        kernels = EmuCnnKernels()

    def emu_block(blk, strip, scale):
        sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
        return shard.sharded_block_forward_staged(None, strip, rank, world, runner=EmuStageRunner(sd, blk.local_filter.n_graphs))

    def emu_runner(blk):
        return EmuStageRunner({k: v.detach().clone() for k, v in blk.state_dict().items()}, blk.local_filter.n_graphs)

    if batched:     # the default form: lock-step stages, one batched exchange per round for the four scales
        ex = shard.ShardedMultiScaleFilter(m, rank, world, stage_runner=emu_runner, cnn_kernels=kernels)
    else:           # an independent staged exchange per block
        ex = shard.ShardedMultiScaleFilter(m, rank, world, block_forward=emu_block, cnn_kernels=kernels)
    a, b = shard.strip_bounds(img.shape[-2], world, ex.ALIGN)[rank]
    strip = img[:, :, a:b].contiguous()
    with torch.no_grad():
        out = ex(strip) if with_blocks else ex.enc_dec(strip)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _run_model(world, img, with_blocks, cnn="torch", batched=True):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_model_worker, args=(r, world, port, q, img, with_blocks, cnn, batched)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=600) for _ in range(world))
    [p.join(60) for p in procs]
    return torch.cat([torch.from_numpy(res[r]) for r in range(world)], dim=2)


@pytest.mark.parametrize("world,cnn", [(2, "torch"), (3, "torch"), (3, "emu")])
def test_sharded_host_cnn_equals_full_image(world, cnn):
    """encoder + decoder of the 4-scale model on row strips (one exchanged row per 3x3 convolution, aligned 2x2 re-sampling)
    against the module on the whole image; strips of unequal height (world 3: 32 + 32 + 16 rows).  cnn = "emu": the
    LocalNonLinearBlocks through host_cnn.py on the emulated glrgtv_pixel_rstd / glrgtv_dwconv_gate (scaled rows exchanged)"""
    img = torch.rand(2, 3, 80 if world == 3 else 64, 32, generator=torch.Generator().manual_seed(8))
    m = _tiny_model()
    with torch.no_grad():
        full = m.enc_dec(img)
    got = _run_model(world, img, with_blocks=False, cnn=cnn)
    assert got.shape == full.shape
    assert float((got - full).abs().max()) < (1e-5 if cnn == "torch" else 5e-5) * float(full.abs().max())


@pytest.mark.parametrize("world,batched", [(2, True), (2, False), (3, True)])
def test_sharded_whole_model_equals_full_image(world, batched):
    """config 4 in miniature: the whole AbtractMultiScaleGraphFilter on row strips - host CNN with row exchanges, the four
    filter blocks through the per-stage halo exchange on the emulated CUDA kernels (batched over the scales, or one block at a
    time) - against the module's CNN + the oracle blocks; world 3 has a first / middle / last rank"""
    img = torch.rand(1, 3, 128 * world, 64, generator=torch.Generator().manual_seed(9))
    m = _tiny_model()
    with torch.no_grad():
        full = m.decode(_oracle_filtering(m, m.encode(img)))
    got = _run_model(world, img, with_blocks=True, batched=batched)
    assert got.shape == full.shape and torch.isfinite(got).all()
    assert float((got - full).norm() / full.norm()) < 1e-5


def test_sharded_model_rejects_unaligned_strips():
    m = _tiny_model()
    with pytest.raises(ValueError, match="multiples of 16"):
        shard.ShardedMultiScaleFilter(m, 0, 1).encode(torch.rand(1, 3, 40, 32))


# ----------------------------------------------------------------------------------------------- partition properties
def test_strip_bounds_properties():
    """strips tile [0, H) in order, every boundary is a multiple of the alignment, heights differ by at most one unit"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(units=st.integers(1, 400), world=st.integers(1, 16), align=st.sampled_from([2, 16]))
    def check(units, world, align):
        H = units * align
        b = shard.strip_bounds(H, world, align)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == H
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all(a % align == 0 and e % align == 0 and e >= a for a, e in b)
        sizes = [(e - a) // align for a, e in b]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)

    check()
    with pytest.raises(ValueError):
        shard.strip_bounds(30, 2, 16)
