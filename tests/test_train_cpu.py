"""Training-loop callers (train.py) on CPU: model factory, the reference's checkpoint layout and resume rule,
bit-exact resume, and the gloo world-2 loop.  The filter blocks have no CPU path, so the loop runs a small stand-in model with
the same encode / decode / forward surface, registered through MODEL_TYPES."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from imagerestoration_development_unrolling_b200 import train as T

pytestmark = pytest.mark.usefixtures("isolated_rng")


class TinyNet(nn.Module):
    def __init__(self, width=4):
        super().__init__()
        self.enc = nn.Conv2d(3, width, 3, padding=1, padding_mode="replicate", bias=False)
        self.dec = nn.Conv2d(width, 3, 1, bias=False)

    def encode(self, img):
        return (torch.tanh(self.enc(img)),)

    def decode(self, coefs):
        return self.dec(coefs[0])

    def forward(self, img):
        return self.decode(self.encode(img))


def _conf(root, total, every=3, bs=2):
    return {"name": "unit", "manual_seed": 2204, "path": {"root_dir": str(root)},
            "datasets": {"train": {"type": "SyntheticNoisyPatches", "dataset_args": {"patch_size": 8, "lambda_noise": 25.0, "max_num_patchs": 64},
                                   "dataloader_args": {"batch_size": bs}}},
            "model": {"type": "tiny", "args": {"width": 4}},
            "train": {"total_iters": total, "checkpoint_every": every, "log_every": 0, "optimizer": {"lr": 1e-2}}}


@pytest.fixture(autouse=True)
def _register():
    T.MODEL_TYPES["tiny"] = TinyNet
    yield
    T.MODEL_TYPES.pop("tiny", None)


def test_model_factory_defaults_and_errors():
    m = T.build_model({"type": "AbtractMultiScaleGraphFilter", "args": {"dims": [8, 8, 8, 8], "hidden_dims": [8, 8, 8, 8], "ngraphs": [2, 2, 2, 2],
                                                                        "num_blocks": [1, 1, 1, 1], "num_blocks_out": 1}})
    assert len(m.encoder_scale_00) == 1 and m.linear_output.out_channels == 3
    assert T.V13_ARGS["ngraphs"] == [8, 16, 16, 32] and T.V13_ARGS["num_blocks"] == [4, 6, 6, 8]
    with pytest.raises(KeyError, match="known types"):
        T.build_model({"type": "nope"})


def test_checkpoint_layout_and_latest(tmp_path):
    m = TinyNet()
    opt, sch = T.build_optimizer(m)
    assert T.checkpoint_name(0, 330000) == "checkpoints_epoch00_iter0330k.pt"          # the file the reference script mentions
    T.save_checkpoint(str(tmp_path), 0, 5000, m, opt, sch)
    p = T.save_checkpoint(str(tmp_path), 0, 10000, m, opt, sch)
    assert T.latest_checkpoint(str(tmp_path)) == p and T.latest_checkpoint(str(tmp_path / "missing")) is None
    state = torch.load(p, weights_only=False)
    assert set(state) == {"i", "model", "optimizer", "lr_scheduler"} and state["i"] == 10000
    m2 = TinyNet()
    opt2, sch2 = T.build_optimizer(m2)
    assert T.load_checkpoint(p, m2, opt2, sch2) == 10000
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_lr_schedule_follows_the_reference():
    opt, sch = T.build_optimizer(TinyNet(), {"step_every": 2, "n_steps": 3, "cosine_iters": 10})
    lrs = []
    for _ in range(8):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    g = 0.5 ** 0.25
    assert lrs[:6] == pytest.approx([4e-4, 4e-4, 4e-4 * g, 4e-4 * g, 4e-4 * g * g, 4e-4 * g * g])
    assert lrs[6] == pytest.approx(5e-5) and lrs[7] < lrs[6]                             # the cosine phase restarts from its own base lr


def test_sampler_shards_and_resumes():
    a = list(T.ResumableShardedSampler(20, 2, rank=0, world=2))
    b = list(T.ResumableShardedSampler(20, 2, rank=1, world=2))
    assert a[:2] == [[0, 1], [4, 5]] and b[:2] == [[2, 3], [6, 7]] and len(a) == 5
    assert list(T.ResumableShardedSampler(20, 2, rank=1, world=2, start_batch=3)) == b[3:]
    d = T.SyntheticNoisyPatches(patch_size=8, max_num_patchs=4)
    n0, c0 = d[1]
    n1, c1 = d[1]
    assert torch.equal(n0, n1) and n0.shape == (8, 8, 3) and float((n0 - c0).std()) == pytest.approx(25 / 255, rel=0.2)


def test_resume_is_bit_exact(tmp_path):
    losses = {}
    straight = T.train(_conf(tmp_path / "a", total=6), torch.device("cpu"), on_step=lambda i, l: losses.setdefault(i, l))
    assert sorted(losses) == list(range(6)) and losses[5] < losses[0]
    T.train(_conf(tmp_path / "b", total=3), torch.device("cpu"))
    seen = []
    resumed = T.train(_conf(tmp_path / "b", total=6), torch.device("cpu"), on_step=lambda i, l: seen.append(i))
    assert seen == [3, 4, 5]
    for a, b in zip(straight.state_dict().values(), resumed.state_dict().values()):
        assert torch.equal(a, b)


def _worker(rank, world, port, root, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    T.MODEL_TYPES["tiny"] = TinyNet
    m = T.train(_conf(root, total=4, every=2), torch.device("cpu"))
    q.put((rank, [v.numpy().copy() for v in m.state_dict().values()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_stay_in_sync_and_rank0_checkpoints(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    for a, b in zip(res[0], res[1]):
        assert (a == b).all()
    folder = T.checkpoints_folder(_conf(tmp_path, 4))
    assert os.listdir(folder) == ["checkpoints_epoch00_iter0000k.pt"]
    assert torch.load(os.path.join(folder, os.listdir(folder)[0]), weights_only=False)["i"] == 3


def test_sampler_partition_properties():
    """over all ranks every global batch is covered exactly once, in order, and a resume point skips exactly the batches before it"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(n=st.integers(0, 300), bs=st.integers(1, 7), world=st.integers(1, 5), start=st.integers(0, 20))
    def check(n, bs, world, start):
        per_rank = [list(T.ResumableShardedSampler(n, bs, r, world, start)) for r in range(world)]
        n_batches = max(n // (bs * world) - start, 0)
        assert all(len(p) == n_batches == len(T.ResumableShardedSampler(n, bs, r, world, start)) for r, p in enumerate(per_rank))
        for k in range(n_batches):
            seen = [i for r in range(world) for i in per_rank[r][k]]
            first = (start + k) * bs * world
            assert seen == list(range(first, first + bs * world))

    check()


def test_validation_psnr_during_training(tmp_path):
    """datasets.val + train.validate_every: PSNR on quantised 0..255 images as the reference's validation loop computes it"""
    conf = _conf(tmp_path, total=4)
    conf["datasets"]["val"] = {"type": "SyntheticNoisyPatches", "dataset_args": {"patch_size": [10, 12], "lambda_noise": 25.0, "max_num_patchs": 3, "seed": 7}}
    conf["train"]["validate_every"] = 2
    seen = []
    model = T.train(conf, torch.device("cpu"), on_step=lambda i, v: seen.append((i, v)))
    psnrs = [(i, v["psnr"]) for i, v in seen if isinstance(v, dict)]
    assert [i for i, _ in psnrs] == [1, 3] and all(5.0 < p < 60.0 for _, p in psnrs)
    assert model.training                                           # validation restores train mode
    # the same number by hand for the last model state: 10x12 is padded to 16x16 (reflect), cropped back, clamped, quantised
    ds = T.SyntheticNoisyPatches(patch_size=[10, 12], lambda_noise=25.0, max_num_patchs=3, seed=7)
    vals = []
    model.eval()
    with torch.no_grad():
        for k in range(3):
            noisy, clean = ds[k]
            x = torch.nn.functional.pad(noisy.permute(2, 0, 1)[None], (0, 4, 0, 6), mode="reflect")
            out = torch.round(model(x)[:, :, :10, :12].clamp(0, 1) * 255)
            ref = torch.round(clean.permute(2, 0, 1)[None].clamp(0, 1) * 255)
            vals.append(20 * torch.log10(255.0 / torch.sqrt(torch.mean((ref - out) ** 2))))
    assert abs(float(torch.stack(vals).mean()) - psnrs[-1][1]) < 1e-4


def test_loop_runs_epochs_until_total_iters(tmp_path):
    """a dataset of 3 global batches trained for 8 iterations: the loop wraps around (epochs) instead of stopping silently, writes
    the final checkpoint, and a resume in the middle of the second epoch reproduces the uninterrupted run bit for bit"""
    conf = _conf(tmp_path / "a", total=8, every=5)
    conf["datasets"]["train"]["dataset_args"]["max_num_patchs"] = 6
    seen = []
    m_full = T.train(conf, device=torch.device("cpu"), on_step=lambda i, v: seen.append(i))
    assert seen == list(range(8))
    last = torch.load(T.latest_checkpoint(T.checkpoints_folder(conf)), weights_only=False)
    assert 7 in [v for v in last.values() if isinstance(v, int)]                  # the final checkpoint (iteration 7) was written
    conf_b = _conf(tmp_path / "b", total=5, every=5)
    conf_b["datasets"]["train"]["dataset_args"]["max_num_patchs"] = 6
    T.train(conf_b, device=torch.device("cpu"))
    conf_b["train"]["total_iters"] = 8
    m_res = T.train(conf_b, device=torch.device("cpu"))
    for a, b in zip(m_full.state_dict().values(), m_res.state_dict().values()):
        assert torch.equal(a, b)
    conf_c = _conf(tmp_path / "c", total=2, bs=4)
    conf_c["datasets"]["train"]["dataset_args"]["max_num_patchs"] = 3
    with pytest.raises(ValueError, match="smaller than one global batch"):
        T.train(conf_c, device=torch.device("cpu"))
