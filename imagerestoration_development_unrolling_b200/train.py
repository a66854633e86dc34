"""Training-loop callers of the hot path: what the reference's `run_train.py` (a stub that stops after
building the dataloader, run_train.py:38-99) and its per-experiment scripts (scripts_v2/run_abtract_lightformer_GGTV_GGLR_sigma25.py)
do around the model, as one YAML-driven, resumable, multi-GPU entry point.

    python -m imagerestoration_development_unrolling_b200.train --conf experiment.yaml
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m imagerestoration_development_unrolling_b200.train --conf experiment.yaml

The YAML is the reference's `experiment_conf/example.yaml` (name, manual_seed, path.root_dir, datasets.train) plus the two
sections it never got: `model: {type, args}` and `train: {...}`.  Kept from the reference: the checkpoint dict
`{'i', 'model', 'optimizer', 'lr_scheduler'}` and its file name `checkpoints_epochEE_iterKKKKk.pt` under
`<root_dir>/experiments/<name>/learning_checkpoints/` (scripts_v2:225-232), automatic resume from the last file of the sorted
folder listing (run_train.py:42-58), Adam(lr 4e-4, eps 1e-8) under SequentialLR[MultiStepLR, CosineAnnealingLR]
(scripts_v2:147-165), the three-term loss (scripts_v2:191-208) and the NHWC batches of the dataloaders.

Multi-GPU: one process per GPU, the batch is sharded by a rank-strided resumable sampler, gradients are averaged with the one
all-reduce of `shard.allreduce_gradients`; rank 0 writes checkpoints.  Datasets are out of the hot path: the built-in
`SyntheticNoisyPatches` stands in for the reference's CSV-driven image datasets (same tensors: noisy / clean [H,W,3] in [0,1],
noise sigma/255 un-clipped); `DATASET_TYPES` takes the real ones."""
import argparse
import logging
import math
import os
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from . import shard

LOG = logging.getLogger("glrgtv.train")


# ---------------------------------------------------------------------------------------------------- model factory
def _v13(**kw):
    from . import deep_multiscale_GGLR_GGTV_v1x0 as M
    return M.AbtractMultiScaleGraphFilter(**kw)


def _v7(device=None, **kw):
    from . import model_GLR_GTV_deep_v7 as M
    return M.MultiScaleSequenceDenoiser(device=torch.device(device) if device is not None else torch.device("cuda"), **kw)


def _v1(device=None, **kw):
    from . import model_GLR_GTV_deep_v1 as M
    return M.MultiScaleSequenceDenoiser(device=torch.device(device) if device is not None else torch.device("cuda"), **kw)


# "MultiScaleSequenceDenoiser" = the v7 model that scripts/run_lightformer_GGTV_GGLR_multiblocks.py trains; "..._v1" = the older
# three-block chain (exploration/model_multiscale_mixture_GLR/lib/model_GLR_GTV_deep_v1.py)
MODEL_TYPES: Dict[str, Callable[..., nn.Module]] = {"AbtractMultiScaleGraphFilter": _v13, "MultiScaleSequenceDenoiser": _v7,
                                                    "MultiScaleSequenceDenoiser_v1": _v1}

V13_ARGS = dict(n_channels_in=3, n_channels_out=3, dims=[48, 96, 192, 384], hidden_dims=[96, 192, 384, 768], nsubnets=[1, 1, 1, 1],
                ngraphs=[8, 16, 16, 32], num_blocks=[4, 6, 6, 8], num_blocks_out=4)       # scripts_v2:120-129


def build_model(conf: dict, device=None) -> nn.Module:
    """conf = {"type": <key of MODEL_TYPES>, "args": {...}}; the shipped v13 arguments are the default for its type.
    `device` goes to the family-A constructor, which takes one (v7:1021); other models are moved by the caller."""
    kind = conf.get("type", "AbtractMultiScaleGraphFilter")
    if kind not in MODEL_TYPES:
        raise KeyError(f"model.type {kind!r}: known types are {sorted(MODEL_TYPES)}")
    args = dict(V13_ARGS) if kind == "AbtractMultiScaleGraphFilter" else {}
    args.update(conf.get("args") or {})
    if kind.startswith("MultiScaleSequenceDenoiser") and device is not None:
        args.setdefault("device", device)
    return MODEL_TYPES[kind](**args)


# ---------------------------------------------------------------------------------------------------- optimiser
def build_optimizer(model: nn.Module, conf: Optional[dict] = None):
    """Adam + SequentialLR[MultiStepLR(gamma = 0.5 ** 0.25 every 50k iterations), CosineAnnealingLR] (scripts_v2:147-165)."""
    from torch.optim import Adam
    from torch.optim.lr_scheduler import CosineAnnealingLR, MultiStepLR, SequentialLR
    c = dict(lr=4e-4, eps=1e-8, step_every=50000, n_steps=12, gamma=math.sqrt(math.sqrt(0.5)), cosine_iters=701000,
             cosine_base_lr=5e-5, eta_min=1e-6)
    c.update(conf or {})
    opt = Adam(model.parameters(), lr=c["lr"], eps=c["eps"])
    milestones = [c["step_every"] * (k + 1) for k in range(c["n_steps"])]
    s1 = MultiStepLR(opt, milestones=milestones, gamma=c["gamma"])
    s2 = CosineAnnealingLR(opt, T_max=c["cosine_iters"], eta_min=c["eta_min"])
    s2.base_lrs = [c["cosine_base_lr"] for _ in opt.param_groups]
    return opt, SequentialLR(opt, schedulers=[s1, s2], milestones=[milestones[-1]])


# ---------------------------------------------------------------------------------------------------- loss
def reference_loss(model, noisy_nhwc: torch.Tensor, clean_nhwc: torch.Tensor, w_mse: float = 0.1, w_stab: float = 0.5,
                   latent_sigma: float = 0.05, generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """scripts_v2:191-208: L1(model(noisy), clean) + 0.1 MSE(dec(enc(clean)), clean) + 0.5 MSE(dec(enc(clean)), dec(enc(clean) + N(0, 0.05))).
    Models without encode / decode (family A) get the L1 term only.  Returns (loss, reconstruction NHWC)."""
    recon = model(noisy_nhwc.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    loss = nn.functional.l1_loss(recon, clean_nhwc)
    if hasattr(model, "encode") and hasattr(model, "decode"):
        latent = model.encode(clean_nhwc.permute(0, 3, 1, 2))
        rec_true = model.decode(latent).permute(0, 2, 3, 1)
        disturbed = tuple(z + latent_sigma * torch.randn(z.shape, device=z.device, dtype=z.dtype, generator=generator) for z in latent)
        rec_dist = model.decode(disturbed).permute(0, 2, 3, 1)
        loss = loss + w_mse * nn.functional.mse_loss(rec_true, clean_nhwc) + w_stab * nn.functional.mse_loss(rec_true, rec_dist)
    return loss, recon


# ---------------------------------------------------------------------------------------------------- checkpoints
def checkpoints_folder(conf: dict) -> str:
    return os.path.join(conf["path"]["root_dir"], "experiments", conf["name"], "learning_checkpoints")       # run_train.py:44-45


def checkpoint_name(epoch: int, i: int, verbose_rate: int = 1000) -> str:
    return f"checkpoints_epoch{str(epoch).zfill(2)}_iter{str(i // verbose_rate).zfill(4)}k.pt"                 # scripts_v2:232


def save_checkpoint(folder: str, epoch: int, i: int, model, optimizer, lr_scheduler, verbose_rate: int = 1000) -> str:
    os.makedirs(folder, exist_ok=True)
    path = os.path.join(folder, checkpoint_name(epoch, i, verbose_rate))
    tmp = path + ".tmp"
    torch.save({"i": i, "model": model.state_dict(), "optimizer": optimizer.state_dict(), "lr_scheduler": lr_scheduler.state_dict()}, tmp)
    os.replace(tmp, path)            # a crash while writing never leaves a truncated file as the latest checkpoint
    return path


def latest_checkpoint(folder: str) -> Optional[str]:
    """run_train.py:46-56: the last entry of the sorted folder listing (zero-padded names sort by iteration)"""
    try:
        names = sorted(n for n in os.listdir(folder) if n.endswith(".pt"))
    except OSError:
        names = []
    return os.path.join(folder, names[-1]) if names else None


def load_checkpoint(path: str, model, optimizer=None, lr_scheduler=None, map_location="cpu") -> int:
    state = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(state["model"])
    if optimizer is not None:
        optimizer.load_state_dict(state["optimizer"])
    if lr_scheduler is not None:
        lr_scheduler.load_state_dict(state["lr_scheduler"])
    return int(state["i"])


# ---------------------------------------------------------------------------------------------------- data
class SyntheticNoisyPatches(torch.utils.data.Dataset):
    """(noisy, clean) [H,W,3] float32 pairs, a pure function of (seed, index): clean ~ U(0,1), noisy = clean + sigma/255 N(0,1),
    not clipped (the reference's `addictive_noise_scale` mode)."""

    def __init__(self, patch_size=64, lambda_noise=25.0, max_num_patchs=1000000, seed=2204, **_):
        self.hw = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.sigma, self.n, self.seed = float(lambda_noise) / 255.0, int(max_num_patchs), int(seed)

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 1000003 + int(idx))
        clean = torch.rand(*self.hw, 3, generator=g)
        return clean + self.sigma * torch.randn(*self.hw, 3, generator=g), clean


DATASET_TYPES: Dict[str, Callable[..., torch.utils.data.Dataset]] = {"SyntheticNoisyPatches": SyntheticNoisyPatches}


class ResumableShardedSampler(torch.utils.data.Sampler):
    """The reference's ResumeableSampler (environ/data/data_sampler.py:6-32: sequential, restartable at a sample) sharded over
    ranks: global batch k is samples [k*B*world, (k+1)*B*world), rank r takes the r-th slice of it; `start_batch` skips."""

    def __init__(self, n_samples: int, batch_size: int, rank: int = 0, world: int = 1, start_batch: int = 0):
        self.n, self.bs, self.rank, self.world, self.start = n_samples, batch_size, rank, world, start_batch

    def __len__(self):
        return max(self.n // (self.bs * self.world) - self.start, 0)

    def __iter__(self):
        per = self.bs * self.world
        for k in range(self.start, self.n // per):
            first = k * per + self.rank * self.bs
            yield list(range(first, first + self.bs))


def validate(model, dataset, device, limit: Optional[int] = None) -> float:
    """Mean PSNR (dB, on 0..255 values after clamp + img_as_ubyte, the reference's validation loop scripts_v2:253-287) of `model`
    over `dataset` items (noisy, clean) [H,W,3]; the model is put in eval mode and returned to its previous mode."""
    from . import evalpipe
    was_training = model.training
    model.eval()
    try:
        n = len(dataset) if limit is None else min(limit, len(dataset))
        noisy, clean = [], []
        for i in range(n):
            a, b = dataset[i]
            noisy.append(a.permute(2, 0, 1)[None].to(device))
            clean.append(evalpipe.to_ubyte(b.permute(2, 0, 1)[None].to(device).clamp(0.0, 1.0)))
        return evalpipe.evaluate(model, noisy, clean)
    finally:
        model.train(was_training)


# ---------------------------------------------------------------------------------------------------- the loop
def train(conf: dict, device: Optional[torch.device] = None, on_step: Optional[Callable[[int, float], None]] = None) -> nn.Module:
    """Run (or resume) the experiment `conf` to `train.total_iters` iterations; returns the trained model."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    tc = dict(total_iters=1000, checkpoint_every=5000, log_every=100, verbose_rate=1000, num_workers=0, optimizer=None,
              w_mse=0.1, w_stab=0.5, latent_sigma=0.05, host_cnn_kernels=True, validate_every=0)
    tc.update(conf.get("train") or {})
    # the host CNN's LocalNonLinearBlocks on libglrgtv as well (host_cnn.py; the default) or on the module's own PyTorch layers
    from . import deep_multiscale_GGLR_GGTV_v1x0 as v13
    v13.set_host_cnn_kernels(bool(tc["host_cnn_kernels"]))
    seed = int(conf.get("manual_seed", 2204))
    torch.manual_seed(seed)                                          # same initial weights on every rank
    device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
    model = build_model(conf.get("model") or {}, device).to(device).train()
    optimizer, lr_scheduler = build_optimizer(model, tc["optimizer"])
    folder = checkpoints_folder(conf)
    i = 0
    last = latest_checkpoint(folder)
    if last is not None:
        i = load_checkpoint(last, model, optimizer, lr_scheduler, map_location=device) + 1
        LOG.info("resumed from %s: next iteration %d", last, i)
    dconf = conf["datasets"]["train"]
    dataset = DATASET_TYPES[dconf.get("type", "SyntheticNoisyPatches")](**(dconf.get("dataset_args") or {}))
    bs = int((dconf.get("dataloader_args") or {}).get("batch_size", 4))
    sampler = ResumableShardedSampler(len(dataset), bs, rank, world, start_batch=i % max(len(dataset) // (bs * world), 1))
    loader = torch.utils.data.DataLoader(dataset, batch_sampler=sampler, num_workers=tc["num_workers"])
    gen = torch.Generator(device=device)
    vconf = conf["datasets"].get("val")
    val_set = DATASET_TYPES[vconf.get("type", "SyntheticNoisyPatches")](**(vconf.get("dataset_args") or {})) if vconf else None
    params = [p for p in model.parameters() if p.requires_grad]
    flat = None
    if len(dataset) < bs * world:
        raise ValueError(f"dataset of {len(dataset)} items is smaller than one global batch ({bs} x {world} ranks)")
    while i < tc["total_iters"]:                                     # epochs: the sampler restarts from batch 0 once exhausted
        for noisy, clean in loader:
            if i >= tc["total_iters"]:
                break
            gen.manual_seed(seed + 7919 * i + rank)                      # the latent disturbance: reproducible across resumes
            optimizer.zero_grad(set_to_none=True)
            loss, _ = reference_loss(model, noisy.to(device), clean.to(device), tc["w_mse"], tc["w_stab"], tc["latent_sigma"], gen)
            loss.backward()
            if world > 1:
                flat = shard.allreduce_gradients(params, average=True, flat=flat)
            optimizer.step()
            lr_scheduler.step()
            if on_step is not None:
                on_step(i, float(loss.detach()))
            if rank == 0 and tc["log_every"] and i % tc["log_every"] == 0:
                LOG.info("iter=%d loss=%.6f lr=%.3e", i, float(loss.detach()), optimizer.param_groups[0]["lr"])
            if rank == 0 and val_set is not None and tc["validate_every"] and (i + 1) % tc["validate_every"] == 0:
                psnr = validate(model, val_set, device)                 # rank 0 only: no collective inside (scripts_v2:253-287)
                LOG.info("FINISH VAL - iter=%d - psnr_testing=%.4f", i, psnr)
                if on_step is not None:
                    on_step(i, {"psnr": psnr})
            if world > 1 and val_set is not None and tc["validate_every"] and (i + 1) % tc["validate_every"] == 0:
                dist.barrier()                                           # the other ranks wait here, not inside the next all-reduce
            if rank == 0 and ((i + 1) % tc["checkpoint_every"] == 0 or i + 1 == tc["total_iters"]):
                save_checkpoint(folder, 0, i, model, optimizer, lr_scheduler, tc["verbose_rate"])
            i += 1
        sampler.start = 0
    if world > 1:
        dist.barrier()
    return model


def main(argv=None):
    import yaml
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--conf", required=True, help="experiment YAML (experiment_conf/example.yaml + model / train sections)")
    a = ap.parse_args(argv)
    with open(a.conf) as f:
        conf = yaml.safe_load(f)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(message)s")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    try:
        train(conf)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
