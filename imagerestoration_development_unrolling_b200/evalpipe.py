"""Full-image evaluation on the device: the reference's per-image test loop
(scripts_v2/run_abtract_lightformer_GGTV_GGLR_sigma25.py:253-287; README.ipynb cell 10) without the host round trip.

    noisy [1,3,h,w] in [0,1] (+noise, not clipped) -> reflect-pad bottom/right to a multiple of 16 (:267-271)
    -> model under no_grad (:273-274) -> crop (:276) -> clamp to [0,1] (:277) -> img_as_ubyte (:279)
    -> MSE against the uint8 clean image as float (:280) -> PSNR = 20 log10(255 / sqrt(mse)) (:285)

Everything stays a CUDA tensor; one scalar (the PSNR) leaves the device per image.  Plain torch ops - the filter blocks
inside `model` are what runs the CUDA kernels."""
import torch
import torch.nn.functional as F

FACTOR = 16


def pad_to_factor(img: torch.Tensor, factor: int = FACTOR) -> torch.Tensor:
    """the reference's padding rule: H = ((h + f) // f) * f, applied only when h % f != 0 (same for w)"""
    h, w = img.shape[-2:]
    H, W = ((h + factor) // factor) * factor, ((w + factor) // factor) * factor
    padh = H - h if h % factor else 0
    padw = W - w if w % factor else 0
    return F.pad(img, (0, padw, 0, padh), mode="reflect") if (padh or padw) else img


def to_ubyte(x: torch.Tensor) -> torch.Tensor:
    """skimage.img_as_ubyte on a float image in [0,1]: round-half-to-even of 255 x, as float values 0..255"""
    return torch.round(x * 255.0).clamp_(0.0, 255.0)


def inference_executor(model, rank: int = 0, world: int = 1, group=None):
    """The network as a callable for inference (`restore_image(inference_executor(model), noisy)`): same results as
    `model(...)`, with the host CNN's LocalNonLinearBlocks on libglrgtv's kernels (host_cnn.py) and, for world > 1, the image
    cut into row strips across ranks (shard.ShardedMultiScaleFilter; pass this rank's strip)."""
    from . import shard
    return shard.ShardedMultiScaleFilter(model, rank, world, group)


@torch.no_grad()
def restore_image(model, noisy: torch.Tensor) -> torch.Tensor:
    """noisy [1,3,h,w] float32 on the model's device -> restored image quantised to 0..255 (float tensor, [1,3,h,w])"""
    h, w = noisy.shape[-2:]
    out = model(pad_to_factor(noisy))
    return to_ubyte(torch.clamp(out[:, :, :h, :w], 0.0, 1.0))


def psnr_255(restored_255: torch.Tensor, clean_255: torch.Tensor) -> torch.Tensor:
    """PSNR in dB between two 0..255-valued images (a 0-dim tensor on the device)"""
    mse = torch.mean((clean_255.float() - restored_255.float()) ** 2)
    return 20.0 * torch.log10(255.0 / torch.sqrt(mse))


@torch.no_grad()
def evaluate(model, noisy_images, clean_images_255) -> float:
    """mean PSNR over an iterable of (noisy [1,3,h,w] in [0,1], clean [1,3,h,w] in 0..255); one host sync at the end"""
    vals = [psnr_255(restore_image(model, n), c) for n, c in zip(noisy_images, clean_images_255)]
    return float(torch.stack(vals).mean())
