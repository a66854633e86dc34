"""imagerestoration_development_unrolling_b200/evalpipe.py against a numpy restatement of the reference's test loop
(scripts_v2/run_abtract_lightformer_GGTV_GGLR_sigma25.py:253-287), with a stand-in model (plain torch, runs on CPU)."""
import numpy as np
import pytest
import torch

from imagerestoration_development_unrolling_b200 import evalpipe as E


class Smooth(torch.nn.Module):
    """a deterministic stand-in for the denoiser: 3x3 box blur with replicate borders"""

    def forward(self, x):
        assert x.shape[-2] % 16 == 0 and x.shape[-1] % 16 == 0      # the model only ever sees padded images
        xp = torch.nn.functional.pad(x, (1, 1, 1, 1), mode="replicate")
        return torch.nn.functional.avg_pool2d(xp, 3, stride=1)


def numpy_loop(img_true_255, noisy):
    h, w = noisy.shape[:2]
    f = 16
    H, W = ((h + f) // f) * f, ((w + f) // f) * f
    padh = H - h if h % f != 0 else 0
    padw = W - w if w % f != 0 else 0
    x = np.pad(noisy, ((0, padh), (0, padw), (0, 0)), mode="reflect")
    xp = np.pad(x, ((1, 1), (1, 1), (0, 0)), mode="edge")
    out = sum(xp[i:i + x.shape[0], j:j + x.shape[1]] for i in range(3) for j in range(3)) / 9.0
    restored = np.clip(out[:h, :w], 0, 1)
    restored = np.clip(np.rint(restored * 255.0), 0, 255).astype(np.float32)       # skimage.img_as_ubyte
    mse = np.square(img_true_255 - restored).mean()
    return restored, 20 * np.log10(255.0 / np.sqrt(mse))


@pytest.mark.parametrize("hw", [(37, 50), (32, 48), (16, 33), (70, 64)])
def test_restore_and_psnr_match_the_reference_loop(hw):
    h, w = hw
    rs = np.random.RandomState(2204)
    clean_255 = rs.randint(0, 256, size=(h, w, 3)).astype(np.float32)
    noisy = (clean_255 / 255.0 + rs.normal(0, 25.0 / 255.0, clean_255.shape)).astype(np.float32)
    ref_img, ref_psnr = numpy_loop(clean_255, noisy.astype(np.float64))
    n = torch.from_numpy(noisy).permute(2, 0, 1).unsqueeze(0)
    c = torch.from_numpy(clean_255).permute(2, 0, 1).unsqueeze(0)
    out = E.restore_image(Smooth(), n)
    got = out[0].permute(1, 2, 0).numpy()
    assert np.abs(got - ref_img).max() <= 1.0                       # fp32 vs fp64 blur: at most one grey level, on ties only
    assert (got != ref_img).mean() < 1e-3
    assert abs(float(E.psnr_255(out, c)) - ref_psnr) < 0.01          # BASELINE north_star: PSNR within 0.01 dB
    assert abs(E.evaluate(Smooth(), [n], [c]) - ref_psnr) < 0.01


def test_padding_rule():
    for h, w in [(16, 16), (17, 31), (2040 // 8, 1392 // 8)]:
        x = torch.zeros(1, 3, h, w)
        p = E.pad_to_factor(x)
        assert p.shape[-2] % 16 == 0 and p.shape[-1] % 16 == 0
        assert p.shape[-2] - h < 16 and p.shape[-1] - w < 16
        assert (h % 16 == 0) == (p.shape[-2] == h)


def test_inference_executor_keeps_the_model_surface():
    """evalpipe.inference_executor(model) is the reference's surface (encode / filtering / decode / enc_dec / call); on CPU tensors
    the host CNN stays on the PyTorch modules, so enc_dec must reproduce the module exactly"""
    import torch
    from imagerestoration_development_unrolling_b200 import evalpipe, deep_multiscale_GGLR_GGTV_v1x0 as M
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(5)
        m = M.AbtractMultiScaleGraphFilter(dims=[8, 8, 8, 8], hidden_dims=[8, 8, 8, 8], ngraphs=[2, 2, 2, 2], num_blocks=[1, 1, 1, 1], num_blocks_out=1).eval()
        img = torch.rand(1, 3, 32, 48)
    net = evalpipe.inference_executor(m)
    for name in ("encode", "filtering", "decode", "enc_dec", "forward"):
        assert callable(getattr(net, name))
    with torch.no_grad():
        assert torch.equal(net.enc_dec(img), m.enc_dec(img))
        lat_a, lat_b = net.encode(img), m.encode(img)
    assert len(lat_a) == 4 and all(torch.equal(a, b) for a, b in zip(lat_a, lat_b))
