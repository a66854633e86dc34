"""Whole drop-in model (host CNN in PyTorch + the four CUDA filter blocks) against a golden produced by the reference's
AbtractMultiScaleGraphFilter; PSNR parity (BASELINE north_star: within 0.01 dB); torch.compile keeps working."""
import os

import numpy as np
import pytest
import torch

from tests.util import rel

CFG = dict(n_channels_in=3, n_channels_out=3, dims=[12, 24, 24, 48], hidden_dims=[24, 48, 48, 96], nsubnets=[1, 1, 1, 1],
           ngraphs=[2, 4, 2, 4], num_blocks=[1, 1, 1, 1], num_blocks_out=1)


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "model_small.npz"))
    return z


def test_model_state_dict_keys_and_order_match_reference(golden_dir):
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    z = _load(golden_dir)
    m = M.AbtractMultiScaleGraphFilter(**CFG)
    assert list(m.state_dict().keys()) == [str(k) for k in z["keys"]]
    m.load_state_dict({str(k): torch.from_numpy(z["sd." + str(k)]) for k in z["keys"]}, strict=True)


@pytest.mark.gpu
def test_model_output_and_psnr_match_reference(golden_dir):
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    z = _load(golden_dir)
    m = M.AbtractMultiScaleGraphFilter(**CFG)
    m.load_state_dict({str(k): torch.from_numpy(z["sd." + str(k)]) for k in z["keys"]}, strict=True)
    m = m.cuda().eval()
    noisy = torch.from_numpy(z["noisy"]).float().cuda()
    clean = torch.from_numpy(z["clean"]).float().cuda()
    with torch.no_grad():
        out = m(noisy)
        enc = m.enc_dec(noisy)
    assert rel(out, torch.from_numpy(z["out"])) < 1e-4
    assert rel(enc, torch.from_numpy(z["enc_dec"])) < 1e-5
    psnr = float(10 * torch.log10(1.0 / ((out - clean) ** 2).mean()))
    assert abs(psnr - float(z["psnr"])) < 0.01, (psnr, float(z["psnr"]))


@pytest.mark.gpu
def test_compiled_block_matches_eager():
    """reference scripts call model.compile(); the fused op is an opaque custom op with a fake kernel and autograd"""
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    torch.manual_seed(0)
    blk = M.LocalLowpassFilteringBlock(12, 1, 2).cuda()
    x = torch.randn(2, 12, 16, 24, device="cuda", requires_grad=True)
    g = torch.randn_like(x)
    out = blk(x)
    gx, = torch.autograd.grad(out, x, g)
    cblk = torch.compile(blk, backend="aot_eager")       # traces through Dynamo + AOTAutograd without a codegen backend
    x2 = x.detach().clone().requires_grad_(True)
    out2 = cblk(x2)
    gx2, = torch.autograd.grad(out2, x2, g)
    assert rel(out2, out) < 1e-6 and rel(gx2, gx) < 1e-5


@pytest.mark.gpu
def test_png_crop_psnr_within_0p01_db_of_the_reference(golden_dir):
    """The evaluation pipeline on the device (evalpipe.restore_image / psnr_255: reflect-pad to 16, model, crop, clamp,
    img_as_ubyte, PSNR) on a real-image crop of the reference's 0020.png with sigma = 25 noise from RandomState(2204), against
    the reference model run through the reference's own evaluation steps in fp64 (tests/golden/make_golden_png.py)."""
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    from imagerestoration_development_unrolling_b200 import evalpipe
    z = _load(golden_dir)
    g = np.load(os.path.join(golden_dir, "png_crop_0020.npz"))
    m = M.AbtractMultiScaleGraphFilter(**CFG)
    m.load_state_dict({str(k): torch.from_numpy(z["sd." + str(k)]) for k in z["keys"]}, strict=True)
    m = m.cuda().eval()
    noisy = torch.from_numpy(g["noisy"]).permute(2, 0, 1)[None].contiguous().cuda()            # [1,3,72,104]: padded to 80x112 inside
    clean = torch.from_numpy(g["clean_u8"].astype(np.float32)).permute(2, 0, 1)[None].cuda()
    restored = evalpipe.restore_image(m, noisy)
    psnr = float(evalpipe.psnr_255(restored, clean))
    assert abs(psnr - float(g["psnr"])) < 0.01, (psnr, float(g["psnr"]))
    ref_u8 = torch.from_numpy(g["out_u8"].astype(np.float32)).permute(2, 0, 1)[None].cuda()
    diff = (restored - ref_u8).abs()
    assert float(diff.max()) <= 1.0 and float((diff > 0).float().mean()) < 0.02      # quantisation ties only
    # the sharded executor gives the same image (its W % 4 / thin-strip guards fall back to the module path)
    restored2 = evalpipe.restore_image(evalpipe.inference_executor(m), noisy)
    assert float((restored2 - restored).abs().max()) <= 1.0
    # evaluate() = the mean of psnr_255 over the image list, one host read
    assert abs(evalpipe.evaluate(m, [noisy, noisy], [clean, clean]) - psnr) < 1e-4


@pytest.mark.gpu
def test_executor_on_a_cbsd68_sized_image(golden_dir):
    """CBSD68 images are 321 x 481 -> padded to 336 x 496: the 1/8-scale planes are 42 x 62 (W % 4 == 2), outside the host-CNN
    kernels and the streaming filter kernels.  The inference executor must fall back per scale and give the module's result
    (shard.py)."""
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    from imagerestoration_development_unrolling_b200 import evalpipe
    z = _load(golden_dir)
    m = M.AbtractMultiScaleGraphFilter(**CFG)
    m.load_state_dict({str(k): torch.from_numpy(z["sd." + str(k)]) for k in z["keys"]}, strict=True)
    m = m.cuda().eval()
    noisy = torch.rand(1, 3, 321, 481, generator=torch.Generator().manual_seed(68)).cuda()
    with torch.no_grad():
        padded = evalpipe.pad_to_factor(noisy)
        assert padded.shape[-2:] == (336, 496)
        ref = m(padded)
        out = evalpipe.inference_executor(m)(padded)
    assert rel(out, ref) < 1e-4, rel(out, ref)
    a, b = evalpipe.restore_image(m, noisy), evalpipe.restore_image(evalpipe.inference_executor(m), noisy)
    assert a.shape == (1, 3, 321, 481) and float((a - b).abs().max()) <= 1.0


@pytest.mark.gpu
def test_four_stream_filtering_equals_the_sequential_blocks(golden_dir):
    """set_filter_streams(True), the default: the four filter blocks on four streams (V1X0:1117-1131: independent) - the same outputs bit for
    bit, the same gradients up to the order of the floating-point atomics, also when captured into a CUDA graph"""
    from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M
    z = _load(golden_dir)
    m = M.AbtractMultiScaleGraphFilter(**CFG)
    m.load_state_dict({str(k): torch.from_numpy(z["sd." + str(k)]) for k in z["keys"]}, strict=True)
    m = m.cuda().train()
    noisy = torch.from_numpy(z["noisy"]).float().cuda()
    params = [p for p in m.parameters()]

    def run():
        out = m(noisy)
        grads = torch.autograd.grad(out.square().mean(), params)
        return out.detach(), grads

    prev = M.set_filter_streams(False)
    try:
        out1, g1 = run()
        M.set_filter_streams(True)
        out2, g2 = run()
        with torch.no_grad():
            m(noisy)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out3 = m(noisy)
            graph.replay()
            torch.cuda.synchronize()
    finally:
        M.set_filter_streams(prev)
    assert torch.equal(out1, out2) and torch.equal(out3, out1)
    for a, b in zip(g1, g2):
        assert rel(b, a) < 1e-4 or float(a.abs().max()) < 1e-12
