"""PSNR golden on a REAL image crop: the reference's AbtractMultiScaleGraphFilter (weights of model_small.npz, CPU fp64) run through
the reference's own evaluation steps (scripts_v2/run_abtract_lightformer_GGTV_GGLR_sigma25.py:253-287) on a 72x104 crop of
exploration/GGTV_GGLR_v1.0/0020.png with sigma = 25 noise from np.random.RandomState(2204) - the fixture SURVEY section 4 names.
The crop is not a multiple of 16, so the reflect padding of :267-271 is exercised.  Run in the build container only."""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/exploration/GGTV_GGLR_v1.0")
import deep_multiscale_GGLR_GGTV_v1x0 as ref  # noqa: E402
from tests.golden.make_golden import to_double  # noqa: E402
from tests.golden.make_golden_model import CFG  # noqa: E402

if __name__ == "__main__":
    z = np.load(os.path.join(HERE, "model_small.npz"))
    m = ref.AbtractMultiScaleGraphFilter(**CFG)
    m.load_state_dict({str(k): torch.from_numpy(z["sd." + str(k)]) for k in z["keys"]}, strict=True)
    m = to_double(m).eval()
    img = np.asarray(Image.open("/root/reference/exploration/GGTV_GGLR_v1.0/0020.png").convert("RGB"))
    clean_u8 = np.ascontiguousarray(img[600:672, 900:1004, :])                       # [72,104,3] uint8
    rs = np.random.RandomState(2204)
    noisy = clean_u8.astype(np.float64) / 255.0 + (25.0 / 255.0) * rs.normal(size=clean_u8.shape)     # not clipped (README.ipynb cell 6)
    x = torch.from_numpy(noisy).permute(2, 0, 1)[None]                                # [1,3,h,w]
    h, w = x.shape[-2:]
    factor = 16
    H, W = ((h + factor) // factor) * factor, ((w + factor) // factor) * factor       # :267-268
    padh, padw = (H - h if h % factor else 0), (W - w if w % factor else 0)
    xp = torch.nn.functional.pad(x, (0, padw, 0, padh), mode="reflect")               # :271
    with torch.no_grad():
        out = m(xp)[:, :, :h, :w]
    out = torch.clamp(out, 0.0, 1.0)[0].permute(1, 2, 0).numpy()
    out_u8 = np.round(out * 255.0).astype(np.uint8)                                   # img_as_ubyte
    mse = np.mean((clean_u8.astype(np.float64) - out_u8.astype(np.float64)) ** 2)
    psnr = 20.0 * np.log10(255.0 / np.sqrt(mse))
    mse_in = np.mean((clean_u8.astype(np.float64) - np.clip(noisy, 0, 1) * 255.0) ** 2)
    np.savez_compressed(os.path.join(HERE, "png_crop_0020.npz"), clean_u8=clean_u8, noisy=noisy.astype(np.float32), out_u8=out_u8,
                        psnr=np.array(psnr), psnr_noisy=np.array(20.0 * np.log10(255.0 / np.sqrt(mse_in))))
    print("psnr", psnr, "noisy input psnr", 20.0 * np.log10(255.0 / np.sqrt(mse_in)), out_u8.shape)
