"""Whole-model golden from the REFERENCE's AbtractMultiScaleGraphFilter (small configuration, CPU fp64):
state-dict key order, output on a synthetic noisy image, PSNR.  Run in the build container only."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/exploration/GGTV_GGLR_v1.0")
import deep_multiscale_GGLR_GGTV_v1x0 as ref  # noqa: E402
from tests.golden.make_golden import to_double  # noqa: E402

CFG = dict(n_channels_in=3, n_channels_out=3, dims=[12, 24, 24, 48], hidden_dims=[24, 48, 48, 96], nsubnets=[1, 1, 1, 1],
           ngraphs=[2, 4, 2, 4], num_blocks=[1, 1, 1, 1], num_blocks_out=1)

if __name__ == "__main__":
    torch.manual_seed(5)
    m = ref.AbtractMultiScaleGraphFilter(**CFG)
    gen = torch.Generator().manual_seed(6)
    with torch.no_grad():      # make the filter blocks matter (default init is ~identity)
        for name, p in m.named_parameters():
            if "localfilter" in name and name.endswith(("muys00", "muys01", "ro00", "ro01")):
                p.copy_(torch.log(0.01 + 0.04 * torch.rand(p.shape, generator=gen)))
            elif "localfilter" in name and name.endswith(("gamma00", "gamma01")):
                p.copy_(torch.log(0.01 + 0.3 * torch.rand(p.shape, generator=gen)))
            elif "localfilter" in name and ("multiM" in name or "stats_kernel_p" in name):
                p.add_(0.2 * torch.randn(p.shape, generator=gen))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = to_double(m)
    clean = torch.rand(1, 3, 48, 64, generator=gen, dtype=torch.float64)
    noisy = clean + (25.0 / 255.0) * torch.randn(clean.shape, generator=gen, dtype=torch.float64)
    with torch.no_grad():
        out = m(noisy)
        enc = m.enc_dec(noisy)
    psnr = float(10 * torch.log10(1.0 / ((out - clean) ** 2).mean()))
    rec = {"clean": clean.numpy(), "noisy": noisy.numpy(), "out": out.numpy(), "enc_dec": enc.numpy(), "psnr": np.array(psnr),
           "keys": np.array(list(sd.keys()))}
    for k, v in sd.items():
        rec["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "model_small.npz"), **rec)
    print("written", len(sd), "tensors; psnr", psnr)
