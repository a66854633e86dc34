"""LocalNonLinearBlock inference forward: the PyTorch module (the reference's op sequence) vs host_cnn.py (libglrgtv's
pixel_rstd / dwconv_gate kernels + cuBLAS fp32 GEMMs), CUDA-event timed, per piece.
    python tools/host_cnn_times.py [--dim 48 --hidden 96 --rows 1080 --cols 3840] [--tf32]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imagerestoration_development_unrolling_b200 import deep_multiscale_GGLR_GGTV_v1x0 as M, host_cnn, ops  # noqa: E402


def flag(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    tf32 = "--tf32" in sys.argv
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    dim, hid, H, W = flag("--dim", 48), flag("--hidden", 96), flag("--rows", 1080), flag("--cols", 3840)
    torch.manual_seed(0)
    blk = M.LocalNonLinearBlock(dim, hid, 1).cuda().eval()
    x = torch.randn(1, dim, H, W, device="cuda")
    w1, w9, w2, s0 = host_cnn.folded_weights(blk)
    with torch.no_grad():
        res = {"shape": [1, dim, H, W], "hidden": hid, "tf32": tf32}
        res["module_ms"] = timed(lambda: blk(x))
        host_cnn.GEMM = "cublas"
        res["fused_cublas_ms"] = timed(lambda: host_cnn.nonlinear_block_forward(blk, x))
        ref_c = host_cnn.nonlinear_block_forward(blk, x)
        host_cnn.GEMM = "auto"
        res["fused_ms"] = timed(lambda: host_cnn.nonlinear_block_forward(blk, x))
        res["gemm1_tc_ms"] = timed(lambda: ops.proj_gemm(w1[0], x.view(1, dim, H * W), False))
        rs = ops.pixel_rstd(x, 1, 1e-5)
        h = torch.matmul(w1, x.view(1, 1, dim, H * W)).view(1, -1, H, W)
        u = ops.dwconv_gate(h, rs, w9)
        res["pixel_rstd_ms"] = timed(lambda: ops.pixel_rstd(x, 1, 1e-5))
        res["gemm1_ms"] = timed(lambda: torch.matmul(w1, x.view(1, 1, dim, H * W)))
        res["dwconv_gate_ms"] = timed(lambda: ops.dwconv_gate(h, rs, w9))
        res["gemm2_tc_ms"] = timed(lambda: ops.proj_gemm(w2[0], u.view(1, hid, H * W), False))
        res["gemm2_skip_ms"] = timed(lambda: torch.addcmul(torch.matmul(w2, u.view(1, 1, hid, H * W)).view_as(x), x, s0))
        px = H * W * 4
        res["pixel_rstd_GBs"] = (dim + 1) * px / res["pixel_rstd_ms"] / 1e6
        res["dwconv_gate_GBs"] = (3 * hid + 1) * px / res["dwconv_gate_ms"] / 1e6
        res["rel_err"] = float((host_cnn.nonlinear_block_forward(blk, x) - blk(x)).norm() / blk(x).norm())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
