"""Check and time the tcgen05 projection GEMMs (csrc/proj_tc.cu) against cuBLAS fp32 / fp64 on the bench shapes.

    python tools/proj_times.py check     # accuracy on small + ragged shapes, one subprocess per shape
    python tools/proj_times.py           # timings on the bench shapes (JSON lines)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHECK = [(1, 16, 8, 128), (1, 96, 48, 256), (2, 96, 48, 4096), (1, 48, 192, 1024), (2, 768, 384, 256), (3, 384, 1536, 256),
         (2, 24, 12, 36), (1, 8, 32, 20), (2, 96, 48, 10416)]
BENCH = [(32, 96, 48, 65536), (32, 48, 192, 16384), (32, 96, 48, 16384), (32, 192, 96, 16384), (32, 96, 384, 4096), (32, 384, 192, 4096),
         (32, 768, 384, 1024), (32, 384, 1536, 256)]


def one(shape):
    import torch
    from imagerestoration_development_unrolling_b200 import ops, _lib as L
    L.load().glrgtv_set_proj_pipeline(int(os.environ.get("PROJ_PIPELINE", "0")))
    B, M, K, N = shape
    gen = torch.Generator().manual_seed(M + K)
    w = torch.randn(M, K, generator=gen).cuda()
    x = torch.randn(B, K, N, generator=gen).cuda()
    gy = torch.randn(B, M, N, generator=gen).cuda()
    rel = lambda a, b: float((a.double() - b).norm() / b.norm())
    w64, x64, gy64 = w.double(), x.double(), gy.double()
    r = {"shape": shape}
    r["fwd"] = rel(ops.proj_gemm(w, x, False), torch.einsum("mk,bkn->bmn", w64, x64))
    r["dgrad"] = rel(ops.proj_gemm(w, gy, True), torch.einsum("mk,bmn->bkn", w64, gy64))
    r["wgrad"] = rel(ops.proj_wgrad(gy, x), torch.einsum("bmn,bkn->mk", gy64, x64))
    torch.cuda.synchronize()
    print(json.dumps(r), flush=True)


def bench():
    import torch
    from imagerestoration_development_unrolling_b200 import ops, _lib as L
    torch.backends.cuda.matmul.allow_tf32 = False
    pipe = int(os.environ.get("PROJ_PIPELINE", "0"))        # 0 = in-place stages (default), 1 = landing ring
    L.load().glrgtv_set_proj_pipeline(pipe)

    def t(f, n=10):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for (B, M, K, N) in BENCH:
        w = torch.randn(M, K, device="cuda")
        x = torch.randn(B, K, N, device="cuda")
        gy = torch.randn(B, M, N, device="cuda")
        we = w.unsqueeze(0).expand(B, -1, -1)
        r = {"pipeline": pipe, "shape": (B, M, K, N), "GFLOP": round(2 * B * M * K * N / 1e9, 1)}
        r["fwd_tc"] = t(lambda: ops.proj_gemm(w, x, False))
        r["fwd_GBs"] = round(4 * B * N * (M + K) / r["fwd_tc"] / 1e6, 0)
        r["fwd_bmm"] = t(lambda: torch.bmm(we, x))
        r["dgrad_tc"] = t(lambda: ops.proj_gemm(w, gy, True))
        r["dgrad_bmm"] = t(lambda: torch.bmm(we.transpose(1, 2), gy))
        r["wgrad_tc"] = t(lambda: ops.proj_wgrad(gy, x))
        r["wgrad_GBs"] = round(4 * B * N * (M + K) / r["wgrad_tc"] / 1e6, 0)
        r["wgrad_bmmsum"] = t(lambda: torch.bmm(gy, x.transpose(1, 2)).sum(0))
        print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        one(tuple(int(v) for v in sys.argv[2].split(",")))
    elif len(sys.argv) > 1 and sys.argv[1] == "check":
        bad = 0
        for sh in CHECK:
            p = subprocess.run([sys.executable, __file__, "one", ",".join(map(str, sh))], capture_output=True, text=True, timeout=120)
            print(p.stdout.strip() or f"{sh}: rc={p.returncode} {p.stderr.strip()[-400:]}", flush=True)
            bad += p.returncode != 0
        sys.exit(1 if bad else 0)
    else:
        bench()
