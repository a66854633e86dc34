import torch, time, sys
sys.path.insert(0,'/root/repo')
from imagerestoration_development_unrolling_b200 import ops
torch.backends.cuda.matmul.allow_tf32=False
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
for (B,M,K,N) in [(32,96,48,65536),(32,48,192,16384),(32,192,96,16384),(32,768,384,1024),(32,384,1536,256)]:
    w=torch.randn(M,K,device='cuda'); x=torch.randn(B,K,N,device='cuda'); gy=torch.randn(B,M,N,device='cuda')
    we=w.unsqueeze(0).expand(B,-1,-1)
    r={}
    r['fwd_tc']=t(lambda: ops.proj_gemm(w,x,False)); r['fwd_bmm']=t(lambda: torch.bmm(we,x))
    r['dgrad_tc']=t(lambda: ops.proj_gemm(w,gy,True)); r['dgrad_bmm']=t(lambda: torch.bmm(we.transpose(1,2),gy))
    r['wgrad_bmmsum']=t(lambda: torch.bmm(gy,x.transpose(1,2)).sum(0))
    if ops.proj_wgrad_supported(M,K,N): r['wgrad_own']=t(lambda: ops.proj_wgrad(gy,x))
    gf=2*B*M*K*N/1e9
    print((B,M,K,N), f"{gf:.1f} GFLOP", {k: round(v,3) for k,v in r.items()})
