"""Library fp64 rates on the box (cuBLAS DGEMM/DSYRK-ish, cuSOLVER potrf through torch).
Reported as the "kernel to beat"; not used by the product path."""
import torch, time, json, sys
def ev_time(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    best=1e9; tot=0
    for _ in range(reps):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t=e0.elapsed_time(e1); best=min(best,t); tot+=t
    return best, tot/reps
out={}
for n in (4096, 8192):
    a=torch.randn(n,n,dtype=torch.float64,device='cuda'); b=torch.randn(n,n,dtype=torch.float64,device='cuda')
    best,avg=ev_time(lambda: torch.matmul(a,b))
    out[f"dgemm_{n}"]={"best_ms":best,"avg_ms":avg,"tflops_best":2*n**3/best*1e-9,"tflops_avg":2*n**3/avg*1e-9}
    g=torch.randn(2*n,n,dtype=torch.float64,device='cuda')
    best,avg=ev_time(lambda: torch.matmul(g.t(),g))
    out[f"gtg_{2*n}x{n}"]={"best_ms":best,"avg_ms":avg,"tflops_full_best":2*n*n*2*n/best*1e-9}
    k=torch.matmul(g.t(),g)+torch.eye(n,dtype=torch.float64,device='cuda')
    best,avg=ev_time(lambda: torch.linalg.cholesky(k), reps=3, warm=1)
    out[f"potrf_{n}"]={"best_ms":best,"avg_ms":avg,"tflops_best":n**3/3/best*1e-9}
    del a,b,g,k
print(json.dumps(out,indent=1))
