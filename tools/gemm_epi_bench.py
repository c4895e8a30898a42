import sys, torch
sys.path.insert(0, ".")
from mr_blip_b200 import ops
def timeit(fn, flush, iters=5):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts)//2]
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
M=61680
for name,N,K in [("fc1",6144,1408),("proj",1408,1408),("fc2",1408,6144),("qkv",4224,1408)]:
    a=(torch.randn(M,K,device="cuda")*0.5).half(); b=(torch.randn(N,K,device="cuda")*0.05).half()
    bias=torch.randn(N,device="cuda"); res=torch.randn(M,N,device="cuda")
    o16=torch.empty(M,N,device="cuda",dtype=torch.float16); o32=torch.empty(M,N,device="cuda")
    fl=2.0*M*N*K
    for label,fn in [("plain16",lambda: ops.gemm(a,b,out=o16)),("bias16",lambda: ops.gemm(a,b,out=o16,bias=bias)),
                     ("gelu16",lambda: ops.gemm(a,b,out=o16,bias=bias,gelu=True)),("plain32",lambda: ops.gemm(a,b,out=o32)),
                     ("resid32",lambda: ops.gemm(a,b,out=res,bias=bias,resid=res)),("cublas16",lambda: torch.matmul(a,b.t(),out=o16))]:
        ms=timeit(fn,flush); print("%-5s %-9s %7.3f ms %7.1f TF/s" % (name,label,ms,fl/ms/1e9), flush=True)
