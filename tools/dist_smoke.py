import os, sys, time, torch
sys.path.insert(0, ".")
t0 = time.time()
def log(*a): print("[rank %s %.1fs]" % (os.environ.get("RANK"), time.time() - t0), *a, flush=True)
from mr_blip_b200 import dist as mdist, ops, _lib
log("init...")
rank, world, local = mdist.init_distributed_mode()
log("init done", rank, world, local, torch.cuda.current_device())
import torch.distributed as tdist
x = torch.ones(1 << 20, device="cuda") * (rank + 1)
tdist.all_reduce(x)
torch.cuda.synchronize()
log("allreduce ok", x[0].item())
a = torch.randn(512, 256, device="cuda").half(); b = torch.randn(256, 256, device="cuda").half()
o = ops.gemm(a, b, out_dtype=torch.float32)
torch.cuda.synchronize()
log("gemm ok", (o - a.float() @ b.float().t()).abs().max().item())
p = [torch.nn.Parameter(torch.zeros(1000, device="cuda")) for _ in range(3)]
for q in p: q.grad = torch.full_like(q, float(rank))
red = mdist.GradAllReducer(p); red(); torch.cuda.synchronize()
log("reducer ok", p[0].grad[0].item())
tdist.barrier(); log("barrier ok")
tdist.destroy_process_group(); log("done")
