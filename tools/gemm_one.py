"""ViT fc1 GEMM (61680 x 6144 x 1408, bias + GELU, fp16) a few times: target for `ncu --set full -k regex:gemm2`."""
import sys, torch
sys.path.insert(0, ".")
from mr_blip_b200 import ops
M, N, K = 61680, 6144, 1408
a = (torch.randn(M, K, device="cuda") * 0.5).half(); b = (torch.randn(N, K, device="cuda") * 0.05).half()
bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.float16)
for _ in range(4):
    ops.gemm(a, b, out=out, bias=bias, gelu=True)
torch.cuda.synchronize()
print("ok")
