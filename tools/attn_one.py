"""T5 encoder attention forward + backward once (L 2037, T5-style bias that saturates 128 positions off the diagonal + mask, bf16;
MRB_ATTN_ONE_DROP=1: with probability dropout 0.1): target for ncu --set full."""
import os
import sys
import torch
sys.path.insert(0, ".")
from mr_blip_b200 import ops
B, H, L, hd = 4, 32, 2037, 64
qkv = (torch.randn(B, L, 3, H, hd, device="cuda") * 0.5).bfloat16()
out = torch.empty(B, L, H, hd, device="cuda", dtype=torch.bfloat16)
rs = 3 * H * hd
idx = (torch.arange(2 * L - 1, device="cuda") - (L - 1)).clamp(-128, 128) + (L - 1)
bias = torch.randn(H, 2 * L - 1, device="cuda")[:, idx].contiguous()
drop = (torch.tensor([12345], dtype=torch.int32, device="cuda"), 0x41, 0.1) if os.environ.get("MRB_ATTN_ONE_DROP", "0") == "1" else None
kmask = torch.ones(B, L, dtype=torch.int32, device="cuda")
lse = torch.empty(B, H, L, device="cuda")
st, ost = (L * rs, rs), (L * H * hd, H * hd)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
dout = torch.randn_like(out); dqkv = torch.empty_like(qkv); ws = torch.empty(B * H * L, device="cuda")
for _ in range(2):
    ops.attention_fwd(q, k, v, out, B, H, L, L, hd, 1.0, st, st, st, ost, bias=bias, bias_zero=L - 1, kmask=kmask, lse=lse, impl="tc", drop=drop)
    ops.attention_bwd(q, k, v, out, dout, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], B, H, L, L, hd, 1.0, st, st, st, ost, ost,
                      lse, ws, bias=bias, bias_zero=L - 1, kmask=kmask, impl="tc", drop=drop)
torch.cuda.synchronize()
print("ok")
