"""The Q-Former cross-attention path alone (batched K/V projection GEMM of the 6 cross-attention layers + the 6 attention cores,
QVH shapes: 240 frames x 257 tokens, train-mode dropout of the probabilities), eager launches: target for
`ncu --set full -k regex:"attn_xq|gemm2" -s 7 -c 2` (second pass: the GEMM and the first core).  Prints the graph-timed ms too."""
import sys

import torch

sys.path.insert(0, ".")
from mr_blip_b200 import dropout as dr, ops  # noqa: E402
from mr_blip_b200.dims import FULL  # noqa: E402

d = FULL
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 240
Hq, nq, T, heads, ncross = d.qf_hidden, d.num_query, d.vit_tokens, d.qf_heads, 6
hd, M = Hq // heads, frames * nq
g = torch.Generator(device="cuda").manual_seed(0)
ie16 = torch.randn((frames * T, d.vit_width), device="cuda", generator=g).half()
kv_w = (torch.randn((ncross * 2 * Hq, d.vit_width), device="cuda", generator=g) * 0.03).half()
kv_b = torch.randn((ncross * 2 * Hq,), device="cuda", generator=g)
qc = (torch.randn((M, Hq), device="cuda", generator=g) * 0.5).half()
ctx = torch.empty((M, Hq), dtype=torch.float16, device="cuda")
kv = torch.empty((frames * T, ncross * 2 * Hq), dtype=torch.float16, device="cuda")
kv_rs = kv.shape[1]
drop = dr.DropState()


def path():
    ops.gemm(ie16, kv_w, out=kv, bias=kv_b)
    for c in range(ncross):
        kbase = kv[:, c * 2 * Hq:]
        ops.attention_fwd(qc, kbase, kbase[:, Hq:], ctx, frames, heads, nq, T, hd, hd ** -0.5, (nq * Hq, Hq), (T * kv_rs, kv_rs),
                          (T * kv_rs, kv_rs), (nq * Hq, Hq), drop=drop.attn(dr.site(dr.QF, 2 * c, dr.CROSS_P), drop.qformer))


for _ in range(3):
    path()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    path()
e1.record()
torch.cuda.synchronize()
print("xattn path, 10 eager passes back to back: %.3f ms per pass" % (e0.elapsed_time(e1) / 10))
