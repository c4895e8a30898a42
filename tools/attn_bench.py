"""Time the attention kernels (tcgen05 vs mma.sync) on the path's two big shapes. python tools/attn_bench.py [shape] [impl]
MRB_ATTN_BENCH_DROP=1 adds the train-mode (probability dropout 0.1) variants of the T5 shapes next to the eval-mode ones."""
import os
import sys

import torch

sys.path.insert(0, ".")
from mr_blip_b200 import ops  # noqa: E402


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


DROP = os.environ.get("MRB_ATTN_BENCH_DROP", "0") == "1"
WORD = None


def main():
    global WORD
    WORD = torch.tensor([12345], dtype=torch.int32, device="cuda")
    only = sys.argv[1] if len(sys.argv) > 1 else None
    impls = (sys.argv[2],) if len(sys.argv) > 2 else ("mma", "tc")
    for name, B, H, L, hd, dt, with_bias in [("vit", 240, 16, 257, 88, torch.float16, False),
                                             ("t5enc", 4, 32, 2037, 64, torch.bfloat16, True),
                                             ("t5enc_nobias", 4, 32, 2037, 64, torch.bfloat16, False),
                                             ("t5enc_4017", 2, 32, 4017, 64, torch.bfloat16, True)]:
        if only and name != only:
            continue
        qkv = (torch.randn(B, L, 3, H, hd, device="cuda") * 0.5).to(dt)
        out = torch.empty(B, L, H, hd, device="cuda", dtype=dt)
        rs = 3 * H * hd
        bias = None
        if with_bias:       # T5-style: bucketed relative bias saturates 128 positions off the diagonal (modeling_t5.py:393-445)
            idx = (torch.arange(2 * L - 1, device="cuda") - (L - 1)).clamp(-128, 128) + (L - 1)
            bias = torch.randn(H, 2 * L - 1, device="cuda")[:, idx].contiguous()
        kmask = torch.ones(B, L, dtype=torch.int32, device="cuda") if with_bias else None
        flops = 4.0 * B * H * L * L * hd
        if name == "vit" and "tc" in impls:                       # the persistent ViT kernel (all 257 rows, CLS included)
            ms = timeit(lambda: ops.attention_vit(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, B, H, L, hd, hd ** -0.5,
                                                  (L * rs, rs), (L * rs, rs), (L * rs, rs), (L * H * hd, H * hd)))
            print("%-14s %-4s %8.3f ms  %7.1f TFLOP/s" % (name, "vit", ms, flops / ms / 1e9), flush=True)
        for impl in impls:
            Lq = L if impl == "mma" else (L // 128) * 128 if name == "vit" else L
            ms = timeit(lambda: ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, B, H, Lq, L, hd, hd ** -0.5,
                                                  (L * rs, rs), (L * rs, rs), (L * rs, rs), (L * H * hd, H * hd), bias=bias,
                                                  bias_zero=L - 1, kmask=kmask, impl=impl))
            print("%-14s %-4s %8.3f ms  %7.1f TFLOP/s" % (name, impl, ms, flops / ms / 1e9), flush=True)
            if DROP and dt == torch.bfloat16:
                ms = timeit(lambda: ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, B, H, Lq, L, hd, hd ** -0.5,
                                                      (L * rs, rs), (L * rs, rs), (L * rs, rs), (L * H * hd, H * hd), bias=bias,
                                                      bias_zero=L - 1, kmask=kmask, impl=impl, drop=(WORD, 0x41, 0.1)))
                print("%-14s %-4s %8.3f ms  %7.1f TFLOP/s  (dropout 0.1)" % (name, impl, ms, flops / ms / 1e9), flush=True)
        if name.startswith("t5enc") and "tc" in impls:          # backward (dK/dV kernel + dQ kernel + delta), 10 L^2 d flops
            q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
            lse = torch.empty(B, H, L, device="cuda")
            ops.attention_fwd(q, k, v, out, B, H, L, L, hd, hd ** -0.5, (L * rs, rs), (L * rs, rs), (L * rs, rs),
                              (L * H * hd, H * hd), bias=bias, bias_zero=L - 1, kmask=kmask, lse=lse, impl="tc")
            dout = torch.randn_like(out)
            dqkv = torch.empty_like(qkv)
            ws = torch.empty(B * H * L, device="cuda")
            st, ost = (L * rs, rs), (L * H * hd, H * hd)
            ms = timeit(lambda: ops.attention_bwd(q, k, v, out, dout, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], B, H, L, L, hd,
                                                  hd ** -0.5, st, st, st, ost, ost, lse, ws, bias=bias, bias_zero=L - 1,
                                                  kmask=kmask, impl="tc"))
            print("%-14s bwd  %8.3f ms  %7.1f TFLOP/s (algorithmic 2.5x fwd)" % (name, ms, 2.5 * flops / ms / 1e9), flush=True)
            if DROP:
                drop = (WORD, 0x41, 0.1)
                ops.attention_fwd(q, k, v, out, B, H, L, L, hd, hd ** -0.5, (L * rs, rs), (L * rs, rs), (L * rs, rs),
                                  (L * H * hd, H * hd), bias=bias, bias_zero=L - 1, kmask=kmask, lse=lse, impl="tc", drop=drop)
                ms = timeit(lambda: ops.attention_bwd(q, k, v, out, dout, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], B, H, L, L, hd,
                                                      hd ** -0.5, st, st, st, ost, ost, lse, ws, bias=bias, bias_zero=L - 1,
                                                      kmask=kmask, impl="tc", drop=drop))
                print("%-14s bwd  %8.3f ms  %7.1f TFLOP/s (dropout 0.1)" % (name, ms, 2.5 * flops / ms / 1e9), flush=True)


def cross():
    """The T5 decoder's cross-attention (16 target rows x 2037 encoder keys per clip, 32 heads, kmask, bf16), forward and backward,
    24 calls (= the decoder's layers) per CUDA-graph replay: us per call in-graph, few-query kernels (auto) vs the tcgen05 kernels."""
    B, H, Lq, Lk, hd = 4, 32, 16, 2037, 64
    word = torch.tensor([12345], dtype=torch.int32, device="cuda")
    q = (torch.randn(B, Lq, H, hd, device="cuda") * 0.5).bfloat16()
    kv = [(torch.randn(B, Lk, 2, H, hd, device="cuda") * 0.5).bfloat16() for _ in range(24)]       # one K/V per layer: 1.6 GB, not L2-resident
    out, dout = torch.empty_like(q), torch.randn_like(q)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv[0])
    kmask = torch.ones(B, Lk, dtype=torch.int32, device="cuda")
    lse = torch.empty(B, H, Lq, device="cuda")
    ws = torch.empty(B * H * Lq, device="cuda")
    qs, ks = (Lq * H * hd, H * hd), (Lk * 2 * H * hd, 2 * H * hd)
    for drop in ((None, (word, 0x41, 0.1)) if DROP else (None,)):
        for impl in ("auto", "tc"):
            def fwd():
                for t in kv:
                    ops.attention_fwd(q, t[:, :, 0], t[:, :, 1], out, B, H, Lq, Lk, hd, 1.0, qs, ks, ks, qs, kmask=kmask, lse=lse, impl=impl, drop=drop)

            def bwd():
                for t in kv:
                    ops.attention_bwd(q, t[:, :, 0], t[:, :, 1], out, dout, dq, dkv[:, :, 0], dkv[:, :, 1], B, H, Lq, Lk, hd, 1.0, qs, ks, ks,
                                      qs, qs, lse, ws, kmask=kmask, impl=impl, drop=drop)
            for nm, fn in (("fwd", fwd), ("bwd", bwd)):
                fn()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                ms = timeit(g.replay)
                print("t5dec_cross    %-4s %-4s %8.1f us per layer%s" % (impl, nm, ms * 1e3 / len(kv), "  (dropout 0.1)" if drop else ""), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cross":
        cross()
    else:
        main()
