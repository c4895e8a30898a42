"""Time the tcgen05 GEMM on the path's shapes (CUDA events, L2 flushed between iterations) next to
cuBLAS (torch.matmul) on the same operands.  Usage: python tools/gemm_bench.py [out.json]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from mr_blip_b200 import ops  # noqa: E402

SHAPES = [  # (name, M, N, K, dtype, bn list)
    ("vit_qkv", 61680, 4224, 1408, torch.float16),
    ("vit_proj", 61680, 1408, 1408, torch.float16),
    ("vit_fc1", 61680, 6144, 1408, torch.float16),
    ("vit_fc2", 61680, 1408, 6144, torch.float16),
    ("qf_kv6", 61680, 9216, 1408, torch.float16),
    ("t5_qkv", 8148, 6144, 2080, torch.bfloat16),
    ("t5_wi", 8148, 10240, 2080, torch.bfloat16),
    ("t5_wo", 8148, 2048, 5152, torch.bfloat16),
    ("lm_head", 64, 32128, 2080, torch.bfloat16),
]


def timeit(fn, flush, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    res = []
    for name, M, N, K, dt in SHAPES:
        a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
        b = (torch.randn(N, K, device="cuda") * 0.05).to(dt)
        out = torch.empty(M, N, device="cuda", dtype=dt)
        flops = 2.0 * M * N * K
        row = {"name": name, "M": M, "N": N, "K": K}
        for bn in (0, 128, 192, 256):
            ms = timeit(lambda: ops.gemm(a, b, out=out, force_bn=bn), flush)
            row["bn%d_ms" % bn] = round(ms, 4)
            row["bn%d_tflops" % bn] = round(flops / ms / 1e9, 1)
        ms = timeit(lambda: torch.matmul(a, b.t(), out=out), flush)
        row["cublas_ms"] = round(ms, 4)
        row["cublas_tflops"] = round(flops / ms / 1e9, 1)
        print(row, flush=True)
        res.append(row)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
