"""Time the GEMM on the path's own shapes WITH their epilogues (bias / bias+GELU / bias+fp32 residual / plain), back to
back (sustained clocks, operands far larger than L2), plus the decoder-sized (M = 56) and LoRA-down (N = 32) shapes with
forced tile widths.  Kernel variants are selected by environment (MRB_GEMM2_EPI, MRB_GEMM2_TAIL, ...): run once per
variant.  Usage: python tools/gemm_sweep.py tag [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from mr_blip_b200 import ops  # noqa: E402

H, BF = torch.float16, torch.bfloat16
BIG = [  # name, M, N, K, dtype, epilogue
    ("vit_qkv", 61680, 4224, 1408, H, "bias"),
    ("vit_proj", 61680, 1408, 1408, H, "bias_resid"),
    ("vit_fc1", 61680, 6144, 1408, H, "bias_gelu"),
    ("vit_fc1_nogelu", 61680, 6144, 1408, H, "bias"),
    ("vit_fc2", 61680, 1408, 6144, H, "bias_resid"),
    ("qf_kv6", 61680, 9216, 1408, H, "bias"),
    ("t5_qkv", 8192, 6144, 2080, BF, "plain"),
    ("t5_o", 8192, 2048, 2080, BF, "resid"),
    ("t5_wi", 8192, 10240, 2080, BF, "plain"),
    ("t5_wo", 8192, 2048, 5152, BF, "resid"),
    ("t5_dwi", 8192, 2048, 10272, BF, "plain"),
]
SMALL = [  # name, M, N, K, dtype, forced widths
    ("dec_2048", 64, 2048, 2080, BF, (0, 32, 64, 128)),
    ("dec_qkv", 64, 6144, 2080, BF, (0, 32, 64, 128)),
    ("dec_wi", 64, 10240, 2080, BF, (0, 64, 128)),
    ("dec_dwi", 64, 2048, 10272, BF, (0, 32, 64)),
    ("lm_head", 64, 32128, 2080, BF, (0, 64, 128, 256)),
    ("down32", 8192, 32, 2048, BF, (0,)),
    ("down32_k10240", 8192, 32, 10240, BF, (0,)),
]


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def time_graphed(fn, calls=40, reps=5):
    """GPU time per call with the host out of the loop: `calls` launches captured into one CUDA graph, replayed."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (reps * calls)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "default"
    res = {"tag": tag, "env": {k: v for k, v in os.environ.items() if k.startswith("MRB_")}, "big": [], "small": []}
    for name, M, N, K, dt, epi in (BIG if tag != "splitk" else []):
        a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
        b = (torch.randn(N, K, device="cuda") * 0.05).to(dt)
        bias = torch.randn(N, device="cuda") if "bias" in epi else None
        resid = torch.randn(M, N, device="cuda") if "resid" in epi else None
        out = resid if resid is not None else torch.empty(M, N, device="cuda", dtype=dt)
        ms = timeit(lambda: ops.gemm(a, b, out=out, bias=bias, gelu="gelu" in epi, resid=resid), 12)
        row = {"name": name, "M": M, "N": N, "K": K, "epi": epi, "ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
        print(tag, row, flush=True)
        res["big"].append(row)
        del a, b, out, resid
    if tag in ("default", "splitk"):            # MRB_GEMM_SPLITK=1 python tools/gemm_sweep.py splitk: the small shapes again, split along K
        for name, M, N, K, dt, bns in SMALL:
            a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
            # several weight copies so that successive calls do not find their weights in L2 (as in the real step)
            ws = [(torch.randn(N, K, device="cuda") * 0.05).to(dt) for _ in range(max(2, int(3e8 // (N * K * 2)) + 1))]
            out = torch.empty(M, N, device="cuda", dtype=dt)
            row = {"name": name, "M": M, "N": N, "K": K}
            for bn in bns:
                it = [0]

                def fn():
                    it[0] += 1
                    ops.gemm(a, ws[it[0] % len(ws)], out=out, force_bn=bn)
                row["bn%d_us" % bn] = round(time_graphed(fn) * 1e3, 2)
            row["hbm_floor_us"] = round((N * K * 2 + M * K * 2) / 6.5e12 * 1e6, 2)
            print(tag, row, flush=True)
            res["small"].append(row)
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
