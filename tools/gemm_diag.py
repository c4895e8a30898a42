"""The path's large GEMM shapes, this library against cuBLAS (torch.matmul) on the same operands, ALTERNATING on one box so that both
see the same clocks / thermal state.  python tools/gemm_diag.py [out.json] [--ncu]   (--ncu: one launch of each between
cudaProfilerStart / Stop for `ncu --profile-from-start off`; shapes then limited to MRB_DIAG_SHAPES, default vit_qkv,t5_qkv)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from mr_blip_b200 import ops  # noqa: E402

SHAPES = [  # name, M, N, K, dtype, epilogue
    ("vit_qkv", 61680, 4224, 1408, torch.float16, "bias"),
    ("vit_proj", 61680, 1408, 1408, torch.float16, "bias_resid"),
    ("vit_fc1", 61680, 6144, 1408, torch.float16, "bias_gelu"),
    ("vit_fc1_nogelu", 61680, 6144, 1408, torch.float16, "bias"),
    ("vit_fc2", 61680, 1408, 6144, torch.float16, "bias_resid"),
    ("t5_qkv", 8192, 6144, 2080, torch.bfloat16, "plain"),
    ("t5_o", 8192, 2048, 2080, torch.bfloat16, "f32"),
    ("t5_wi", 8192, 10240, 2080, torch.bfloat16, "plain"),
    ("t5_wo", 8192, 2048, 5152, torch.bfloat16, "f32"),
    ("t5_dwi", 8192, 2048, 10272, torch.bfloat16, "plain"),
    ("t5_dqkv", 8192, 2048, 6176, torch.bfloat16, "plain"),
]


def timeit(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out_path = next((a for a in sys.argv[1:] if not a.startswith("--")), None)
    ncu = "--ncu" in sys.argv
    only = os.environ.get("MRB_DIAG_SHAPES", "vit_qkv,t5_qkv" if ncu else "").split(",")
    rows = []
    for name, M, N, K, dt, epi in SHAPES:
        if only != [""] and name not in only:
            continue
        a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
        b = (torch.randn(N, K, device="cuda") * 0.05).to(dt)
        bias = torch.randn(N, device="cuda") if "bias" in epi else None
        f32 = "resid" in epi or epi == "f32"
        out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else dt)
        ref = torch.empty(M, N, device="cuda", dtype=dt)
        kw = dict(out=out, bias=bias, gelu="gelu" in epi)
        if "resid" in epi:
            kw["resid"] = out
        ours = lambda: ops.gemm(a, b, **kw)  # noqa: E731
        cub = lambda: torch.matmul(a, b.t(), out=ref)  # noqa: E731
        for _ in range(3):
            ours(); cub()
        torch.cuda.synchronize()
        if ncu:
            torch.cuda.cudart().cudaProfilerStart()
            ours(); cub()
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
            continue
        reps = max(10, int(20e-3 / (2.0 * M * N * K / 1.1e15)))       # ~20 ms per burst
        t_o, t_c = [], []
        for _ in range(4):
            t_o.append(timeit(ours, reps)); t_c.append(timeit(cub, reps))
        fl = 2.0 * M * N * K
        mo, mc = sorted(t_o)[1], sorted(t_c)[1]
        row = dict(name=name, M=M, N=N, K=K, epi=epi, ms=round(mo, 4), tflops=round(fl / mo / 1e9, 1), cublas_ms=round(mc, 4),
                   cublas_tflops=round(fl / mc / 1e9, 1), ratio=round(mc / mo, 3))
        print(os.environ.get("MRB_LIB_VARIANT", "default"), row, flush=True)
        rows.append(row)
        del a, b, out, ref
    if out_path and rows:
        json.dump({"variant": os.environ.get("MRB_LIB_VARIANT", ""), "rows": rows}, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
