"""Parity tests of the split-K path of the small-M / 32-column GEMMs (csrc/gemm.cu mrb_gemm_splitk: K ranges per CTA, fp32 partials
in a workspace, ordered reduce pass with the epilogue) against an fp32 torch reference, and of a whole training step + generate
with it on and off.  First run on a B200 in round 2 (profiles/r02_call1.md); split-K is the default since (MRB_GEMM_SPLITK=0
switches it off)."""
import math
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mr_blip_b200 import _lib
    _lib.load()
    return _lib


def _rand(shape, dtype, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def _splitk(lib, a, b, out, ws, M, K, bias=None, gelu=False, resid=None, force_bn=0, max_splits=8):
    from mr_blip_b200.ops import _DT, _ptr
    N = b.shape[0]
    lib.call("mrb_gemm_splitk", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), M, N, K, _DT[a.dtype], _ptr(bias), int(gelu),
             _ptr(resid), resid.stride(0) if resid is not None else 0, out.data_ptr(), _DT[out.dtype], out.stride(0), 0, force_bn,
             ws.data_ptr(), ws.numel() * 4, max_splits, torch.cuda.current_stream().cuda_stream)


def _close(got, want, rtol, atol, what):
    err = (got.float() - want.float()).abs()
    bad = (err > atol + rtol * want.float().abs()).sum().item()
    assert bad == 0, "%s: %d/%d out of tolerance, max err %.4g" % (what, bad, err.numel(), err.max().item())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("M,N,K,bn,ms", [
    (56, 2048, 2080, 0, 8),       # decoder q/k/v/o: K tail block (2080 = 32.5 x 64), 7 splits of 128-wide tiles
    (56, 2048, 2080, 64, 4),      # forced 64-wide tiles, 4 splits
    (56, 6144, 2080, 0, 8),       # fused qkv
    (56, 10240, 2080, 0, 8),      # wi_0 | wi_1: 256-wide tiles
    (64, 2048, 10272, 0, 8),      # wi dgrad: long K
    (5, 2048, 2080, 0, 8),        # one clip, beam search
    (128, 2048, 5152, 0, 3),      # full row tile, 3 splits
    (8132, 32, 2048, 0, 8),       # LoRA down-projection: 64 row tiles x 2 splits
    (8132, 32, 10240, 0, 8),
    (56, 32128, 2080, 0, 8),      # lm_head: enough tiles, must take the unsplit path
    (56, 2048, 192, 0, 8),        # 3 K blocks: too short to split
])
def test_splitk_matches_fp32_reference(lib, dtype, M, N, K, bn, ms):
    a = _rand((M, K), dtype, 1.0, 1)
    b = _rand((N, K), dtype, 1.0 / math.sqrt(K), 2)
    ws = torch.full((48 << 18,), float("nan"), device="cuda")            # NaN-poisoned: every partial read must have been written
    want = a.float() @ b.float().t()
    out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    _splitk(lib, a, b, out, ws, M, K, force_bn=bn, max_splits=ms)
    _close(out, want, 2e-3, 2e-3, "fp32 out")
    out_h = torch.empty((M, N), dtype=dtype, device="cuda")
    _splitk(lib, a, b, out_h, ws, M, K, force_bn=bn, max_splits=ms)
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-3
    _close(out_h, want, tol, tol, "16-bit out")
    # agreement with the unsplit kernel up to fp32 summation order
    from mr_blip_b200 import ops
    _close(out, ops.gemm(a, b, out_dtype=torch.float32, force_bn=bn), 1e-4, 1e-4, "vs mrb_gemm")


def test_splitk_epilogues_strides_and_small_workspace(lib):
    M, N, K = 56, 2048, 2080
    big = _rand((M, K + 40), torch.bfloat16, 1.0, 3)                      # row stride > K
    a = big[:, :K]
    b = _rand((N, K), torch.bfloat16, 1.0 / math.sqrt(K), 4)
    bias = _rand((N,), torch.float32, 1.0, 5)
    resid = _rand((M, N), torch.float32, 1.0, 6)
    ws = torch.full((8 * M * N,), float("nan"), device="cuda")
    ref = a.float() @ b.float().t() + bias
    out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    _splitk(lib, a, b, out, ws, M, K, bias=bias)
    _close(out, ref, 2e-3, 2e-3, "bias")
    _splitk(lib, a, b, out, ws, M, K, bias=bias, gelu=True)
    _close(out, torch.nn.functional.gelu(ref), 2e-3, 2e-3, "bias + gelu")
    x = resid.clone()
    _splitk(lib, a, b, x, ws, M, K, resid=x)                              # in-place fp32 residual stream (T5 o / wo)
    _close(x, resid + a.float() @ b.float().t(), 2e-3, 2e-3, "residual in place")
    outbuf = torch.zeros((M, N + 64), dtype=torch.bfloat16, device="cuda")
    _splitk(lib, a, b, outbuf[:, :N], ws, M, K)
    _close(outbuf[:, :N], a.float() @ b.float().t(), 1e-2, 1e-2, "strided 16-bit out")
    assert outbuf[:, N:].abs().max().item() == 0
    # a workspace too small for two splits: silently the unsplit kernel, the workspace stays untouched
    tiny = torch.full((M * N,), float("nan"), device="cuda")
    _splitk(lib, a, b, out, tiny, M, K)
    _close(out, a.float() @ b.float().t(), 2e-3, 2e-3, "small workspace")
    assert torch.isnan(tiny).all()
    # deterministic: fixed summation order
    o1 = torch.empty((M, N), dtype=torch.float32, device="cuda")
    o2 = torch.empty_like(o1)
    _splitk(lib, a, b, o1, ws, M, K)
    _splitk(lib, a, b, o2, ws, M, K)
    assert torch.equal(o1, o2)


def test_splitk_whole_model_step(monkeypatch, tiny_sd):
    """The tiny-config (full widths, 2 layers) training step and generate with split-K switched on, against the same model
    with it off: same loss, LoRA / t5_proj gradients and decoded strings up to fp32 summation order."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mr_blip_b200 import ops
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from mr_blip_b200.dims import TINY
    from oracle import synth
    samples = synth.make_samples(batch=2, frames=4, seed=0)

    def run(flag):
        monkeypatch.setattr(ops, "SPLITK", flag)
        m = BLIP2_MR(dims=TINY, state_dict=tiny_sd, cuda_graphs=False, train_dropout=False).cuda().train()
        if flag:
            vit, qf, t5 = m.engines()
            ops.splitk_register(t5.side)
            ops.splitk_register(vit.side)
        loss = m(samples)["loss"]
        loss.backward()
        grads = torch.cat([p.grad.flatten() for p in m.parameters() if p.requires_grad])
        with torch.no_grad():                                # decode path (M = clips x beams rows): first-step logits of both beams
            logits = m.eval().forward_mr(samples, want_logits=True)["logits"].float()
            pred = m.generate(samples, num_beams=2, max_length=8)["raw_prediction"]
        return loss.item(), grads, (logits, pred)

    l0, g0, p0 = run(False)
    l1, g1, p1 = run(True)
    assert abs(l0 - l1) <= 2e-4 * abs(l0), (l0, l1)
    # a different fp32 summation order moves the bf16 roundings of everything downstream: compare in norm, and elementwise
    # against the scale of the gradient (B200, round 2: 3.3 % of the elements moved by more than 2 % of their own value, the
    # largest absolute difference was 8.1e-4 = 0.35 % of the largest gradient element)
    assert ((g1 - g0).norm() / g0.norm()).item() < 1e-2
    _close(g1, g0, 2e-2, 1e-2 * g0.abs().max().item(), "grads")
    # on random weights the beam search sits on near-ties, so decoded strings may legitimately differ between two summation orders
    # (they did on the B200); the logits they are decoded from must agree
    assert ((p1[0] - p0[0]).norm() / p0[0].norm()).item() < 1e-2      # B200: 4e-3 through 2 + 2 bf16 layers
    assert len(p0[1]) == len(p1[1])


_FOLD_SCRIPT = r"""
import math, sys, torch
sys.path.insert(0, sys.argv[2])
from mr_blip_b200 import _lib
from mr_blip_b200.ops import _DT, _ptr
_lib.load()
outs = []
for i, (M, N, K, dt, gelu, use_bias, use_resid, od) in enumerate([
        (56, 2048, 2080, torch.bfloat16, False, False, True, torch.float32), (64, 10240, 2080, torch.bfloat16, False, False, False, torch.bfloat16),
        (64, 2048, 10272, torch.bfloat16, False, True, False, torch.float32), (5, 2048, 2080, torch.float16, True, True, False, torch.float16),
        (8132, 32, 2048, torch.bfloat16, False, False, False, torch.bfloat16), (128, 2048, 5152, torch.bfloat16, False, True, True, torch.float32)]):
    g = torch.Generator(device="cuda").manual_seed(100 + i)
    a = torch.randn((M, K), generator=g, device="cuda").to(dt)
    b = (torch.randn((N, K), generator=g, device="cuda") / math.sqrt(K)).to(dt)
    bias = torch.randn((N,), generator=g, device="cuda") if use_bias else None
    resid = torch.randn((M, N), generator=g, device="cuda") if use_resid else None
    ws = torch.full((48 << 18,), float("nan"), device="cuda")
    out = torch.empty((M, N), dtype=od, device="cuda")
    for rep in range(3):          # the arrival counters must come back to zero: repeated launches on one workspace agree
        _lib.call("mrb_gemm_splitk", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), M, N, K, _DT[dt], _ptr(bias), int(gelu),
                  _ptr(resid), resid.stride(0) if resid is not None else 0, out.data_ptr(), _DT[od], out.stride(0), 0, 0,
                  ws.data_ptr(), ws.numel() * 4, 8, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        if rep == 0:
            first = out.clone()
        assert torch.equal(out, first), (i, rep)
    outs.append(out.cpu())
torch.save(outs, sys.argv[1])
"""


def test_folded_reduce_is_bit_identical_to_the_two_pass_reduce(lib, tmp_path):
    """The reduce folded into the split-K launch (the CTA that arrives last at an output tile sums the partials in split order and
    applies the epilogue; csrc/gemm.cu, default) against the separate splitk_reduce_kernel launch (MRB_SPLITK_FUSED=0): the same
    arithmetic in the same order, so the outputs must be EQUAL -- for every epilogue, and launch after launch on one workspace."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "fold.py"
    script.write_text(_FOLD_SCRIPT)
    res = {}
    for fused in ("1", "0"):
        env = dict(os.environ, MRB_SPLITK_FUSED=fused)
        f = str(tmp_path / ("out%s.pt" % fused))
        subprocess.check_call([sys.executable, str(script), f, root], env=env)
        res[fused] = torch.load(f)
    for x, y in zip(res["1"], res["0"]):
        assert torch.isfinite(x.float()).all()
        assert torch.equal(x, y)
