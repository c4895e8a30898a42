"""Kernel-level parity on the GPU: every C-ABI entry point against a plain PyTorch fp32 evaluation
of the same op on the same (already rounded) 16-bit inputs.  Tolerances are stated per test."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mr_blip_b200 import ops as _ops, _lib
    _lib.load()
    return _ops


def _rand(shape, dtype, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def _close(got, want, rtol, atol, what=""):
    got, want = got.float(), want.float()
    err = (got - want).abs()
    bound = atol + rtol * want.abs()
    bad = (err > bound).sum().item()
    assert bad == 0, "%s: %d/%d out of tolerance, max err %.4g (max |want| %.4g)" % (
        what, bad, err.numel(), err.max().item(), want.abs().max().item())


# ------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 256),        # one tile, one k-block
    (128, 256, 256, 256),       # k pipeline
    (300, 384, 200, 128),       # ragged M, K not a multiple of 64 (TMA zero fill)
    (1000, 1408, 1408, 0),      # ViT proj shape (N = 5.5 x 256)
    (777, 4224, 1408, 0),       # ViT qkv -> 192-wide tiles
    (2048, 6144, 1408, 0),      # ViT fc1
    (514, 1408, 6144, 0),       # ViT fc2, long K
    (64, 32128, 2048, 0),       # lm_head
    (4100, 2048, 2080, 0),      # T5 linear with LoRA-extended K
    (56, 2048, 2080, 0),        # decoder step: narrow tiles spread over the SMs
    (56, 2048, 2080, 64),
    (56, 2048, 2080, 32),       # M <= 64: 64-row A stages, deeper TMA ring
    (64, 2048, 10272, 0),       # decoder wi dgrad: long K at tiny M
    (8132, 32, 2048, 0),        # LoRA down-projection
])
def test_gemm_plain(ops, dtype, M, N, K, bn):
    a = _rand((M, K), dtype, 1.0, 1)
    b = _rand((N, K), dtype, 1.0 / math.sqrt(K), 2)
    out = ops.gemm(a, b, out_dtype=torch.float32, force_bn=bn)
    want = a.float() @ b.float().t()
    _close(out, want, 2e-3, 2e-3, "gemm fp32 out")      # fp32 accumulate; only summation order differs
    out_h = ops.gemm(a, b, force_bn=bn)
    _close(out_h, want, 1e-2 if dtype == torch.bfloat16 else 2e-3, 1e-2 if dtype == torch.bfloat16 else 2e-3, "gemm 16-bit out")


@pytest.mark.parametrize("M", [771, 4100])          # 4100 rows route through the 2-CTA (cta_group::2) kernel
def test_gemm_epilogues(ops, M):
    N, K = 1408, 1408
    a = _rand((M, K), torch.float16, 1.0, 3)
    b = _rand((N, K), torch.float16, 1.0 / math.sqrt(K), 4)
    bias = _rand((N,), torch.float32, 1.0, 5)
    resid = _rand((M, N), torch.float32, 1.0, 6)
    ref = a.float() @ b.float().t() + bias
    _close(ops.gemm(a, b, bias=bias, out_dtype=torch.float32), ref, 2e-3, 2e-3, "bias")
    _close(ops.gemm(a, b, bias=bias, gelu=True, out_dtype=torch.float32), torch.nn.functional.gelu(ref), 2e-3, 2e-3, "gelu")
    _close(ops.gemm(a, b, bias=bias, gelu=True), torch.nn.functional.gelu(ref), 4e-3, 4e-3, "gelu fp16 out (ViT fc1 epilogue)")
    _close(ops.gemm(a, b, bias=bias), ref, 4e-3, 4e-3, "bias fp16 out (ViT qkv epilogue)")
    x = resid.clone()
    ops.gemm(a, b, out=x, bias=bias, resid=x)            # in-place fp32 residual stream
    _close(x, resid + ref, 2e-3, 2e-3, "residual in place")
    # strided A (row stride > K) and strided output
    big = _rand((M, K + 72), torch.float16, 1.0, 7)
    outbuf = torch.zeros((M, N + 64), dtype=torch.float16, device="cuda")
    ops.gemm(big[:, :K], b, out=outbuf[:, :N])
    _close(outbuf[:, :N], big[:, :K].float() @ b.float().t(), 3e-3, 3e-3, "strided")
    assert outbuf[:, N:].abs().max().item() == 0


@pytest.mark.parametrize("F", [3, 20])
def test_gemm_patch_embed_remap(ops, F):
    G, C, K = 256, 1408, 592
    a = _rand((F * G, K), torch.float16, 1.0, 8)
    w = _rand((C, K), torch.float16, 0.05, 9)
    bias = _rand((C,), torch.float32, 1.0, 10)
    pos = _rand((G + 1, C), torch.float32, 1.0, 11)
    x = torch.zeros((F * (G + 1), C), dtype=torch.float32, device="cuda")
    ops.gemm(a, w, out=x, bias=bias, resid=pos, row_group=G)
    want = (a.float() @ w.float().t() + bias).view(F, G, C) + pos[1:]
    got = x.view(F, G + 1, C)
    _close(got[:, 1:], want, 2e-3, 2e-3, "patch rows")
    assert got[:, 0].abs().max().item() == 0            # cls slot untouched


# ------------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, scale, bias=None, kmask=None, causal=False, q_pos0=0):
    # q [B,Lq,H,hd] etc, fp32 math
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    Lq, Lk = s.shape[-2:]
    if bias is not None:
        s = s + bias
    if kmask is not None:
        s = s.masked_fill(kmask[:, None, None, :] == 0, float("-inf"))
    if causal:
        i = torch.arange(Lq, device=s.device)[:, None] + q_pos0
        j = torch.arange(Lk, device=s.device)[None, :]
        s = s.masked_fill(j > i, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, vf).permute(0, 2, 1, 3), torch.logsumexp(s, dim=-1)


@pytest.mark.parametrize("impl", ["mma", "tc"])
@pytest.mark.parametrize("B,H,L,hd,dtype", [(3, 16, 257, 88, torch.float16), (2, 12, 32, 64, torch.float16),
                                            (2, 32, 200, 64, torch.bfloat16), (2, 4, 640, 64, torch.bfloat16),
                                            (1, 2, 129, 80, torch.float16)])
def test_attention_fwd_fused_qkv_layout(ops, B, H, L, hd, dtype, impl):
    qkv = _rand((B, L, 3, H, hd), dtype, 1.0, 12)
    out = torch.zeros((B, L, H, hd), dtype=dtype, device="cuda")
    rs = 3 * H * hd
    ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, B, H, L, L, hd, hd ** -0.5,
                      (L * rs, rs), (L * rs, rs), (L * rs, rs), (L * H * hd, H * hd), impl=impl)
    want, _ = _attn_ref(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], hd ** -0.5)
    tol = 2e-2 if dtype == torch.bfloat16 else 4e-3       # P is rounded to 16 bit before P.V
    _close(out, want, tol, tol, "attention fwd")


@pytest.mark.parametrize("frames,H,hd,dtype,peaked", [(3, 16, 88, torch.float16, 0), (1, 1, 88, torch.float16, 0), (10, 16, 88, torch.float16, 0),
                                                      (40, 16, 88, torch.float16, 1), (2, 4, 96, torch.bfloat16, 0), (2, 3, 72, torch.float16, 2),
                                                      (5, 16, 88, torch.bfloat16, 1)])
def test_attention_vit_persistent_kernel(ops, frames, H, hd, dtype, peaked):
    """csrc/attention_vit.cu (eva_vit.py:128-145): all 257 query rows -- CLS as key and as query on CUDA cores, the 256 patches as
    128 x 256 tcgen05 tiles with P in tensor memory -- against an fp32 softmax.  Item counts below, at and far above the number
    of SMs (persistent loop, barrier phases, triple-buffered CLS rows); `peaked` plants scores far above the first keys' (the
    online-softmax correction of the written P) -- 1: in the later key chunks, 2: on the CLS key and on the last patch key."""
    L = 257
    qkv = _rand((frames, L, 3, H, hd), dtype, 1.0, 12 + frames)
    if peaked == 1:
        for j in (40, 100, 130, 200, 255):                    # one query direction meets keys of growing alignment further down the row
            qkv[:, j, 1] = qkv[:, 7, 0] * (0.4 + j / 200.0)
    if peaked == 2:
        qkv[:, 0, 1] = qkv[:, 9, 0] * 2.0
        qkv[:, 256, 1] = qkv[:, 9, 0] * 3.0
    out = torch.full((frames, L, H, hd), 7.0, dtype=dtype, device="cuda")
    rs = 3 * H * hd
    ops.attention_vit(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], out, frames, H, L, hd, hd ** -0.5,
                      (L * rs, rs), (L * rs, rs), (L * rs, rs), (L * H * hd, H * hd))
    want, _ = _attn_ref(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], hd ** -0.5)
    tol = 2e-2 if dtype == torch.bfloat16 else 4e-3           # P is rounded to 16 bit before P.V
    _close(out[:, 0], want[:, 0], tol, tol, "CLS query row")
    _close(out[:, 1:129], want[:, 1:129], tol, tol, "patch rows, group 0")
    _close(out[:, 129:], want[:, 129:], tol, tol, "patch rows, group 1")
    # same answer as the round-1 path (generic flash kernel, all rows)
    old = torch.zeros_like(out)
    ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], old, frames, H, L, L, hd, hd ** -0.5,
                      (L * rs, rs), (L * rs, rs), (L * rs, rs), (L * H * hd, H * hd), impl="mma")
    _close(out, old, 2 * tol, 2 * tol, "vs mma.sync kernel")


def test_attention_single_row(ops):
    B, H, L, hd = 5, 16, 257, 88
    qkv = _rand((B, L, 3, H, hd), torch.float16, 1.0, 42)
    out = torch.zeros((B, L, H, hd), dtype=torch.float16, device="cuda")
    rs = 3 * H * hd
    ops.attention_row(qkv[:, L - 1, 0], qkv[:, :, 1], qkv[:, :, 2], out[:, L - 1], B, H, L, hd, hd ** -0.5,
                      L * rs, (L * rs, rs), (L * rs, rs), L * H * hd)
    want, _ = _attn_ref(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], hd ** -0.5)
    _close(out[:, L - 1], want[:, L - 1], 4e-3, 4e-3, "single row")
    assert out[:, :L - 1].abs().max().item() == 0


@pytest.mark.parametrize("Lq,Lk,dtype", [(32, 257, torch.float16), (20, 100, torch.bfloat16), (32, 320, torch.float16),
                                         (7, 65, torch.float16), (32, 321, torch.float16)])
def test_attention_cross_small_q(ops, Lq, Lk, dtype):
    """Q-Former cross-attention shape (32 queries x 257 keys): the one-shot K/V kernel (keys split over the warps,
    partial softmax merged through shared memory); Lk = 321 falls back to the generic kernel."""
    B, H, hd = 5, 12, 64
    q = _rand((B, Lq, H, hd), dtype, 1.0, 13)
    kv = _rand((B, Lk, 2, H, hd), dtype, 1.0, 14)
    out = torch.zeros_like(q)
    ops.attention_fwd(q, kv[:, :, 0], kv[:, :, 1], out, B, H, Lq, Lk, hd, 0.125, (Lq * H * hd, H * hd),
                      (Lk * 2 * H * hd, 2 * H * hd), (Lk * 2 * H * hd, 2 * H * hd), (Lq * H * hd, H * hd))
    want, _ = _attn_ref(q, kv[:, :, 0], kv[:, :, 1], 0.125)
    tol = 2e-2 if dtype == torch.bfloat16 else 4e-3
    _close(out, want, tol, tol, "cross attention")


@pytest.mark.parametrize("impl", ["mma", "tc"])
@pytest.mark.parametrize("Lq,Lk,causal,sat", [(150, 150, False, 0), (21, 21, True, 0), (21, 333, False, 0), (300, 300, True, 0),
                                              (257, 400, False, 0), (513, 513, False, 0), (700, 700, False, 128),
                                              (700, 700, False, 40), (14, 650, False, 128), (16, 2037, False, 0),
                                              (30, 1000, False, 0)])
def test_attention_t5_bias_mask_fwd_bwd(ops, Lq, Lk, causal, sat, impl):
    """sat > 0: T5-style bias that saturates `sat` positions off the diagonal (bucketed relative positions,
    modeling_t5.py:393-445), so most KV tiles see one bias value -- the constant-bias fast paths of the tcgen05 kernels."""
    B, H, hd = 2, 32, 64
    dt = torch.bfloat16
    q, k, v = (_rand((B, L, H, hd), dt, 0.5, s) for L, s in ((Lq, 15), (Lk, 16), (Lk, 17)))
    dout = _rand((B, Lq, H, hd), dt, 1.0, 18)
    table = _rand((H, Lq + Lk - 1), torch.float32, 1.0, 19)          # bias by (j - i) + (Lq - 1)
    if sat:
        idx = (torch.arange(Lq + Lk - 1, device="cuda") - (Lq - 1)).clamp(-sat, sat) + (Lq - 1)
        table = table[:, idx.clamp(0, Lq + Lk - 2)].contiguous()
    kmask = torch.ones((B, Lk), dtype=torch.int32, device="cuda")
    kmask[1, Lk - 7:] = 0
    i = torch.arange(Lq, device="cuda")[:, None]
    j = torch.arange(Lk, device="cuda")[None, :]
    bias_full = table[:, (j - i) + (Lq - 1)][None]                   # [1,H,Lq,Lk]
    out = torch.zeros_like(q)
    lse = torch.zeros((B, H, Lq), dtype=torch.float32, device="cuda")
    st = lambda L: (L * H * hd, H * hd)
    ops.attention_fwd(q, k, v, out, B, H, Lq, Lk, hd, 1.0, st(Lq), st(Lk), st(Lk), st(Lq), bias=table, bias_zero=Lq - 1,
                      kmask=kmask, causal=causal, lse=lse, impl=impl)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    want, want_lse = _attn_ref(qr, kr, vr, 1.0, bias_full, kmask, causal)
    _close(out, want, 2e-2, 2e-2, "t5 attention fwd")
    _close(lse, want_lse, 1e-3, 1e-3, "lse")
    want.backward(dout.float())
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    ws = torch.zeros((B * H * Lq,), dtype=torch.float32, device="cuda")
    ops.attention_bwd(q, k, v, out, dout, dq, dk, dv, B, H, Lq, Lk, hd, 1.0, st(Lq), st(Lk), st(Lk), st(Lq), st(Lq),
                      lse, ws, bias=table, bias_zero=Lq - 1, kmask=kmask, causal=causal, impl=impl)
    # 16-bit P/dS operands: errors scale with the gradient magnitude
    for got, ref, nm in ((dq, qr.grad, "dq"), (dk, kr.grad, "dk"), (dv, vr.grad, "dv")):
        _close(got, ref, 3e-2, 3e-2 * ref.abs().max().item(), nm)


# ------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("C,mode", [(1408, 0), (768, 0), (2048, 1)])
def test_norm(ops, C, mode):
    rows = 1003
    x = _rand((rows, C), torch.float32, 2.0, 20) + 0.3
    add = _rand((rows, C), torch.float32, 1.0, 21)
    w = _rand((C,), torch.float32, 0.2, 22) + 1.0
    b = _rand((C,), torch.float32, 0.2, 23) if mode == 0 else None
    o32 = torch.empty_like(x)
    oh = torch.empty((rows, C + 32), dtype=torch.bfloat16, device="cuda")
    so = torch.empty_like(x)
    ops.norm(x, w, b, 1e-6, mode, add=add, out_f32=o32, out_h=oh, sum_out=so)
    s = x + add
    if mode == 0:
        want = torch.nn.functional.layer_norm(s, (C,), w, b, 1e-6)
    else:
        want = w * (s * torch.rsqrt(s.pow(2).mean(-1, keepdim=True) + 1e-6))
    _close(o32, want, 1e-5, 1e-5, "norm fp32")
    _close(oh[:, :C], want, 8e-3, 8e-3, "norm bf16")
    assert torch.equal(so, s)


def test_rmsnorm_bwd(ops):
    rows, C = 517, 2048
    x = _rand((rows, C), torch.float32, 2.0, 24).requires_grad_(True)
    w = _rand((C,), torch.float32, 0.2, 25) + 1.0
    dy = _rand((rows, C), torch.float32, 1.0, 26)
    y = w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6))
    y.backward(dy)
    dres = _rand((rows, C), torch.float32, 1.0, 27)
    want = dres + x.grad
    ops.rmsnorm_bwd(x.detach(), w, dy, 1e-6, dres)
    _close(dres, want, 1e-4, 1e-5, "rmsnorm bwd")
    # 16-bit extended dgrad buffer with the LoRA term folded in: dy_eff = dy + dy[:, C:C+R] . A
    R = 24
    ext = _rand((rows, C + 32), torch.bfloat16, 1.0, 40)
    A = _rand((R, C), torch.float32, 0.05, 41)
    dy_eff = ext[:, :C].float() + ext[:, C:C + R].float() @ A
    x2 = x.detach().clone().requires_grad_(True)
    y2 = w * (x2 * torch.rsqrt(x2.pow(2).mean(-1, keepdim=True) + 1e-6))
    y2.backward(dy_eff)
    dres2 = torch.zeros((rows, C), device="cuda")
    ops.rmsnorm_bwd(x.detach(), w, ext, 1e-6, dres2, lora_A=A, R=R)
    _close(dres2, x2.grad, 1e-4, 1e-5, "rmsnorm bwd + lora")
    # standalone fix-up kernel, in place and accumulating
    e2 = ext.clone()
    ops.lora_up_add(e2, A, R, rows, C)
    _close(e2[:, :C], dy_eff, 8e-3, 8e-3, "lora_up_add in place")
    acc = torch.ones((rows, C), device="cuda")
    ops.lora_up_add(ext.clone(), A, R, rows, C, acc=acc)
    _close(acc, dy_eff + 1.0, 1e-5, 1e-5, "lora_up_add acc")


@pytest.mark.parametrize("rows", [64, 8192, 517])
def test_fused_residual_kernels_equal_their_two_pass_forms(ops, rows):
    """mrb_dropout_add_norm == mrb_dropout_add then mrb_norm(mode 1), mrb_rmsnorm_bwd_drop == mrb_rmsnorm_bwd then mrb_dropout
    (fp32 -> bf16): bit for bit on the device, decoder- and encoder-sized (the train step fuses them from 512 rows)."""
    C, p, site = 2048, 0.1, 0x1046
    word = torch.tensor([0x1234567], dtype=torch.int32, device="cuda")
    x, br = _rand((rows, C), torch.float32, 1.5, 51), _rand((rows, C), torch.float32, 1.0, 52)
    w = _rand((C,), torch.float32, 0.2, 53) + 1.0
    out1, xn1 = torch.empty_like(x), torch.zeros((rows, C + 32), dtype=torch.bfloat16, device="cuda")
    ops.dropout_add(x, br, out1, word, site, p)
    ops.norm(out1, w, None, 1e-6, 1, out_h=xn1)
    out2, xn2 = torch.empty_like(x), torch.zeros((rows, C + 32), dtype=torch.bfloat16, device="cuda")
    ops.dropout_add_norm(x, br, out2, w, 1e-6, xn2, word, site, p)
    assert torch.equal(out1, out2) and torch.equal(xn1, xn2)
    kept = (out1 != x).float().mean().item()
    assert abs(kept - (1.0 - 26.0 / 256.0)) < 0.01                       # 0.1 quantised to 26 / 256 of the branch elements dropped
    dy = _rand((rows, C + 32), torch.bfloat16, 1.0, 54)
    d1, n1 = _rand((rows, C), torch.float32, 1.0, 55), torch.zeros((rows, C + 32), dtype=torch.bfloat16, device="cuda")
    d2, n2 = d1.clone(), torch.zeros_like(n1)
    ops.rmsnorm_bwd(x, w, dy, 1e-6, d1)
    ops.dropout(d1, n1, rows, C, word, site + 3, p)
    ops.rmsnorm_bwd_drop(x, w, dy, 1e-6, d2, n2, word, site + 3, p)
    assert torch.equal(d1, d2) and torch.equal(n1, n2) and n2[:, C:].abs().max().item() == 0


# ------------------------------------------------------------------------------------------- small kernels
def test_patchify_matches_conv_unfold(ops):
    F, S, P = 3, 224, 14
    img = _rand((F, 3, S, S), torch.float32, 1.0, 28)
    out = torch.full((F * 256, 592), 7.0, dtype=torch.float16, device="cuda")
    ops.patchify(img, out, S, P)
    want = torch.nn.functional.unfold(img, kernel_size=P, stride=P).transpose(1, 2).reshape(F * 256, 588)
    assert torch.equal(out[:, :588], want.half())
    assert out[:, 588:].abs().max().item() == 0


def test_patchify_uint8_fused_normalise_is_bit_exact(ops):
    """Raw uint8 frames + fused (x/255 - mean) / std == the processors' host normalisation followed by the fp32 path.
    The reference normalises on the CPU (dataloader workers), where torch divides exactly; torch's CUDA `x / 255.0`
    multiplies by a reciprocal instead and differs in the last bit for half of the byte values."""
    F, S, P = 3, 224, 14
    g = torch.Generator().manual_seed(77)
    u8 = torch.randint(0, 256, (F, 3, S, S), generator=g, dtype=torch.uint8)
    mean = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
    std = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)
    x = (u8.float() / 255.0 - mean) / std                     # ToTensorVideo + NormalizeVideo (blip_processors.py:355-395), CPU
    a = torch.zeros((F * 256, 592), dtype=torch.float16, device="cuda")
    b = torch.ones((F * 256, 592), dtype=torch.float16, device="cuda")
    ops.patchify(x.cuda().contiguous(), a, S, P)
    ops.patchify_u8(u8.cuda(), b, S, P, mean.flatten().tolist(), std.flatten().tolist())
    assert torch.equal(a, b)


def test_gated_gelu_fwd_bwd(ops):
    M, Fd = 333, 5120
    ab = _rand((M, 2 * Fd), torch.bfloat16, 1.0, 29)
    h = torch.empty((M, Fd + 32), dtype=torch.bfloat16, device="cuda")
    ops.gated_gelu_fwd(ab, h, M, Fd)
    a, b = ab[:, :Fd].float().requires_grad_(True), ab[:, Fd:].float().requires_grad_(True)
    want = torch.nn.functional.gelu(a) * b
    _close(h[:, :Fd], want, 8e-3, 8e-3, "gated gelu")
    dh = _rand((M, Fd), torch.bfloat16, 1.0, 30)
    want.backward(dh.float())
    dab = torch.empty_like(ab)
    ops.gated_gelu_bwd(ab, dh, dab, M, Fd)
    _close(dab[:, :Fd], a.grad, 1e-2, 1e-2, "d wi0")
    _close(dab[:, Fd:], b.grad, 1e-2, 1e-2, "d wi1")


def test_gather_scatter_rows(ops):
    C = 2048
    emb = _rand((500, C), torch.float32, 1.0, 31)
    frames = _rand((40, C), torch.float32, 1.0, 32)
    idx = torch.tensor([ops.INT_MIN, 3, -1, -40, 499, -7, 0], dtype=torch.int32, device="cuda")
    out = torch.full((7, C), 9.0, device="cuda")
    ops.gather_rows(idx, emb, frames, out)
    want = torch.stack([torch.zeros(C, device="cuda"), emb[3], frames[0], frames[39], emb[499], frames[6], emb[0]])
    assert torch.equal(out, want)
    d = torch.zeros_like(frames)
    ops.scatter_frames(idx, out, d)
    assert torch.equal(d[0], out[2]) and torch.equal(d[39], out[3]) and torch.equal(d[6], out[5]) and d[1].abs().max() == 0


def test_cross_entropy(ops):
    rows, V = 37, 32128
    logits = _rand((rows, V), torch.float32, 3.0, 33)
    labels = torch.randint(0, V, (rows,), device="cuda")
    labels[5] = -100
    labels[36] = -100
    n_valid = (labels >= 0).sum().item()
    row_loss = torch.zeros(rows, device="cuda")
    dl = torch.zeros((rows, V), dtype=torch.bfloat16, device="cuda")
    loss = torch.zeros(1, device="cuda")
    ops.cross_entropy(logits, labels, row_loss, dl, 1.0 / n_valid, loss_sum=loss)
    lg = logits.clone().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(lg, labels, ignore_index=-100)
    want.backward()
    assert abs(row_loss.sum().item() / n_valid - want.item()) < 1e-4
    assert abs(loss.item() - want.item()) < 1e-4
    _close(dl, lg.grad, 1e-2, 1e-7, "dlogits")
    # gscale < 0: the mean over valid targets is taken on the device (graph-capturable step)
    dl2, loss2 = torch.zeros_like(dl), torch.zeros(1, device="cuda")
    ops.cross_entropy(logits, labels, None, dl2, -1.0, loss_sum=loss2)
    assert abs(loss2.item() - want.item()) < 1e-4 and torch.equal(dl2, dl)


def test_lora_down_and_wgrad(ops):
    M, K, R = 1001, 2048, 24
    x = torch.zeros((M, K + 32), dtype=torch.bfloat16, device="cuda")
    x[:, :K] = _rand((M, K), torch.bfloat16, 1.0, 34)
    x[:, K:] = 5.0
    A = _rand((R, K), torch.float32, 0.02, 35)
    ops.lora_down(x, A, M, K, R)
    want = x[:, :K].float() @ A.t()
    _close(x[:, K:K + R], want, 8e-3, 8e-3, "lora down")
    assert x[:, K + R:].abs().max().item() == 0
    # dB[n, r] = sum_m dy[m, n] xa[m, r]
    N = 1000
    dy = _rand((M, N), torch.bfloat16, 1.0, 36)
    dB = torch.zeros((N, 8), dtype=torch.float32, device="cuda")
    xa = x[:, K + 8:K + 16]
    ops.skinny_wgrad(dy.data_ptr(), dy.stride(0), xa.data_ptr(), x.stride(0), M, N, dB, False, ops.BF16, impl="cc")
    _close(dB, dy.float().t() @ xa.float(), 1e-3, 1e-3, "dB")
    for tr in (False, True):                                  # tensor-core variant, ragged M and C, both output layouts
        o = torch.zeros((8, N) if tr else (N, 8), dtype=torch.float32, device="cuda")
        ops.skinny_wgrad(dy.data_ptr(), dy.stride(0), xa.data_ptr(), x.stride(0), M, N, o, tr, ops.BF16, impl="tc")
        want = dy.float().t() @ xa.float()
        _close(o.t() if tr else o, want, 2e-3, 2e-3, "dB tc")
    # two adjacent slots in one pass (dA of q|k, wi_0|wi_1, cross k|v)
    q16 = _rand((M, 32), torch.bfloat16, 1.0, 43)
    o1, o2 = torch.zeros((8, N), dtype=torch.float32, device="cuda"), torch.zeros((8, N), dtype=torch.float32, device="cuda")
    ops.skinny_wgrad_pair(dy.data_ptr(), dy.stride(0), q16.data_ptr() + 16, q16.stride(0), M, N, o1, o2, True, ops.BF16)
    _close(o1.t(), dy.float().t() @ q16[:, 8:16].float(), 2e-3, 2e-3, "pair slot 0")
    _close(o2.t(), dy.float().t() @ q16[:, 16:24].float(), 2e-3, 2e-3, "pair slot 1")
    # small-M down-projection kernel
    Wd = _rand((32, K), torch.bfloat16, 0.05, 39)
    xs = x[:56]
    outd = torch.zeros((56, 32), dtype=torch.bfloat16, device="cuda")
    ops.down32(xs[:, :K], Wd, outd, 56)
    _close(outd, xs[:, :K].float() @ Wd.float().t(), 1e-2, 1e-2, "small down")
    dA = torch.zeros((8, K), dtype=torch.float32, device="cuda")
    ops.skinny_wgrad(x.data_ptr(), x.stride(0), dy[:, 8:16].contiguous().data_ptr(), 8, M, K, dA, True, ops.BF16, impl="cc")
    _close(dA, dy[:, 8:16].float().t() @ x[:, :K].float(), 1e-3, 1e-3, "dA")
    # decoder-sized M (the no-atomics kernel): accumulates into a non-zero output, ragged C, both layouts
    for Ms, C in ((56, 1000), (13, 2048), (256, 72)):
        P, Q = _rand((Ms, C), torch.bfloat16, 1.0, 40), _rand((Ms, 8), torch.bfloat16, 1.0, 41)
        for tr in (False, True):
            o = torch.ones((8, C) if tr else (C, 8), dtype=torch.float32, device="cuda")
            ops.skinny_wgrad(P.data_ptr(), P.stride(0), Q.data_ptr(), Q.stride(0), Ms, C, o, tr, ops.BF16, impl="cc")
            want = 1.0 + P.float().t() @ Q.float()
            _close(o.t() if tr else o, want, 1e-4, 1e-4, "small-M wgrad")


def test_casts_transpose_colsum(ops):
    x = _rand((300, 776), torch.float32, 1.0, 37)
    out = torch.empty((300, 776), dtype=torch.bfloat16, device="cuda")
    ops.cast_to(x, out)
    assert torch.equal(out, x.bfloat16())
    o2 = torch.zeros((300, 800), dtype=torch.float16, device="cuda")
    ops.cast2d(x, o2, 300, 776)
    assert torch.equal(o2[:, :776], x.half())
    t = torch.empty((776, 304), dtype=torch.bfloat16, device="cuda")
    ops.transpose16(out, t, 300, 776)
    assert torch.equal(t[:, :300], out.t())
    cs = torch.zeros(776, device="cuda")
    ops.colsum(x, cs)
    _close(cs, x.sum(0), 1e-4, 1e-4, "colsum")
    y = _rand((300, 776), torch.float32, 1.0, 38)
    y0 = y.clone()
    ops.axpby(x, y, 0.5, 2.0)
    _close(y, 0.5 * x + 2.0 * y0, 1e-6, 1e-6, "axpby")
    g = torch.zeros((6, 776), device="cuda")
    ops.group_mean(x, g, 6, 50, 776)
    _close(g, x.view(6, 50, 776).mean(1), 1e-5, 1e-5, "group mean")
