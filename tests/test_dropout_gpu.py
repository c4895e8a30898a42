"""Parity tests of the train-mode dropout path (csrc/dropmask.cuh, csrc/dropout.cu, the DROP instantiations of the attention
kernels, T5Engine / QFormerEngine / BLIP2_MR with train_dropout) against the CPU oracle, which evaluates the SAME counter-hash
masks (oracle/dropout.py) -- so every comparison is elementwise, not statistical.

First run on a B200 in round 2 (profiles/r02_call1.md: green); train-mode dropout is the default of BLIP2_MR.train() since.
"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]

from mr_blip_b200.dims import TINY, T5_PREFIX  # noqa: E402

SEED = 0x9E3779B1


@pytest.fixture(scope="module")
def word():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mr_blip_b200 import _lib
    _lib.load()
    return torch.tensor([SEED - (1 << 32)], dtype=torch.int32, device="cuda")      # the kernels read the word as uint32


def _rand(shape, dtype, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


def _mask(site, rows, cols, p):
    """keep * scale as a float32 CUDA tensor [rows, cols] from the oracle's restatement of the mask function."""
    from oracle import dropout as od
    return (torch.from_numpy(od.keep_mask(SEED, site, rows, cols, p)).float() * float(od.scale_of(p))).cuda()


def _relfro(got, want):
    got, want = torch.as_tensor(got).float().cpu(), torch.as_tensor(want).float().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    return ((got - want).norm() / want.norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------ one-pass kernels
@pytest.mark.parametrize("p", [0.1, 0.05, 0.0])
def test_dropout_kernels_match_oracle_mask_exactly(word, p):
    from mr_blip_b200 import ops
    rows, cols, site = 77, 2048, 0x1234
    m = _mask(site, rows, cols, p)
    x = _rand((rows, cols), torch.float32, 1.0, 1)
    out = ops.dropout(x, torch.empty_like(x), rows, cols, word, site, p)
    assert torch.equal(out, x * m)                                               # fp32 -> fp32: bit exact
    for dt in (torch.bfloat16, torch.float16):
        o16 = ops.dropout(x, torch.empty((rows, cols), dtype=dt, device="cuda"), rows, cols, word, site, p)
        assert torch.equal(o16, (x * m).to(dt))                                  # fp32 -> 16 bit (the masked dgrad operand)
        big = torch.zeros((rows, cols + 32), dtype=dt, device="cuda")            # strided, in place (final-norm dropout on x_ext)
        big[:, :cols] = x.to(dt)
        want = (big[:, :cols].float() * m).to(dt)
        ops.dropout(big[:, :cols], big[:, :cols], rows, cols, word, site, p)
        assert torch.equal(big[:, :cols], want) and big[:, cols:].abs().max().item() == 0
    r = _rand((rows, cols), torch.float32, 1.0, 2)
    got = ops.dropout_add(r, x, torch.empty_like(x), word, site, p)
    assert torch.allclose(got, r + x * m, rtol=1e-6, atol=1e-6)                  # the kernel may contract the scale into an FMA
    # another site / another seed word: different masks
    other = ops.dropout(x, torch.empty_like(x), rows, cols, word, site + 1, p)
    if p > 0:
        assert not torch.equal(other, out)
        w2 = torch.tensor([12345], dtype=torch.int32, device="cuda")
        assert not torch.equal(ops.dropout(x, torch.empty_like(x), rows, cols, w2, site, p), out)


def test_gated_gelu_with_inner_dropout(word):
    from mr_blip_b200 import ops
    M, F, site, p = 130, 5120, 77, 0.1
    ab = _rand((M, 2 * F), torch.bfloat16, 1.0, 3)
    m = _mask(site, M, F, p)
    a, b = ab[:, :F].float().requires_grad_(True), ab[:, F:].float().requires_grad_(True)
    want = torch.nn.functional.gelu(a) * b * m
    h = torch.zeros((M, F + 32), dtype=torch.bfloat16, device="cuda")
    ops.gated_gelu_fwd_drop(ab, h, M, F, word, site, p)
    assert _relfro(h[:, :F], want) < 4e-3
    assert torch.equal(h[:, :F] == 0, (want.detach().to(torch.bfloat16) == 0))  # zeros exactly where the mask (or the value) is zero
    dh = _rand((M, F), torch.bfloat16, 1.0, 4)
    want.backward(dh.float())
    dab = torch.zeros((M, 2 * F + 32), dtype=torch.bfloat16, device="cuda")
    ops.gated_gelu_bwd_drop(ab, dh, dab, M, F, word, site, p)
    assert _relfro(dab[:, :F], a.grad) < 6e-3 and _relfro(dab[:, F:2 * F], b.grad) < 6e-3


# ------------------------------------------------------------------------------------------------ LoRA input dropout
@pytest.mark.parametrize("M,K,nlin", [(8148, 2048, 3), (8148, 5120, 1), (56, 2048, 3), (56, 2048, 2), (300, 5120, 1), (2049, 2048, 2)])
def test_lora_dropout_kernels(word, M, K, nlin):
    from mr_blip_b200 import ops
    p, site0 = 0.05, 0x2108
    x_ext = torch.zeros((M, K + 32), dtype=torch.bfloat16, device="cuda")
    x_ext[:, :K] = _rand((M, K), torch.bfloat16, 1.0, 5)
    A = torch.zeros((32, K), dtype=torch.bfloat16, device="cuda")
    A[:8 * nlin] = _rand((8 * nlin, K), torch.bfloat16, 1.0 / math.sqrt(K), 6)
    x = x_ext[:, :K].float()
    masks = [_mask(site0 + j, M, K, p) for j in range(nlin)]
    # forward: u_j = drop_j(x) A_j^T into the extension columns, zeros in the unused ones
    x_ext[:, K:] = 7.0
    ops.lora_down_drop(x_ext[:, :K], A, x_ext[:, K:], M, K, nlin, word, site0, p)
    for j in range(nlin):
        want = (x * masks[j]) @ A[8 * j:8 * j + 8].float().t()
        assert _relfro(x_ext[:, K + 8 * j:K + 8 * j + 8], want) < 5e-3, j
    assert x_ext[:, K + 8 * nlin:].abs().max().item() == 0
    # backward: dA_j += q_j^T drop_j(x)
    q = torch.zeros((M, 32), dtype=torch.bfloat16, device="cuda")
    q[:, :8 * nlin] = _rand((M, 8 * nlin), torch.bfloat16, 1.0, 7)
    for j in range(nlin):
        dA = torch.ones((8, K), dtype=torch.float32, device="cuda")
        ops.lora_wgrad_drop(x_ext.data_ptr(), x_ext.stride(0), q.data_ptr() + 16 * j, q.stride(0), M, K, dA, ops.BF16, word, site0 + j, p)
        want = q[:, 8 * j:8 * j + 8].float().t() @ (x * masks[j])
        assert _relfro(dA - 1.0, want) < 2e-3, j
    # backward: dx += sum_j mask_j * (q_j A_j), 16-bit and fp32 destinations
    want = sum(masks[j] * (q[:, 8 * j:8 * j + 8].float() @ A[8 * j:8 * j + 8].float()) for j in range(nlin))
    base = _rand((M, K), torch.float32, 1.0, 8)
    d32 = base.clone()
    ops.lora_dx_drop(q, A, nlin, d32, M, K, word, site0, p)
    assert _relfro(d32 - base, want) < 1e-3
    d16 = torch.zeros((M, K + 32), dtype=torch.bfloat16, device="cuda")
    d16[:, :K] = base.to(torch.bfloat16)
    ops.lora_dx_drop(q, A, nlin, d16[:, :K], M, K, word, site0, p)
    assert _relfro(d16[:, :K], base.to(torch.bfloat16).float() + want) < 4e-3
    assert d16[:, K:].abs().max().item() == 0


# ------------------------------------------------------------------------------------------------ attention probabilities
def _ref_attention(q, k, v, scale, bias, kmask, causal, mask):
    """fp32 torch reference with the dropout mask on the probabilities; q [B, Lq, H, hd] etc."""
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    B, H, Lq, Lk = s.shape
    if bias is not None:
        i = torch.arange(Lq, device="cuda")[:, None]
        j = torch.arange(Lk, device="cuda")[None, :]
        s = s + bias[:, (j - i) + (Lq - 1)][None]
    if kmask is not None:
        s = s.masked_fill(kmask[:, None, None, :] == 0, float("-inf"))
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Lq, Lk, device="cuda", dtype=torch.bool), 1), float("-inf"))
    pr = torch.softmax(s, -1)
    return torch.matmul(pr * mask.view(B, H, Lq, Lk), vf).permute(0, 2, 1, 3)


@pytest.mark.parametrize("impl,dtype,B,H,Lq,Lk,has_bias,has_mask,causal", [
    ("mma", torch.bfloat16, 2, 4, 72, 72, True, True, False),       # tiny-config encoder shape
    ("mma", torch.bfloat16, 2, 4, 9, 9, True, True, True),          # decoder self-attention
    ("mma", torch.bfloat16, 2, 4, 9, 72, False, True, False),       # decoder cross-attention
    ("mma", torch.float16, 3, 12, 32, 32, False, False, False),     # Q-Former self-attention
    ("mma", torch.float16, 3, 12, 32, 257, False, False, False),    # Q-Former cross-attention (generic kernel in dropout mode)
    ("tc", torch.bfloat16, 2, 4, 300, 300, True, True, False),      # tcgen05 encoder: ragged last tiles
    ("tc", torch.bfloat16, 1, 8, 2037, 2037, True, True, False),    # QVH encoder length
    ("tc", torch.bfloat16, 2, 4, 16, 600, False, True, False),      # decoder cross-attention over a long encoder
    ("auto", torch.bfloat16, 2, 4, 16, 2037, False, True, False),   # the same through the few-query kernels (keys split over a cluster)
    ("auto", torch.bfloat16, 1, 3, 21, 700, False, True, False),    #   two 16-row blocks
    ("mma", torch.float16, 1, 2, 5, 515, False, False, False),
    ("tc", torch.bfloat16, 2, 4, 200, 200, True, False, True),      # causal
])
def test_attention_probability_dropout_fwd_bwd(word, impl, dtype, B, H, Lq, Lk, has_bias, has_mask, causal):
    from mr_blip_b200 import ops
    hd, p, site = 64, 0.1, 0x1041
    scale = 1.0 if dtype == torch.bfloat16 else hd ** -0.5
    q = _rand((B, Lq, H, hd), dtype, 0.5, 11).requires_grad_(True)
    k = _rand((B, Lk, H, hd), dtype, 0.5, 12).requires_grad_(True)
    v = _rand((B, Lk, H, hd), dtype, 1.0, 13).requires_grad_(True)
    bias = _rand((H, Lq + Lk - 1), torch.float32, 1.0, 14) if has_bias else None
    kmask = None
    if has_mask:
        kmask = torch.ones((B, Lk), dtype=torch.int32, device="cuda")
        kmask[-1, Lk - Lk // 6:] = 0
    mask = _mask(site, B * H * Lq, Lk, p)
    want = _ref_attention(q, k, v, scale, bias, kmask, causal, mask)
    dout = _rand((B, Lq, H, hd), dtype, 1.0, 15)
    want.backward(dout.float())
    out = torch.empty((B, Lq, H, hd), dtype=dtype, device="cuda")
    lse = torch.empty((B, H, Lq), dtype=torch.float32, device="cuda")
    rs = H * hd
    drop = (word, site, p)
    ops.attention_fwd(q.detach(), k.detach(), v.detach(), out, B, H, Lq, Lk, hd, scale, (Lq * rs, rs), (Lk * rs, rs), (Lk * rs, rs),
                      (Lq * rs, rs), bias=bias, bias_zero=Lq - 1, kmask=kmask, causal=causal, lse=lse, impl=impl, drop=drop)
    assert _relfro(out, want) < 1e-2
    if impl == "mma" and not (has_bias or has_mask or causal) and Lq <= 32 and Lk > 64:
        # forward-only call without lse: the one-shot Q-Former cross-attention kernel (attn_xq_kernel<T, DROP>)
        o2 = torch.empty_like(out)
        ops.attention_fwd(q.detach(), k.detach(), v.detach(), o2, B, H, Lq, Lk, hd, scale, (Lq * rs, rs), (Lk * rs, rs), (Lk * rs, rs),
                          (Lq * rs, rs), impl=impl, drop=drop)
        assert _relfro(o2, want) < 1e-2
    dq, dk, dv = (torch.empty_like(t) for t in (q, k, v))
    ws = torch.empty((B * H * Lq,), dtype=torch.float32, device="cuda")
    ops.attention_bwd(q.detach(), k.detach(), v.detach(), out, dout, dq, dk, dv, B, H, Lq, Lk, hd, scale, (Lq * rs, rs), (Lk * rs, rs),
                      (Lk * rs, rs), (Lq * rs, rs), (Lq * rs, rs), lse, ws, bias=bias, bias_zero=Lq - 1, kmask=kmask, causal=causal,
                      impl=impl, drop=drop)
    assert _relfro(dv, v.grad) < 1.5e-2
    assert _relfro(dq, q.grad) < 2e-2
    assert _relfro(dk, k.grad) < 2e-2
    # the mask, not just its rate: the eval-mode kernel differs by far more than the tolerance
    ev = torch.empty_like(out)
    ops.attention_fwd(q.detach(), k.detach(), v.detach(), ev, B, H, Lq, Lk, hd, scale, (Lq * rs, rs), (Lk * rs, rs), (Lk * rs, rs),
                      (Lq * rs, rs), bias=bias, bias_zero=Lq - 1, kmask=kmask, causal=causal, impl=impl)
    assert _relfro(ev, want) > 5e-2


# ------------------------------------------------------------------------------------------------ engines and model
def _t5_inputs(d):
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 72, d.d_model, generator=g) * 2.0
    mask = torch.ones(2, 72, dtype=torch.long)
    mask[1, 60:] = 0
    labels = torch.randint(2, 1000, (2, 9), generator=g)
    labels[:, -1] = 1
    labels[1, 6:] = -100
    labels[1, 5] = 1
    return emb, mask, labels


@pytest.fixture(scope="module")
def model(tiny_sd):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mr_blip_b200.blip2_mr import BLIP2_MR
    return BLIP2_MR(dims=TINY, state_dict=tiny_sd, train_dropout=True, cuda_graphs=False).cuda()


def test_t5_engine_train_mode_vs_oracle(model, tiny_sd):
    """T5 loss, logits, d inputs_embeds and every LoRA gradient with all dropout sites on (0.1 / LoRA 0.05), same masks in the
    oracle.  Tolerances as tests/test_model_gpu.py (bf16 operands vs fp32)."""
    from oracle import t5 as ot5
    from oracle.dropout import Dropper
    _, _, t5 = model.engines()
    emb, mask, labels = _t5_inputs(TINY)
    t5.drop = model.drop_state
    seed = model.drop_state.set_seed(0xC0FFEE11)
    try:
        t5.zero_grads()
        out = t5.loss(emb.cuda(), mask, labels, (labels != -100).long(), backward=True, want_logits=True)
    finally:
        t5.drop = None
    sd = dict(tiny_sd)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in sd if "lora_" in k}
    sd.update(leaves)
    e = emb.clone().requires_grad_(True)
    o = ot5.t5_forward(sd, TINY, e, mask, labels, (labels != -100).long(), drop=Dropper(seed))
    o["loss"].backward()
    assert abs(out["loss"].item() - o["loss"].item()) < 5e-3
    assert _relfro(out["logits"], o["logits"]) < 2e-2
    assert _relfro(out["d_inputs_embeds"], e.grad) < 4e-2
    grads = {id(p): g for p, g in t5.param_grads()}
    for k, leaf in leaves.items():
        assert _relfro(grads[id(model._get(k))], leaf.grad) < 4e-2, k
    # eval-mode oracle is far away: the comparison above can fail
    with torch.no_grad():
        ev = ot5.t5_forward(dict(tiny_sd), TINY, emb, mask, labels, (labels != -100).long())
    # (the two losses sit within 1e-2 of each other on these random weights -- both ~ ln(vocab) -- so the distance is taken on the logits)
    assert _relfro(ev["logits"], o["logits"]) > 0.3


@pytest.mark.parametrize("agg", [None, "mean"])
def test_model_train_step_with_dropout_vs_oracle(model, tiny_sd, agg):
    """BLIP2_MR.forward in train() with train_dropout: Q-Former (frozen, train mode) + T5 + LoRA dropout, loss and all
    trainable gradients against the oracle with the same seed word; eval() is unaffected; a second step draws other masks."""
    from oracle import blip2_mr as ob, synth
    from oracle.dropout import Dropper
    samples = synth.make_samples(batch=2, frames=3, seed=3)
    model.frame_token_aggregation = agg
    model.train()
    for q in model.parameters():
        q.grad = None
    res = model.forward_mr(samples, want_logits=True)
    res["loss"].backward()
    seed = model.drop_state.seed
    sd = dict(tiny_sd)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in sd if "lora_" in k or k.startswith("t5_proj.")}
    sd.update(leaves)
    o = ob.forward_mr(sd, TINY, model.t5_tokenizer, samples, frame_token_aggregation=agg, drop=Dropper(seed))
    o["loss"].backward()
    assert _relfro(res["qformer"], o["qformer"]) < 2e-3
    assert abs(res["loss"].item() - o["loss"].item()) < 5e-3
    assert _relfro(res["logits"], o["logits"]) < 2e-2
    for k, leaf in leaves.items():
        assert _relfro(model._get(k).grad, leaf.grad) < 4e-2, k
    l2 = model.forward_mr(samples)["loss"].item()
    assert model.drop_state.seed != seed and abs(l2 - res["loss"].item()) > 1e-4
    model.eval()
    with torch.no_grad():
        ev = model.forward_mr(samples, want_logits=True)
        oe = ob.forward_mr(dict(tiny_sd), TINY, model.t5_tokenizer, samples, frame_token_aggregation=agg)
    assert abs(ev["loss"].item() - oe["loss"].item()) < 5e-3
    model.frame_token_aggregation = None


def test_graphed_step_draws_fresh_masks_and_matches_eager(tiny_sd):
    """With CUDA graphs the seed word is rewritten before every replay: replays differ from each other and each equals the
    eager step run with the same seed word."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from oracle import synth
    samples = synth.make_samples(batch=2, frames=3, seed=5)
    g = BLIP2_MR(dims=TINY, state_dict=tiny_sd, train_dropout=True, cuda_graphs=True).cuda().train()
    e = BLIP2_MR(dims=TINY, state_dict=tiny_sd, train_dropout=True, cuda_graphs=False, graph_bucket=None).cuda().train()
    losses = []
    for _ in range(4):                                       # eager, capture + replay, replay, replay
        losses.append(g(samples)["loss"].item())
    assert len({round(x, 5) for x in losses}) == 4
    # the second model advances its own state identically (same base seed, same step count) and takes the same bucketed path
    # (the mask rows follow the padded shapes) but never captures: its seen-count is reset before every step
    el = []
    e.graph_bucket = g.graph_bucket
    e.cuda_graphs = True
    for _ in range(4):
        e._seen.clear()
        e._steps.clear()
        el.append(e(samples)["loss"].item())
    for a, b in zip(losses, el):
        assert abs(a - b) < 2e-5 * max(1.0, abs(a)), (losses, el)
