"""Torch-on-CPU stand-ins for the C-ABI ops that T5Engine / QFormerEngine call (mr_blip_b200/ops.py), same signatures and buffer
conventions (extended 16-bit buffers, raw-pointer operands of the weight-gradient ops, strided attention operands), so that the
HOST logic of the engines -- which op runs on which buffer, the hand-written backward's chain rule, the dropout sites and the
LoRA-dropout decomposition -- can be checked against the oracle in the CPU suite, where no kernel can run.  Test infrastructure
only: nothing here is a product path, and kernel arithmetic itself is only tested on the GPU (tests/test_*_gpu.py).

load_engine_module(name) re-compiles an engine module (t5 / vision) with its "cuda" device strings pointing at the CPU and its
`ops` bound to this module.
"""
import contextlib
import ctypes
import importlib.util
import os

import numpy as np
import torch
import torch.nn.functional as F

from oracle import dropout as od

F16, BF16, F32 = 0, 1, 2
_TDT = {F16: torch.float16, BF16: torch.bfloat16, F32: torch.float32}
SPLITK = False
USE_TC_ATTENTION = False          # VitEngine then sends all 257 query rows through one attention_fwd call
GEMM_PROFILE = None
INT_MIN = -(1 << 31)


def splitk_register(side_stream=None):
    pass


def gemm_sm_limit(sms):
    pass


@contextlib.contextmanager
def phase(name):
    yield


def _from_ptr(ptr, rows, cols, ld, dtype):
    """[rows, cols] view (row stride ld elements) of the memory at a raw address (what the C ABI receives)."""
    es = torch.empty((), dtype=dtype).element_size()
    n = ((rows - 1) * ld + cols) * es
    buf = (ctypes.c_char * n).from_address(ptr)
    return torch.frombuffer(buf, dtype=dtype).as_strided((rows, cols), (ld, 1))


def _seed(word):
    return int(word.view(-1)[0].item()) & 0xFFFFFFFF


def _mask(word, site, rows, cols, p):
    if p <= 0.0:
        return torch.ones((rows, cols))
    return torch.from_numpy(od.keep_mask(_seed(word), site, rows, cols, p)).float() * float(od.scale_of(p))


# ---------------------------------------------------------------------------------------------- GEMM family
def gemm(a, b, out=None, bias=None, gelu=False, resid=None, out_dtype=None, row_group=0, out_rows=None, force_bn=0, M=None, K=None):
    M = a.shape[0] if M is None else M
    K = a.shape[1] if K is None else K
    N = b.shape[0]
    y = a[:M, :K].float() @ b[:, :K].float().t()
    if bias is not None:
        y = y + bias
    if gelu:
        y = F.gelu(y)
    if row_group > 0:                                        # patch embed: token 0 of every frame is the cls row, written elsewhere
        m = torch.arange(M)
        orow = (m // row_group) * (row_group + 1) + 1 + m % row_group
        if resid is not None:
            y = y + resid[1 + m % row_group, :N]
        out[orow, :N] = y.to(out.dtype)
        return out
    if resid is not None:
        y = y + resid[:M, :N]
    if out is None:
        out = torch.empty((out_rows or M, N), dtype=out_dtype or a.dtype)
    out[:M, :N] = y.to(out.dtype)
    return out


def patchify(img, out, img_size, patch):
    Fr, g = img.shape[0], img_size // patch
    x = img.view(Fr, 3, g, patch, g, patch).permute(0, 2, 4, 1, 3, 5).reshape(Fr * g * g, 3 * patch * patch)
    out.zero_()
    out[:, :3 * patch * patch] = x.to(out.dtype)


def cls_pos(cls, pos, x, frames, tokens, C):
    x.view(frames, tokens, C)[:, 0] = cls + pos[0]


def scatter_frames(idx, dout, dframes):
    idx = idx.long()
    neg = (idx < 0) & (idx != INT_MIN)
    dframes.index_add_(0, -(idx[neg] + 1), dout[neg])


def group_mean(x, out, groups, n, C):
    out.copy_(x.view(groups, n, C).mean(1))


def group_mean_bwd(dout, dx, groups, n, C):
    dx.view(groups, n, C).copy_((dout / n)[:, None, :].expand(groups, n, C))


def colsum(x, out):
    out += x.sum(0)


def axpby(x, y, a, b):
    y.copy_(a * x + b * y)


def lora_pack(table, n, blocks_per_linear):
    """LoraPackDesc records (csrc/elementwise.cu): A [8,K] / B [N,8] fp32 -> their four bf16 operand slots."""
    for rec in table.tolist():
        a_p, b_p, ext_p, ext_ld, bd_p, bd_ld, ad_p, ad_ld, eb_p, eb_ld, K, N, sbits = rec
        scale = float(np.int64(sbits).view(np.float64))
        A = _from_ptr(a_p, 8, K, K, torch.float32)
        B = _from_ptr(b_p, N, 8, 8, torch.float32) * scale
        _from_ptr(ext_p, N, 8, ext_ld, torch.bfloat16).copy_(B.to(torch.bfloat16))
        _from_ptr(bd_p, 8, N, bd_ld, torch.bfloat16).copy_(B.t().to(torch.bfloat16))
        _from_ptr(ad_p, 8, K, ad_ld, torch.bfloat16).copy_(A.to(torch.bfloat16))
        _from_ptr(eb_p, K, 8, eb_ld, torch.bfloat16).copy_(A.t().to(torch.bfloat16))


def down32(x, W, out, M):
    out[:M, :32] = (x[:M].float() @ W.float().t()).to(out.dtype)
    return out


def transpose16(x, out, rows, cols):
    out[:cols, :rows] = x[:rows, :cols].t()


def cast2d(x, out, rows, cols):
    out[:rows, :cols] = x[:rows, :cols].to(out.dtype)


def cast_to(x, out):
    out.copy_(x.to(out.dtype))
    return out


def skinny_wgrad(P, ldp, Q, ldq, M, C, out, transposed_out, dtype, impl="auto"):
    p = _from_ptr(P, M, C, ldp, _TDT[dtype]).float()
    q = _from_ptr(Q, M, 8, ldq, _TDT[dtype]).float()
    r = p.t() @ q                                            # [C, 8]
    out += r.t() if transposed_out else r


def skinny_wgrad_pair(P, ldp, Q, ldq, M, C, out, out2, transposed_out, dtype):
    es = 2
    skinny_wgrad(P, ldp, Q, ldq, M, C, out, transposed_out, dtype)
    skinny_wgrad(P, ldp, Q + 8 * es, ldq, M, C, out2, transposed_out, dtype)


# ---------------------------------------------------------------------------------------------- norms, activations, loss
def norm(x, w, bias, eps, mode, add=None, out_f32=None, out_h=None, sum_out=None, ld_h=None):
    rows, C = x.shape
    v = x if add is None else x + add
    if mode == 1:
        y = w * (v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps))
    else:
        y = F.layer_norm(v, (C,), w, bias, eps)
    if out_f32 is not None:
        out_f32.copy_(y)
    if out_h is not None:
        out_h[:, :C] = y.to(out_h.dtype)


def rmsnorm_bwd(x, w, dy, eps, dres, lora_A=None, R=0):
    assert lora_A is None
    rows, C = x.shape
    xx = x.detach().clone().requires_grad_(True)
    y = w * (xx * torch.rsqrt(xx.pow(2).mean(-1, keepdim=True) + eps))
    y.backward(dy[:rows, :C].float())
    dres += xx.grad


def _gg(ab, F_):
    return ab[:, :F_].float(), ab[:, F_:2 * F_].float()


def gated_gelu_fwd(ab, h, M, F_):
    a, b = _gg(ab[:M], F_)
    h[:M, :F_] = (F.gelu(a) * b).to(h.dtype)


def gated_gelu_fwd_drop(ab, h, M, F_, seed, site, p):
    a, b = _gg(ab[:M], F_)
    h[:M, :F_] = (F.gelu(a) * b * _mask(seed, site, M, F_, p)).to(h.dtype)


def _gelu_grad(a):
    return 0.5 * (1.0 + torch.erf(a * 0.7071067811865476)) + a * torch.exp(-0.5 * a * a) * 0.3989422804014327


def gated_gelu_bwd(ab, dh, dab, M, F_, mask=None):
    a, b = _gg(ab[:M], F_)
    d = dh[:M, :F_].float()
    if mask is not None:
        d = d * mask
    dab[:M, :F_] = (d * b * _gelu_grad(a)).to(dab.dtype)
    dab[:M, F_:2 * F_] = (d * F.gelu(a)).to(dab.dtype)


def gated_gelu_bwd_drop(ab, dh, dab, M, F_, seed, site, p):
    gated_gelu_bwd(ab, dh, dab, M, F_, _mask(seed, site, M, F_, p))


def gather_rows(idx, emb, frames, out):
    idx = idx.long()
    rows = torch.zeros_like(out)
    pos = idx >= 0
    rows[pos] = emb[idx[pos]].float()
    neg = (idx < 0) & (idx != INT_MIN)
    if neg.any():
        rows[neg] = frames[-(idx[neg] + 1)]
    out.copy_(rows)


def cross_entropy(logits, labels, row_loss=None, dlogits=None, gscale=1.0, loss_sum=None):
    rows, V = logits.shape
    valid = labels >= 0
    gs = 1.0 / max(int(valid.sum()), 1) if gscale < 0 else gscale
    lp = torch.log_softmax(logits, -1)
    safe = labels.clamp_min(0)
    nll = -lp.gather(1, safe[:, None])[:, 0] * valid
    if row_loss is not None:
        row_loss.copy_(nll)
    if loss_sum is not None:
        loss_sum += (nll * gs).sum()
    if dlogits is not None:
        g = lp.exp()
        g[torch.arange(rows), safe] -= 1.0
        dlogits[:, :V] = (g * (valid[:, None] * gs)).to(dlogits.dtype)


# ---------------------------------------------------------------------------------------------- attention
def _view4(t, B, L, H, hd, strides):
    return torch.as_strided(t, (B, L, H, hd), (strides[0], strides[1], hd, 1), t.storage_offset())


def _scores(q, k, scale, bias, bias_zero, kmask, causal, q_pos0):
    s = torch.matmul(q.permute(0, 2, 1, 3), k.permute(0, 2, 3, 1)) * scale          # [B, H, Lq, Lk]
    Lq, Lk = s.shape[-2:]
    if bias is not None:
        i = torch.arange(Lq)[:, None] + q_pos0
        j = torch.arange(Lk)[None, :]
        s = s + bias[:, (j - i) + bias_zero][None]
    if kmask is not None:
        s = s.masked_fill(kmask[:, None, None, :] == 0, float("-inf"))
    if causal:
        i = torch.arange(Lq)[:, None] + q_pos0
        j = torch.arange(Lk)[None, :]
        s = s.masked_fill(j > i, float("-inf"))
    return s


def _attn(q, k, v, scale, bias, bias_zero, kmask, causal, q_pos0, drop):
    s = _scores(q, k, scale, bias, bias_zero, kmask, causal, q_pos0)
    pr = torch.softmax(s, -1)
    B, H, Lq, Lk = s.shape
    if drop is not None:
        pr = pr * _mask(drop[0], drop[1], B * H * Lq, Lk, drop[2]).view(B, H, Lq, Lk)
    return torch.matmul(pr, v.permute(0, 2, 1, 3)).permute(0, 2, 1, 3), torch.logsumexp(s, -1)


def attention_fwd(q, k, v, out, B, H, Lq, Lk, hd, scale, q_strides, k_strides, v_strides, o_strides, bias=None, bias_zero=0,
                  kmask=None, causal=False, q_pos0=0, lse=None, kv_div=1, impl="auto", drop=None):
    assert kv_div == 1
    o, l = _attn(_view4(q, B, Lq, H, hd, q_strides).float(), _view4(k, B, Lk, H, hd, k_strides).float(),
                 _view4(v, B, Lk, H, hd, v_strides).float(), scale, bias, bias_zero, kmask, causal, q_pos0, drop)
    _view4(out, B, Lq, H, hd, o_strides).copy_(o.to(out.dtype))
    if lse is not None:
        lse.copy_(l)
    return out


USE_VIT_ATTENTION = True


def attention_vit_ok(L, hd):
    return USE_VIT_ATTENTION and L == 257 and 64 < hd <= 96 and hd % 8 == 0


def attention_vit(q, k, v, out, frames, H, L, hd, scale, q_strides, k_strides, v_strides, o_strides):
    return attention_fwd(q, k, v, out, frames, H, L, L, hd, scale, q_strides, k_strides, v_strides, o_strides)


def attention_bwd(q, k, v, o, dout, dq, dk, dv, B, H, Lq, Lk, hd, scale, q_strides, k_strides, v_strides, o_strides, do_strides,
                  lse, delta_ws, bias=None, bias_zero=0, kmask=None, causal=False, q_pos0=0, impl="auto", drop=None):
    qf = _view4(q, B, Lq, H, hd, q_strides).float().clone().requires_grad_(True)
    kf = _view4(k, B, Lk, H, hd, k_strides).float().clone().requires_grad_(True)
    vf = _view4(v, B, Lk, H, hd, v_strides).float().clone().requires_grad_(True)
    out, _ = _attn(qf, kf, vf, scale, bias, bias_zero, kmask, causal, q_pos0, drop)
    out.backward(_view4(dout, B, Lq, H, hd, do_strides).float())
    _view4(dq, B, Lq, H, hd, q_strides).copy_(qf.grad.to(dq.dtype))
    _view4(dk, B, Lk, H, hd, k_strides).copy_(kf.grad.to(dk.dtype))
    _view4(dv, B, Lk, H, hd, v_strides).copy_(vf.grad.to(dv.dtype))


# ---------------------------------------------------------------------------------------------- train-mode dropout
def dropout(x, out, rows, cols, seed, site, p):
    out[:rows, :cols] = (x[:rows, :cols].float() * _mask(seed, site, rows, cols, p)).to(out.dtype)
    return out


def dropout_add(resid, branch, out, seed, site, p):
    rows, cols = branch.shape
    out.copy_(resid + branch * _mask(seed, site, rows, cols, p))
    return out


def dropout_add_norm(resid, branch, out, w, eps, xn, seed, site, p):
    dropout_add(resid, branch, out, seed, site, p)
    norm(out, w, None, eps, 1, out_h=xn)
    return out


def rmsnorm_bwd_drop(x, w, dy, eps, dres, dy_next, seed, site, p):
    rmsnorm_bwd(x, w, dy, eps, dres)
    rows, C = x.shape
    dropout(dres, dy_next, rows, C, seed, site, p)


def lora_down_drop(x, A_down, out, M, K, nlin, seed, site0, p):
    out[:M, :32] = 0
    for j in range(nlin):
        xj = x[:M, :K].float() * _mask(seed, site0 + j, M, K, p)
        out[:M, 8 * j:8 * j + 8] = (xj @ A_down[8 * j:8 * j + 8, :K].float().t()).to(out.dtype)
    return out


def lora_wgrad_drop(x, ldx, q, ldq, M, K, dA, dtype, seed, site, p):
    xx = _from_ptr(x, M, K, ldx, _TDT[dtype]).float() * _mask(seed, site, M, K, p)
    qq = _from_ptr(q, M, 8, ldq, _TDT[dtype]).float()
    dA += qq.t() @ xx


def lora_dx_drop(q, A_down, nlin, dx, M, K, seed, site0, p):
    add = torch.zeros((M, K))
    for j in range(nlin):
        add += _mask(seed, site0 + j, M, K, p) * (q[:M, 8 * j:8 * j + 8].float() @ A_down[8 * j:8 * j + 8, :K].float())
    dx[:M, :K] = (dx[:M, :K].float() + add).to(dx.dtype)
    return dx


# ---------------------------------------------------------------------------------------------- engine modules on the CPU
class _NoStream:
    cuda_stream = 0

    def wait_stream(self, other):
        pass


def load_engine_module(name, ops_module=None):
    """mr_blip_b200/<name>.py compiled with its device strings pointing at the CPU and `ops` bound to this module (or to
    `ops_module`, e.g. the product's own ops.py over a HostCAbi)."""
    import mr_blip_b200
    import sys
    path = os.path.join(os.path.dirname(mr_blip_b200.__file__), name + ".py")
    src = open(path).read()
    for a, b in (('device="cuda"', 'device="cpu"'), ('.to("cuda"', '.to("cpu"'), ('.to(device="cuda"', '.to(device="cpu"'),
                 ("assert self.emb.is_cuda", "pass"), ("torch.cuda.Stream()", "_NoStream()"),
                 ("torch.cuda.is_available()", "True"), ('self.device.type != "cuda"', "False"), (".pin_memory()", ""),
                 ('@registry.register_model("blip2_mr")', ""), ('@registry.register_model("blip2_t5")', ""),
                 ("DropState(base_seed=", 'DropState(device="cpu", base_seed=')):
        src = src.replace(a, b)
    assert 'device="cuda"' not in src and '.to("cuda"' not in src, [l for l in src.splitlines() if '"cuda"' in l]
    spec = importlib.util.spec_from_loader("mr_blip_b200._%s_cpu" % name, loader=None)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "mr_blip_b200"
    mod.__dict__["_NoStream"] = _NoStream
    exec(compile(src, path + " (cpu emulation)", "exec"), mod.__dict__)
    mod.ops = ops_module if ops_module is not None else sys.modules[__name__]
    return mod


class HostCAbi:
    """Stands in for mr_blip_b200._lib.call on the CPU: every C-ABI entry point whose kernels are not tcgen05 / TMA code is
    served by the KERNEL SOURCE itself, compiled over tests/cuda_host_shim (csrc/elementwise.cu, dropout.cu, attention.cu); the
    tcgen05 entry points are mapped onto their same-contract siblings (attention_*_tc -> the mma.sync kernels, skinny_wgrad_tc ->
    the CUDA-core kernel) and the GEMM onto a torch matmul over the raw pointers.  With the product's ops.py on top this runs
    wrapper -> ctypes signature -> C entry point -> kernel source end to end without a GPU."""
    ALIAS = {"mrb_attention_fwd_tc": "mrb_attention_fwd", "mrb_attention_bwd_tc": "mrb_attention_bwd",
             "mrb_attention_fwd_tc_drop": "mrb_attention_fwd_drop", "mrb_attention_bwd_tc_drop": "mrb_attention_bwd_drop",
             "mrb_skinny_wgrad_tc": "mrb_skinny_wgrad"}

    def __init__(self, libs):
        from mr_blip_b200 import _lib
        self.libs, self.sigs, self.calls = libs, _lib.SIGNATURES, {}

    def _gemm(self, A, lda, B, ldb, M, N, K, dtype, bias, gelu, resid, ldr, out, out_dtype, ldc, row_group, force_bn, stream):
        y = _from_ptr(A, M, K, lda, _TDT[dtype]).float() @ _from_ptr(B, N, K, ldb, _TDT[dtype]).float().t()
        if bias:
            y = y + _from_ptr(bias, 1, N, N, torch.float32)
        if gelu:
            y = F.gelu(y)
        if row_group > 0:            # patch embed: out_row = (m / G) * (G + 1) + 1 + m % G, resid_row = 1 + m % G
            m = torch.arange(M)
            orow = (m // row_group) * (row_group + 1) + 1 + m % row_group
            if resid:
                y = y + _from_ptr(resid, row_group + 1, N, ldr, torch.float32)[1 + m % row_group]
            o = _from_ptr(out, (M // row_group) * (row_group + 1), N, ldc, _TDT[out_dtype])
            o[orow] = y.to(o.dtype)
            return
        if resid:
            y = y + _from_ptr(resid, M, N, ldr, torch.float32)
        o = _from_ptr(out, M, N, ldc, _TDT[out_dtype])
        o.copy_(y.to(o.dtype))

    def call(self, name, *args):
        self.calls[name] = self.calls.get(name, 0) + 1
        if name == "mrb_gemm":
            return self._gemm(*args)
        if name == "mrb_attention_vit":                      # tcgen05 kernel -> the same-contract mma.sync kernel, all 257 rows
            (q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, frames, H, L, hd, dt, scale, st) = args
            return self.call("mrb_attention_fwd", q, q_bs, q_rs, k, k_bs, k_rs, v, v_bs, v_rs, o, o_bs, o_rs, frames, H, L, L, hd, dt,
                             scale, None, 0, 0, None, 1, 0, 0, None, st)
        if name == "mrb_gemm_sm_limit":                      # scheduling only
            return None
        if name == "mrb_gemm_splitk":                        # same contract + (workspace, bytes, max splits) before the stream
            return self._gemm(*args[:17], args[-1])
        if name == "mrb_skinny_wgrad_tc2":
            P, ldp, Q, ldq, M, C, out, out2, tr, dt, st = args
            self.call("mrb_skinny_wgrad", P, ldp, Q, ldq, M, C, out, tr, dt, st)
            return self.call("mrb_skinny_wgrad", P, ldp, Q + 16, ldq, M, C, out2, tr, dt, st)
        sym = self.ALIAS.get(name, name)
        fn = next((getattr(l, sym) for l in self.libs if hasattr(l, sym)), None)
        assert fn is not None, "no host build of " + sym
        fn.argtypes, fn.restype = self.sigs[name], ctypes.c_int
        rc = fn(*args)
        assert rc == 0, (name, rc)


def load_model_module(ops_module=None):
    """mr_blip_b200/blip2_mr.py on the CPU: its engines are the CPU-compiled engine modules, its ops this module (or
    `ops_module`).  Build the model with cuda_graphs=False (graph capture is CUDA-only)."""
    vision, t5 = load_engine_module("vision", ops_module), load_engine_module("t5", ops_module)
    mod = load_engine_module("blip2_mr", ops_module)
    mod.VitEngine, mod.QFormerEngine, mod.T5Engine = vision.VitEngine, vision.QFormerEngine, t5.T5Engine
    return mod


def load_blip2_t5_module():
    """mr_blip_b200/blip2_t5.py (the thin `blip2_t5` sibling) on the CPU, same arrangement."""
    vision, t5 = load_engine_module("vision"), load_engine_module("t5")
    mod = load_engine_module("blip2_t5")
    mod.VitEngine, mod.QFormerEngine, mod.T5Engine = vision.VitEngine, vision.QFormerEngine, t5.T5Engine
    return mod
