"""Torch-tensor front end of the C ABI: checks devices/dtypes/contiguity, passes raw device pointers
and the current CUDA stream.  PyTorch here is plumbing only (allocation, streams); all arithmetic on
the hot path happens inside libmrblip_b200.so."""
import os

import torch

from . import _lib

F16, BF16, F32 = 0, 1, 2
_DT = {torch.float16: F16, torch.bfloat16: BF16, torch.float32: F32}
INT_MIN = -2 ** 31
USE_TC_ATTENTION = True
USE_FQ_ATTENTION = os.environ.get("MRB_ATTN_FQ", "0") == "1"       # 1: decoder cross-attention through the few-query kernels (measured slower, off)
USE_VIT_ATTENTION = os.environ.get("MRB_ATTN_VIT", "1") != "0"     # 0: the generic flash kernel + the single-row kernel (round 1 path)
GEMM_PROFILE = None      # set to a list to record (M, N, K, cuda start event, cuda end event) per GEMM launch


# Split-K for the decoder-sized (M <= 128) and 32-column GEMMs (csrc/gemm.cu mrb_gemm_splitk).  On by default since round 2
# (tests/test_splitk_gpu.py green on a B200; 19.1 -> 10.9 us at M 64 N 2048 K 2080, 54.6 -> 17.6 us at K 10272,
# 14.9 -> 12.0 us for the LoRA down-projection, step 261.2 -> 257.3 ms: profiles/r02_call1.md).  MRB_GEMM_SPLITK=0 switches it off.
SPLITK = os.environ.get("MRB_GEMM_SPLITK", "1") != "0"
SPLITK_MAX = int(os.environ.get("MRB_GEMM_SPLITK_MAX", "8"))
SPLITK_WS_BYTES = 48 << 20          # 8 splits x 128 rows x 10240 columns of fp32, rounded up
_SPLITK_MAIN = None
_SPLITK_SIDE = {}                   # cuda stream handle -> workspace of that side stream


def splitk_register(side_stream=None):
    """Allocate the split-K workspaces OUTSIDE any graph capture: one for the main chain (whatever stream it runs on; a
    process issues it from one stream at a time) and one per registered side stream, since a workspace is busy until the
    GEMM that uses it has finished on its stream."""
    global _SPLITK_MAIN
    if not SPLITK:
        return
    if _SPLITK_MAIN is None:
        _SPLITK_MAIN = torch.empty(SPLITK_WS_BYTES // 4, dtype=torch.float32, device="cuda")
    if side_stream is not None and side_stream.cuda_stream not in _SPLITK_SIDE:
        _SPLITK_SIDE[side_stream.cuda_stream] = torch.empty(SPLITK_WS_BYTES // 4, dtype=torch.float32, device="cuda")


NVTX = os.environ.get("MRB_NVTX", "0") == "1"     # MRB_NVTX=1: NVTX ranges around the phases of a step (nsys / ncu --nvtx)


class phase:
    """`with ops.phase("vit"):` -- an NVTX range when MRB_NVTX=1, nothing otherwise (the reference has no tracing hooks,
    SURVEY.md §5; ranges are host-side, so under graph replay they bracket the capture / eager launches only)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push("mrb:" + self.name)

    def __exit__(self, *a):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def gemm_sm_limit(sms):
    """The large GEMMs this thread launches from now on occupy at most `sms` SMs (even; 0 = no cap): csrc/gemm.cu mrb_gemm_sm_limit."""
    _lib.call("mrb_gemm_sm_limit", int(sms))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _check(t, *dtypes):
    assert t.is_cuda, "hot-path tensors must live on the GPU (no CPU fallback)"
    if dtypes:
        assert t.dtype in dtypes, (t.dtype, dtypes)
    return t


def gemm(a, b, out=None, bias=None, gelu=False, resid=None, out_dtype=None, row_group=0, out_rows=None, force_bn=0,
         M=None, K=None):
    """out = epi(a[M,K] @ b[N,K]^T).  a/b fp16 or bf16 2-D (row stride >= K, unit column stride)."""
    _check(a, torch.float16, torch.bfloat16)
    _check(b, a.dtype)
    assert a.stride(1) == 1 and b.stride(1) == 1
    M = a.shape[0] if M is None else M
    K = a.shape[1] if K is None else K
    N = b.shape[0]
    assert b.shape[1] >= K or b.shape[1] == K, (a.shape, b.shape)
    if out is None:
        out_dtype = out_dtype or a.dtype
        out = torch.empty((out_rows or M, N), dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1
    if bias is not None:
        _check(bias, torch.float32)
    if resid is not None:
        _check(resid, torch.float32)
        assert resid.stride(1) == 1
    if GEMM_PROFILE is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    if SPLITK and row_group == 0 and (M <= 128 or N <= 32):
        h = _stream()
        if _SPLITK_MAIN is None:
            splitk_register()
        ws = _SPLITK_SIDE.get(h, _SPLITK_MAIN)
        _lib.call("mrb_gemm_splitk", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), M, N, K, _DT[a.dtype], _ptr(bias),
                  int(gelu), _ptr(resid), resid.stride(0) if resid is not None else 0, out.data_ptr(), _DT[out.dtype],
                  out.stride(0), row_group, force_bn, ws.data_ptr(), ws.numel() * 4, SPLITK_MAX, h)
    else:
        _lib.call("mrb_gemm", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), M, N, K, _DT[a.dtype], _ptr(bias),
                  int(gelu), _ptr(resid), resid.stride(0) if resid is not None else 0, out.data_ptr(), _DT[out.dtype],
                  out.stride(0), row_group, force_bn, _stream())
    if GEMM_PROFILE is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        GEMM_PROFILE.append((M, N, K, ev0, ev1))
    return out


def few_queries(Lq, Lk, hd):
    """<= 32 query rows against >= 512 keys, hd 64 (the T5 decoder's cross-attention over the encoder output): the mma.sync entry
    points split the KEYS over a thread-block cluster for this shape (csrc/attention.cu attn_fq_*_kernel; MRB_ATTN_FQ=0 disables),
    the tcgen05 kernels would run one CTA per (clip, head) with 16 of their 256 rows in use."""
    return USE_FQ_ATTENTION and hd == 64 and Lq <= 32 and Lk >= 512


def attention_fwd(q, k, v, out, B, H, Lq, Lk, hd, scale, q_strides, k_strides, v_strides, o_strides, bias=None,
                  bias_zero=0, kmask=None, causal=False, q_pos0=0, lse=None, kv_div=1, impl="auto", drop=None):
    """q/k/v/out: tensors whose data_ptr is row 0 / head 0; *_strides = (batch stride, row stride) in elements.
    impl: "tc" = tcgen05 kernel, "mma" = mma.sync kernel, "auto" = tcgen05 for >= 128 query rows and hd 64 / 72..96.
    drop: None, or (seed word tensor, site, p) = train-mode dropout of the attention probabilities."""
    _check(q, torch.float16, torch.bfloat16)
    if kmask is not None:
        _check(kmask, torch.int32)
    tc_ok = (hd == 64 or (64 < hd <= 96 and hd % 8 == 0))
    use_tc = impl == "tc" or (impl == "auto" and USE_TC_ATTENTION and tc_ok and (Lq >= 128 or Lk >= 512) and not few_queries(Lq, Lk, hd))
    args = (q.data_ptr(), q_strides[0], q_strides[1], k.data_ptr(), k_strides[0], k_strides[1],
            v.data_ptr(), v_strides[0], v_strides[1], out.data_ptr(), o_strides[0], o_strides[1], B, H, Lq, Lk, hd,
            _DT[q.dtype], float(scale), _ptr(bias), bias.shape[1] if bias is not None else 0, bias_zero, _ptr(kmask),
            kv_div, int(causal), q_pos0, _ptr(lse))
    if drop is not None:
        use_tc = use_tc and hd == 64 and q.dtype == torch.bfloat16
        _lib.call("mrb_attention_fwd_tc_drop" if use_tc else "mrb_attention_fwd_drop", *args, drop[0].data_ptr(), drop[1],
                  float(drop[2]), _stream())
    else:
        _lib.call("mrb_attention_fwd_tc" if use_tc else "mrb_attention_fwd", *args, _stream())
    return out


def attention_row(q, k, v, out, B, H, Lk, hd, scale, q_bs, k_strides, v_strides, o_bs):
    """One query row per (batch, head): q/out data_ptr = that row of batch 0."""
    _lib.call("mrb_attention_row", q.data_ptr(), q_bs, k.data_ptr(), k_strides[0], k_strides[1], v.data_ptr(), v_strides[0],
              v_strides[1], out.data_ptr(), o_bs, B, H, Lk, hd, _DT[q.dtype], float(scale), _stream())


def attention_vit_ok(L, hd):
    """Shapes the persistent ViT attention kernel (csrc/attention_vit.cu) is specialised for."""
    return USE_VIT_ATTENTION and L == 257 and 64 < hd <= 96 and hd % 8 == 0


def attention_vit(q, k, v, out, frames, H, L, hd, scale, q_strides, k_strides, v_strides, o_strides):
    """EVA ViT self-attention over every (frame, head): all 257 query rows in one launch (CLS row included)."""
    _check(q, torch.float16, torch.bfloat16)
    _lib.call("mrb_attention_vit", q.data_ptr(), q_strides[0], q_strides[1], k.data_ptr(), k_strides[0], k_strides[1],
              v.data_ptr(), v_strides[0], v_strides[1], out.data_ptr(), o_strides[0], o_strides[1], frames, H, L, hd,
              _DT[q.dtype], float(scale), _stream())
    return out


def attention_bwd(q, k, v, o, dout, dq, dk, dv, B, H, Lq, Lk, hd, scale, q_strides, k_strides, v_strides, o_strides,
                  do_strides, lse, delta_ws, bias=None, bias_zero=0, kmask=None, causal=False, q_pos0=0, impl="auto", drop=None):
    """drop: the (seed word tensor, site, p) the forward call was given, or None."""
    use_tc = impl == "tc" or (impl == "auto" and USE_TC_ATTENTION and hd == 64 and (Lq >= 128 or Lk >= 512) and not few_queries(Lq, Lk, hd))
    args = (q.data_ptr(), q_strides[0], q_strides[1], k.data_ptr(), k_strides[0], k_strides[1],
            v.data_ptr(), v_strides[0], v_strides[1], o.data_ptr(), o_strides[0], o_strides[1], dout.data_ptr(),
            do_strides[0], do_strides[1], dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, H, Lq, Lk, hd, _DT[q.dtype],
            float(scale), _ptr(bias), bias.shape[1] if bias is not None else 0, bias_zero, _ptr(kmask), int(causal),
            q_pos0, lse.data_ptr(), delta_ws.data_ptr())
    if drop is not None:
        _lib.call("mrb_attention_bwd_tc_drop" if use_tc else "mrb_attention_bwd_drop", *args, drop[0].data_ptr(), drop[1],
                  float(drop[2]), _stream())
    else:
        _lib.call("mrb_attention_bwd_tc" if use_tc else "mrb_attention_bwd", *args, _stream())


def norm(x, w, bias, eps, mode, add=None, out_f32=None, out_h=None, sum_out=None, ld_h=None):
    """mode 0 LayerNorm / 1 RMSNorm over the last dim of fp32 x [rows, C] (+ add)."""
    _check(x, torch.float32)
    rows, C = x.shape
    assert x.is_contiguous()
    _lib.call("mrb_norm", x.data_ptr(), _ptr(add), w.data_ptr(), _ptr(bias), float(eps), rows, C, mode, _ptr(out_f32),
              _ptr(out_h), _DT[out_h.dtype] if out_h is not None else 0,
              (ld_h if ld_h is not None else (out_h.stride(0) if out_h is not None else 0)), _ptr(sum_out), _stream())


def rmsnorm_bwd(x, w, dy, eps, dres, lora_A=None, R=0):
    """dres += RMSNorm'(x; w) . dy; dy fp32 or 16-bit (row stride dy.stride(0)); lora_A folds dy[:, C:C+R] . A in."""
    rows, C = x.shape
    assert dres.is_contiguous() and x.is_contiguous()
    _lib.call("mrb_rmsnorm_bwd", x.data_ptr(), w.data_ptr(), dy.data_ptr(), _DT[dy.dtype], dy.stride(0), _ptr(lora_A), R,
              float(eps), rows, C, dres.data_ptr(), _stream())


def lora_up_add(x_ext, A, R, M, K, acc=None):
    _lib.call("mrb_lora_up_add", x_ext.data_ptr(), x_ext.stride(0), A.data_ptr(), R, M, K, _DT[x_ext.dtype], _ptr(acc), _stream())


def patchify(img, out, img_size, patch):
    _check(img, torch.float32)
    assert img.is_contiguous()
    _lib.call("mrb_patchify", img.data_ptr(), out.data_ptr(), _DT[out.dtype], img.shape[0], img_size, patch,
              out.stride(0), _stream())


def patchify_u8(img, out, img_size, patch, mean, std):
    """Raw uint8 frames [F,3,S,S] -> normalised 16-bit patch matrix ((x/255 - mean) / std fused)."""
    _check(img, torch.uint8)
    assert img.is_contiguous()
    _lib.call("mrb_patchify_u8", img.data_ptr(), out.data_ptr(), _DT[out.dtype], img.shape[0], img_size, patch, out.stride(0),
              float(mean[0]), float(mean[1]), float(mean[2]), float(std[0]), float(std[1]), float(std[2]), _stream())


def cls_pos(cls, pos, x, frames, tokens, C):
    _lib.call("mrb_cls_pos", cls.data_ptr(), pos.data_ptr(), x.data_ptr(), frames, tokens, C, _stream())


def gated_gelu_fwd(ab, h, M, F):
    _lib.call("mrb_gated_gelu_fwd", ab.data_ptr(), h.data_ptr(), M, F, h.stride(0), _DT[ab.dtype], _stream())


def gated_gelu_bwd(ab, dh, dab, M, F):
    _lib.call("mrb_gated_gelu_bwd", ab.data_ptr(), dh.data_ptr(), dh.stride(0), dab.data_ptr(), dab.stride(0), M, F,
              _DT[ab.dtype], _stream())


def gather_rows(idx, emb, frames, out):
    _check(idx, torch.int32)
    rows, C = out.shape
    _lib.call("mrb_gather_rows", idx.data_ptr(), emb.data_ptr(), _ptr(frames), out.data_ptr(), rows, C, _stream())


def scatter_frames(idx, dout, dframes):
    rows, C = dout.shape
    _lib.call("mrb_scatter_frames", idx.data_ptr(), dout.data_ptr(), dframes.data_ptr(), rows, C, _stream())


def group_mean(x, out, groups, n, C):
    _lib.call("mrb_group_mean", x.data_ptr(), out.data_ptr(), groups, n, C, _stream())


def group_mean_bwd(dout, dx, groups, n, C):
    _lib.call("mrb_group_mean_bwd", dout.data_ptr(), dx.data_ptr(), groups, n, C, _stream())


def cross_entropy(logits, labels, row_loss=None, dlogits=None, gscale=1.0, loss_sum=None):
    _check(logits, torch.float32)
    _check(labels, torch.int64)
    rows, V = logits.shape
    assert logits.is_contiguous()
    _lib.call("mrb_cross_entropy", logits.data_ptr(), labels.data_ptr(), rows, V, _ptr(row_loss), _ptr(dlogits),
              _DT[dlogits.dtype] if dlogits is not None else 0, dlogits.stride(0) if dlogits is not None else 0,
              float(gscale), _ptr(loss_sum), _stream())


def lora_down(x_ext, A, M, K, R):
    _check(A, torch.float32)
    _lib.call("mrb_lora_down", x_ext.data_ptr(), x_ext.stride(0), A.data_ptr(), M, K, R, _DT[x_ext.dtype], _stream())


def lora_pack(table, n, blocks_per_linear):
    """table: int64 [n, 13] device array of LoraPackDesc records (T5Engine.refresh)."""
    _check(table, torch.int64)
    _lib.call("mrb_lora_pack", table.data_ptr(), n, blocks_per_linear, BF16, _stream())


def skinny_wgrad(P, ldp, Q, ldq, M, C, out, transposed_out, dtype, impl="auto"):
    """out (+)= P[M,C]^T . Q[M,8]; tensor-core kernel for large M (Q then needs 16 readable columns), CUDA-core otherwise."""
    tc = impl == "tc" or (impl == "auto" and M >= 1024)
    _lib.call("mrb_skinny_wgrad_tc" if tc else "mrb_skinny_wgrad", P, ldp, Q, ldq, M, C, out.data_ptr(), int(transposed_out), dtype, _stream())


def skinny_wgrad_pair(P, ldp, Q, ldq, M, C, out, out2, transposed_out, dtype):
    """Tensor-core pass over P for two adjacent 8-column slots of Q: out += P^T Q[:, :8], out2 += P^T Q[:, 8:16]."""
    _lib.call("mrb_skinny_wgrad_tc2", P, ldp, Q, ldq, M, C, out.data_ptr(), out2.data_ptr(), int(transposed_out), dtype, _stream())


def down32(x, W, out, M):
    """out[:, :32] = x[:M] . W^T for a 32-row 16-bit W: tensor-core GEMM, or the one-block-per-row kernel for tiny M."""
    if M >= 256:
        return gemm(x, W, out=out, M=M)
    _lib.call("mrb_small_down", x.data_ptr(), x.stride(0), W.data_ptr(), W.stride(0), M, x.shape[1], out.data_ptr(),
              out.stride(0), _DT[x.dtype], _stream())
    return out


def cast_to(x, out):
    assert x.is_contiguous() and out.is_contiguous()
    _lib.call("mrb_cast_f32_to_h", x.data_ptr(), out.data_ptr(), x.numel(), _DT[out.dtype], _stream())
    return out


def cast2d(x, out, rows, cols):
    _lib.call("mrb_cast2d_f32_to_h", x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, cols, _DT[out.dtype], _stream())


def transpose16(x, out, rows, cols):
    _lib.call("mrb_transpose16", x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, cols, _stream())


def colsum(x, out):
    rows, C = x.shape
    _lib.call("mrb_colsum", x.data_ptr(), rows, C, out.data_ptr(), _stream())


def axpby(x, y, a, b):
    _lib.call("mrb_axpby", x.data_ptr(), y.data_ptr(), x.numel(), float(a), float(b), _stream())


# ---------------------------------------------------------------------------------------------- train-mode dropout
# (csrc/dropmask.cuh, csrc/dropout.cu, mr_blip_b200/dropout.py).  `seed` is a one-element int32 / uint32 CUDA tensor the host
# rewrites before each step; `site` the id of the dropout call; p the drop probability.
def dropout(x, out, rows, cols, seed, site, p):
    """out = drop(x): fp32 -> fp32 / 16-bit, or 16-bit -> same type; 2-D operands with unit column stride (out may be x)."""
    _lib.call("mrb_dropout", x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, cols, _DT[x.dtype], _DT[out.dtype],
              seed.data_ptr(), site, float(p), _stream())
    return out


def dropout_add_norm(resid, branch, out, w, eps, xn, seed, site, p):
    """out = resid + drop(branch) (fp32, contiguous) and xn[:, :C] = T5 RMSNorm(out) * w (16-bit, row stride xn.stride(0)): one pass."""
    assert resid.is_contiguous() and branch.is_contiguous() and out.is_contiguous() and xn.stride(1) == 1
    rows, C = branch.shape
    _lib.call("mrb_dropout_add_norm", resid.data_ptr(), branch.data_ptr(), w.data_ptr(), float(eps), rows, C, xn.data_ptr(),
              _DT[xn.dtype], xn.stride(0), out.data_ptr(), seed.data_ptr(), site, float(p), _stream())
    return out


def rmsnorm_bwd_drop(x, w, dy, eps, dres, dy_next, seed, site, p):
    """rmsnorm_bwd that also writes dy_next[:, :C] = drop_site(dres) (16-bit): the next sublayer's masked dgrad operand."""
    rows, C = x.shape
    assert dres.is_contiguous() and x.is_contiguous() and dy_next.stride(1) == 1
    _lib.call("mrb_rmsnorm_bwd_drop", x.data_ptr(), w.data_ptr(), dy.data_ptr(), _DT[dy.dtype], dy.stride(0), float(eps), rows, C,
              dres.data_ptr(), dy_next.data_ptr(), _DT[dy_next.dtype], dy_next.stride(0), seed.data_ptr(), site, float(p), _stream())


def dropout_add(resid, branch, out, seed, site, p):
    """out = resid + drop(branch): fp32 [rows, cols], all contiguous."""
    assert resid.is_contiguous() and branch.is_contiguous() and out.is_contiguous()
    rows, cols = branch.shape
    _lib.call("mrb_dropout_add", resid.data_ptr(), branch.data_ptr(), out.data_ptr(), rows, cols, seed.data_ptr(), site, float(p),
              _stream())
    return out


def gated_gelu_fwd_drop(ab, h, M, F, seed, site, p):
    _lib.call("mrb_gated_gelu_fwd_drop", ab.data_ptr(), h.data_ptr(), M, F, h.stride(0), _DT[ab.dtype], seed.data_ptr(), site,
              float(p), _stream())


def gated_gelu_bwd_drop(ab, dh, dab, M, F, seed, site, p):
    _lib.call("mrb_gated_gelu_bwd_drop", ab.data_ptr(), dh.data_ptr(), dh.stride(0), dab.data_ptr(), dab.stride(0), M, F,
              _DT[ab.dtype], seed.data_ptr(), site, float(p), _stream())


def lora_down_drop(x, A_down, out, M, K, nlin, seed, site0, p):
    """out[:M, :32] = [drop_j(x) . A_j^T]_j (16-bit): x [M, >=K], A_down [32, K] = the group's stacked lora_A."""
    _lib.call("mrb_lora_down_drop", x.data_ptr(), x.stride(0), A_down.data_ptr(), A_down.stride(0), M, K, nlin, out.data_ptr(),
              out.stride(0), _DT[x.dtype], seed.data_ptr(), site0, float(p), _stream())
    return out


def lora_wgrad_drop(x, ldx, q, ldq, M, K, dA, dtype, seed, site, p):
    """dA[8, K] += (q[M, 8])^T drop(x[M, K]); x / q are raw device addresses (column offsets applied by the caller)."""
    _lib.call("mrb_lora_wgrad_drop", x, ldx, q, ldq, M, K, dA.data_ptr(), dtype, seed.data_ptr(), site, float(p), _stream())


def lora_dx_drop(q, A_down, nlin, dx, M, K, seed, site0, p):
    """dx[:M, :K] += sum_j mask_j * (q[:, 8j:8j+8] . A_j): q = the 32 extension columns of the dgrad operand."""
    _lib.call("mrb_lora_dx_drop", q.data_ptr(), q.stride(0), A_down.data_ptr(), A_down.stride(0), nlin, dx.data_ptr(), dx.stride(0),
              _DT[dx.dtype], M, K, _DT[q.dtype], seed.data_ptr(), site0, float(p), _stream())
    return dx
