"""FlanT5-XL with LoRA on the GPU: teacher-forced forward + hand-written backward (LoRA A/B
gradients and the gradient w.r.t. inputs_embeds), encoder-only forward, and incremental decoding
with cached K/V.  Restates lavis/models/blip2_models/modeling_t5.py (T5Stack.forward :1021-1282,
T5Attention :474-620, T5LayerFF :331-347, T5ForConditionalGeneration.forward :1734-1893) and peft's
lora.Linear as configured at blip2_mr.py:193-200, kernel by kernel through the C ABI.

Data layout (HBM):
  * residual stream h: fp32 [M, d_model] (reference keeps fp32 residuals under bf16 autocast).
  * every Linear input lives in an "extended" bf16 buffer [M, K+32]: columns [0,K) hold x, columns
    [K, K+8j+8) hold the LoRA down-projections x.A_j^T of the (up to 3) Linears sharing that input
    (one N = 32 GEMM).  The frozen weight is stored once as W_ext = [W | sB_0 sB_1 sB_2 | 0]
    (bf16 [N, K+32]), so base(x) + B(A x) is ONE tcgen05 GEMM.  Backward mirrors this with
    [M, N+32] gradient buffers and [W^T | A^T | 0] weights (see LoraGroup).
  * LoRA A/B change every optimiser step: refresh() re-copies them into their 8-column slots.
Train-mode dropout (T5 0.1, LoRA inputs 0.05): on while T5Engine.drop is a dropout.DropState (BLIP2_MR sets it in train mode
unless built with train_dropout=False / MRB_TRAIN_DROPOUT=0); masks are the counter hash of csrc/dropmask.cuh.
"""
import math

import torch

from . import dropout as dr
from . import ops
from .dims import Dims, T5_PREFIX

BF = torch.bfloat16
EXT = 32


def _f(t):
    return t.detach().to(device="cuda", dtype=torch.float32).contiguous()


class LoraGroup:
    """Linears that share one input (q|k|v, wi_0|wi_1, or a single Linear) packed so that forward AND backward are each
    one skinny "down" GEMM (N = 32) plus one full tcgen05 GEMM:
        forward : x_ext[:, K:]  = x . A_down^T            ;  y  = x_ext  . [W | sB]^T
        backward: dy_ext[:, N:] = dy . B_down^T (= dy.sB) ;  dx = dy_ext . [W^T | A^T]^T   (= dy.W + (dy.sB).A)
    LoRA weight gradients: dB_j = dy_j^T (x A_j^T)  and  dA_j = (dy_j sB_j)^T x  -- both skinny reductions over M."""

    def __init__(self, get, names, scale, eng=None):
        self.eng = eng            # T5Engine: owner of the side stream that runs the weight-gradient reductions
        self.names = names
        Ws = [get(n + ".base_layer.weight") for n in names]
        self.K = K = Ws[0].shape[1]
        self.Ns = [w.shape[0] for w in Ws]
        self.N = N = sum(self.Ns)
        self.n = len(names)
        self.site0 = dr.lora_site(names[0])                   # LoRA input dropout: Linear j of the group draws at site0 + j
        assert [dr.lora_site(n) for n in names] == [self.site0 + j for j in range(self.n)]
        self.scale = scale
        self.A_params = [get(n + ".lora_A.default.weight") for n in names]
        self.B_params = [get(n + ".lora_B.default.weight") for n in names]
        assert all(a.shape[0] == 8 for a in self.A_params), "kernels are specialised for LoRA r = 8"
        self.offs = [sum(self.Ns[:j]) for j in range(self.n)]
        self.ext = torch.zeros((N, K + EXT), dtype=BF, device="cuda")          # [W | sB | 0]
        off = 0
        for w in Ws:
            self.ext[off:off + w.shape[0], :K] = w.detach().to("cuda")
            off += w.shape[0]
        self.ext_b = torch.zeros((K, N + EXT), dtype=BF, device="cuda")        # [W^T | A^T | 0]
        ops.transpose16(self.ext, self.ext_b, N, K)
        self.A_down = torch.zeros((EXT, K), dtype=BF, device="cuda")
        self.B_down = torch.zeros((EXT, N), dtype=BF, device="cuda")
        self.dA = [torch.zeros((8, K), dtype=torch.float32, device="cuda") for _ in names]
        self.dB = [torch.zeros((n_, 8), dtype=torch.float32, device="cuda") for n_ in self.Ns]
        self._dst = []
        for j in range(self.n):
            o, n_ = self.offs[j], self.Ns[j]
            self._dst += [self.ext[o:o + n_, K + 8 * j:K + 8 * j + 8], self.B_down[8 * j:8 * j + 8, o:o + n_],
                          self.A_down[8 * j:8 * j + 8], self.ext_b[:, N + 8 * j:N + 8 * j + 8]]
        self.refresh()

    def refresh(self):
        """Copy the (updated) LoRA A/B of this group into their 16-bit slots; T5Engine.refresh does all groups in one launch."""
        src = []
        for j in range(self.n):
            b = self.B_params[j].detach()
            if self.scale != 1.0:
                b = b * self.scale
            a = self.A_params[j].detach()
            src += [b, b.t(), a, a.t()]
        torch._foreach_copy_(self._dst, src)

    def pack_records(self):
        """One LoraPackDesc (csrc/elementwise.cu) per Linear of the group: 13 x int64."""
        import numpy as np
        recs = []
        for j in range(self.n):
            ext_slot, bdown_slot, adown_slot, extb_slot = self._dst[4 * j:4 * j + 4]
            a, b = self.A_params[j], self.B_params[j]
            assert a.is_contiguous() and b.is_contiguous() and a.dtype == torch.float32 and b.dtype == torch.float32
            recs.append([a.data_ptr(), b.data_ptr(), ext_slot.data_ptr(), self.ext.stride(0), bdown_slot.data_ptr(),
                         self.B_down.stride(0), adown_slot.data_ptr(), self.A_down.stride(0), extb_slot.data_ptr(),
                         self.ext_b.stride(0), self.K, self.Ns[j], int(np.float64(self.scale).view(np.int64))])
        return recs

    def bind_grads(self, flat, off):
        """Re-point dA/dB at consecutive views of one flat fp32 buffer (order = grads()). -> new offset."""
        for j in range(self.n):
            nA, nB = 8 * self.K, self.Ns[j] * 8
            self.dA[j] = flat[off:off + nA].view(8, self.K)
            off += nA
            self.dB[j] = flat[off:off + nB].view(self.Ns[j], 8)
            off += nB
        return off

    def _drop(self):
        dp = self.eng.drop if self.eng is not None else None
        return dp if (dp is not None and dp.lora > 0.0) else None

    def down(self, x_ext, M):
        """x_ext[:, K:K+32] = x_ext[:, :K] . A_down^T  (the LoRA down-projections of the Linears in this group); in train mode
        every Linear j sees its own drop_j(x) (peft: lora_B(lora_A(dropout(x))))."""
        dp = self._drop()
        if dp is not None:
            ops.lora_down_drop(x_ext[:, :self.K], self.A_down, x_ext[:, self.K:], M, self.K, self.n, dp.word, self.site0, dp.lora)
        else:
            ops.down32(x_ext[:, :self.K], self.A_down, x_ext[:, self.K:], M)

    def dgrad(self, dy_ext, M, out=None, resid=None, out_dtype=BF):
        """dx = dy . W + sum_j mask_j * ((dy sB_j) . A_j): one GEMM over the extended K in eval mode; with LoRA dropout the dense
        part is the GEMM over K = N and the rank-8 terms are added under their masks (mrb_lora_dx_drop)."""
        dp = self._drop()
        if dp is None:
            return ops.gemm(dy_ext, self.ext_b, out=out, resid=resid, out_dtype=out_dtype, M=M)
        dx = ops.gemm(dy_ext, self.ext_b, out=out, resid=resid, out_dtype=out_dtype, M=M, K=self.N)
        return ops.lora_dx_drop(dy_ext[:, self.N:], self.A_down, self.n, dx, M, self.K, dp.word, self.site0, dp.lora)

    def forward(self, x_ext, M, out=None, resid=None, out_dtype=BF):
        self.down(x_ext, M)
        return ops.gemm(x_ext, self.ext, out=out, resid=resid, out_dtype=out_dtype, M=M)

    def backward(self, dy_ext, x_ext, M, out=None, resid=None, out_dtype=BF):
        """dy_ext [M, N+32] (columns [0,N) filled by the caller), x_ext = the saved forward input [M, K+32].
        -> dx [M, K] (complete, LoRA term included); accumulates dA_j, dB_j."""
        K, N = self.K, self.N
        ops.down32(dy_ext[:, :N], self.B_down, dy_ext[:, N:], M)
        if self.eng is not None and self.eng.overlap:
            # dA / dB feed nothing downstream: reduce them on the side stream while the main stream goes on with dgrad
            with self.eng.side_block(hold=(dy_ext, x_ext)):
                self._wgrads(dy_ext, x_ext, M)
            return self.dgrad(dy_ext, M, out=out, resid=resid, out_dtype=out_dtype)
        dx = self.dgrad(dy_ext, M, out=out, resid=resid, out_dtype=out_dtype)
        self._wgrads(dy_ext, x_ext, M)
        return dx

    def _wgrads(self, dy_ext, x_ext, M):
        K, N = self.K, self.N
        es = 2
        for j in range(self.n):
            o, n_ = self.offs[j], self.Ns[j]
            ops.skinny_wgrad(dy_ext.data_ptr() + o * es, dy_ext.stride(0), x_ext.data_ptr() + (K + 8 * j) * es,
                             x_ext.stride(0), M, n_, self.dB[j], False, ops.BF16)
        dp = self._drop()
        if dp is not None:                                  # dA_j = (dy sB_j)^T drop_j(x): the mask is recomputed per Linear
            for j in range(self.n):
                ops.lora_wgrad_drop(x_ext.data_ptr(), x_ext.stride(0), dy_ext.data_ptr() + (N + 8 * j) * es, dy_ext.stride(0), M, K,
                                    self.dA[j], ops.BF16, dp.word, self.site0 + j, dp.lora)
            return
        j = 0
        while j < self.n:                                   # dA_j = (dy sB_j)^T x: two slots per pass over x when M is large
            q = dy_ext.data_ptr() + (N + 8 * j) * es
            if M >= 1024 and j + 1 < self.n:
                ops.skinny_wgrad_pair(x_ext.data_ptr(), x_ext.stride(0), q, dy_ext.stride(0), M, K, self.dA[j],
                                      self.dA[j + 1], True, ops.BF16)
                j += 2
            else:
                ops.skinny_wgrad(x_ext.data_ptr(), x_ext.stride(0), q, dy_ext.stride(0), M, K, self.dA[j], True, ops.BF16)
                j += 1

    def zero_grads(self):
        for g in self.dA + self.dB:
            g.zero_()

    def grads(self):
        out = []
        for j in range(self.n):
            out += [(self.A_params[j], self.dA[j]), (self.B_params[j], self.dB[j])]
        return out


_BUCKET_CACHE = {}


def shift_right(labels):
    """_shift_right (modeling_t5.py:919-948): decoder_start_token_id = pad = 0, -100 -> pad."""
    dec_ids = torch.zeros_like(labels)
    dec_ids[:, 1:] = labels[:, :-1]
    dec_ids.masked_fill_(dec_ids == -100, 0)
    return dec_ids


def _bucket_index(Lq, Lk, bidirectional, num_buckets, max_distance):
    """Bucket id for every delta = j - i in [-(Lq-1), Lk-1], computed on the HOST with the reference's own
    float formula (modeling_t5.py:393-445) so bucket boundaries agree bit-for-bit with the CPU oracle."""
    key = (Lq, Lk, bidirectional, num_buckets, max_distance)
    if key not in _BUCKET_CACHE:
        rel = torch.arange(-(Lq - 1), Lk, dtype=torch.long)
        nb = num_buckets
        buckets = torch.zeros_like(rel)
        if bidirectional:
            nb //= 2
            buckets = buckets + (rel > 0).to(torch.long) * nb
            rel = torch.abs(rel)
        else:
            rel = -torch.min(rel, torch.zeros_like(rel))
        max_exact = nb // 2
        is_small = rel < max_exact
        large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                             * (nb - max_exact)).to(torch.long)
        large = torch.min(large, torch.full_like(large, nb - 1))
        buckets = buckets + torch.where(is_small, rel, large)
        _BUCKET_CACHE[key] = buckets.to("cuda")
    return _BUCKET_CACHE[key]


class T5Engine:
    def __init__(self, d: Dims, get, prefix=T5_PREFIX):
        self.d = d
        D = d.d_model
        scale = d.lora_alpha / d.lora_r
        self.emb = get(prefix + "shared.weight")          # fp32 [V, D] (embed_tokens is tied to it)
        assert self.emb.is_cuda
        self.groups = []
        self.drop = None                                  # dropout.DropState while a train-mode step with dropout runs

        # Side stream (forked / joined inside the captured step): LoRA weight-gradient reductions and the decoder's
        # encoder-sized cross-attention K/V GEMMs run next to the latency-bound main chain.  MRB_OVERLAP=0 disables.
        import os
        self.overlap = os.environ.get("MRB_OVERLAP", "1") != "0"
        self.side = torch.cuda.Stream()
        ops.splitk_register(self.side)
        # SM cap of the side stream's GEMMs (the decoder's 48 encoder-sized cross K/V GEMMs): uncapped, each holds every SM's
        # shared memory for ~100 us and the main chain's small kernels wait behind it.  Decoder chain in-graph on a B200 (call 27):
        # 18.36 ms uncapped, 17.73 at 132 SMs, 17.88 / 17.74 / 18.12 at 116 / 100 / 84.  MRB_SIDE_SMS=0 lifts the cap.
        self.side_sms = int(os.environ.get("MRB_SIDE_SMS", "132"))      # applied by side_block() around its launches
        self._hold = []
        self._side_open = False
        # Train-mode fusions of the residual stream's elementwise passes (MRB_T5_FUSE_NORM=0 disables): dropout-add + the next
        # T5LayerNorm in one kernel, RMSNorm backward + the next sublayer's masked 16-bit gradient in one kernel.
        # B200, each phase as its own graph (call 28, fused off -> on): encoder forward 28.66-28.74 -> 27.92-28.29 ms, encoder
        # backward 52.49-52.95 -> 51.82-51.98 ms, but the decoder chain (64 rows) 17.83 -> 18.15-18.19 ms: encoder-sized inputs only
        # (MRB_T5_FUSE_MIN_ROWS, default 512; 0 = every size).
        self.fuse_norm = os.environ.get("MRB_T5_FUSE_NORM", "1") != "0"
        self.fuse_min_rows = int(os.environ.get("MRB_T5_FUSE_MIN_ROWS", "512"))
        self._pre_norm = None                             # (residual stream, ln weight, normalised operand) of the fused add
        self._pre_grad = None                             # (gradient stream, site, masked 16-bit operand) of the fused backward

        def grp(names):
            g = LoraGroup(get, names, scale, self)
            self.groups.append(g)
            return g

        def attn(p):
            return dict(qkv=grp([p + ".q", p + ".k", p + ".v"]), o=grp([p + ".o"]))

        self.enc, self.dec = [], []
        for i in range(d.t5_layers):
            b = f"{prefix}encoder.block.{i}."
            L = attn(b + "layer.0.SelfAttention")
            L.update(ln0=_f(get(b + "layer.0.layer_norm.weight")), ln1=_f(get(b + "layer.1.layer_norm.weight")),
                     wi=grp([b + "layer.1.DenseReluDense.wi_0", b + "layer.1.DenseReluDense.wi_1"]),
                     wo=grp([b + "layer.1.DenseReluDense.wo"]))
            L["stack"], L["li"] = dr.ENC, i
            self.enc.append(L)
        self.enc_bias = _f(get(prefix + "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"))
        self.enc_final_ln = _f(get(prefix + "encoder.final_layer_norm.weight"))
        for i in range(d.t5_dec_layers):
            b = f"{prefix}decoder.block.{i}."
            L = attn(b + "layer.0.SelfAttention")
            c = b + "layer.1.EncDecAttention"
            L.update(ln0=_f(get(b + "layer.0.layer_norm.weight")), ln1=_f(get(b + "layer.1.layer_norm.weight")),
                     ln2=_f(get(b + "layer.2.layer_norm.weight")),
                     cq=grp([c + ".q"]), ckv=grp([c + ".k", c + ".v"]), co=grp([c + ".o"]),
                     wi=grp([b + "layer.2.DenseReluDense.wi_0", b + "layer.2.DenseReluDense.wi_1"]),
                     wo=grp([b + "layer.2.DenseReluDense.wo"]))
            L["stack"], L["li"] = dr.DEC, i
            self.dec.append(L)
        self.dec_bias = _f(get(prefix + "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"))
        self.dec_final_ln = _f(get(prefix + "decoder.final_layer_norm.weight"))
        self.lm_head = grp([prefix + "lm_head"])
        self._pack_table = None
        self._dec_states = {}
        self._dec_pool = None
        self.decode_graphs = os.environ.get("MRB_CUDA_GRAPHS", "1") != "0"

    # ------------------------------------------------------------------ side stream
    def side_block(self, hold=()):
        """Context: the enclosed launches go to the side stream, ordered after everything issued so far on the main
        stream.  `hold` tensors stay referenced until side_join() so the allocator cannot recycle them early."""
        eng = self

        class _Ctx:
            def __enter__(self):
                eng._hold.extend(hold)
                eng.side.wait_stream(torch.cuda.current_stream())
                eng._side_open = True
                self.cm = torch.cuda.stream(eng.side)
                self.cm.__enter__()
                if eng.side_sms > 0:
                    ops.gemm_sm_limit(eng.side_sms)          # this thread's large GEMMs leave SMs to the main chain

            def __exit__(self, *a):
                if eng.side_sms > 0:
                    ops.gemm_sm_limit(0)
                return self.cm.__exit__(*a)

        return _Ctx()

    def side_join(self):
        """Main stream waits for the side stream; held tensors are released."""
        if self._side_open:
            torch.cuda.current_stream().wait_stream(self.side)
            self._side_open = False
        self._hold.clear()

    # ------------------------------------------------------------------ helpers
    def refresh(self):
        """Re-pack every LoRA A/B after an optimiser step: one kernel over a device-resident descriptor table."""
        if self._pack_table is None:
            import numpy as np
            recs = [r for g in self.groups for r in g.pack_records()]
            self._pack_table = torch.from_numpy(np.asarray(recs, dtype=np.int64)).to("cuda")
            self._pack_blocks = max(1, (max(max(g.K, max(g.Ns)) for g in self.groups) + 2047) // 2048)
        ops.lora_pack(self._pack_table, self._pack_table.shape[0], self._pack_blocks)

    def n_grad_elems(self):
        return sum(8 * g.K + 8 * n for g in self.groups for n in g.Ns)

    def bind_grads(self, flat, off=0):
        for g in self.groups:
            off = g.bind_grads(flat, off)
        return off

    def scale_grads(self):
        """dB of a LoRA with alpha != r carries the alpha / r factor (no-op for the reference's alpha = r = 8)."""
        for g in self.groups:
            if g.scale != 1.0:
                torch._foreach_mul_(list(g.dB), g.scale)

    def zero_grads(self):
        for g in self.groups:
            g.zero_grads()

    def param_grads(self):
        out = []
        for g in self.groups:
            out += g.grads()
        return out

    def _bias(self, table, Lq, Lk, bidirectional):
        d = self.d
        idx = _bucket_index(Lq, Lk, bidirectional, d.rel_buckets, d.rel_max_dist)
        return table[idx].t().contiguous()                 # [H, Lq+Lk-1], zero delta at column Lq-1

    def _ext(self, M, K):
        return torch.empty((M, K + EXT), dtype=BF, device="cuda")

    def _pdrop(self, site):
        """Attention-probability dropout (modeling_t5.py:600) of one attention call: (seed word, site, p) or None."""
        return self.drop.attn(site, self.drop.t5) if self.drop is not None else None

    def _self_attn(self, qkv, out_ext, B, L, bias, kmask, causal, lse, site=None):
        d = self.d
        inner = d.t5_heads * d.d_kv
        rs = qkv.stride(0)
        ops.attention_fwd(qkv, qkv[:, inner:], qkv[:, 2 * inner:], out_ext, B, d.t5_heads, L, L, d.d_kv, 1.0,
                          (L * rs, rs), (L * rs, rs), (L * rs, rs), (L * out_ext.stride(0), out_ext.stride(0)),
                          bias=bias, bias_zero=L - 1, kmask=kmask, causal=causal, lse=lse, drop=self._pdrop(site))

    def _site(self, L, slot):
        return dr.site(L["stack"], L["li"], slot)

    def _p(self):
        """T5's dropout_rate while a train-mode step with dropout runs, else 0."""
        return self.drop.t5 if self.drop is not None else 0.0

    def _grad_ext(self, dh, M, site=None):
        """fp32 residual-stream gradient -> 16-bit extended dgrad operand [M, d_model + 32]; with `site` it is the gradient
        of hidden + dropout(branch) w.r.t. the branch, i.e. dh under that site's mask."""
        pre, self._pre_grad = self._pre_grad, None
        if pre is not None and pre[0] is dh and pre[1] == site:
            return pre[2]                                    # written by the previous sublayer's fused RMSNorm backward
        dy = self._ext(M, self.d.d_model)
        if site is not None and self._p() > 0.0:
            ops.dropout(dh, dy, M, self.d.d_model, self.drop.word, site, self._p())
        else:
            ops.cast2d(dh, dy, M, self.d.d_model)
        return dy

    def _branch(self, grp, x_ext, M, h, site, next_ln=None):
        """h + dropout(Linear(x)) (modeling_t5.py:346,652,690) -> new fp32 residual stream.  Eval mode: the residual add is
        the GEMM's epilogue; train mode: the branch leaves the GEMM in fp32 and one pass applies mask and add -- and, when the
        caller names the T5LayerNorm weight that reads the new stream next (`next_ln`), that norm too (picked up by _norm_ext)."""
        out = torch.empty_like(h)
        if self._p() > 0.0:
            br = grp.forward(x_ext, M, out_dtype=torch.float32)
            if next_ln is not None and self.fuse_norm and M >= self.fuse_min_rows:
                xn = self._ext(M, self.d.d_model)
                ops.dropout_add_norm(h, br, out, next_ln, self.d.t5_ln_eps, xn, self.drop.word, site, self._p())
                self._pre_norm = (out, next_ln, xn)
                return out
            return ops.dropout_add(h, br, out, self.drop.word, site, self._p())
        grp.forward(x_ext, M, out=out, resid=h)
        return out

    def _norm_ext(self, h, ln, M):
        """T5LayerNorm(h) * ln as an extended 16-bit operand [M, d_model + 32] (modeling_t5.py:263-277)."""
        pre, self._pre_norm = self._pre_norm, None
        if pre is not None and pre[0] is h and pre[1] is ln:
            return pre[2]                                    # written by the fused residual add of the previous sublayer
        xn = self._ext(M, self.d.d_model)
        ops.norm(h, ln, None, self.d.t5_ln_eps, 1, out_h=xn)
        return xn

    def _rmsnorm_bwd(self, x, ln, dy, dh, M, next_site=None):
        """dh += T5LayerNorm'(x; ln) . dy; with `next_site` (the residual-dropout site of the sublayer whose backward runs next)
        the same pass writes that sublayer's masked 16-bit dgrad operand (picked up by _grad_ext)."""
        if next_site is not None and self._p() > 0.0 and self.fuse_norm and M >= self.fuse_min_rows:
            dyn = self._ext(M, self.d.d_model)
            ops.rmsnorm_bwd_drop(x, ln, dy, self.d.t5_ln_eps, dh, dyn, self.drop.word, next_site, self._p())
            self._pre_grad = (dh, next_site, dyn)
            return
        ops.rmsnorm_bwd(x, ln, dy, self.d.t5_ln_eps, dh)

    def _drop_inplace(self, t, rows, cols, site):
        if self._p() > 0.0:
            ops.dropout(t, t, rows, cols, self.drop.word, site, self._p())
        return t

    def _ff(self, L, h, M, save, next_ln=None):
        d = self.d
        xn = self._norm_ext(h, L["ln_ff"], M)
        ab = L["wi"].forward(xn, M)
        hm = self._ext(M, d.d_ff)
        if self._p() > 0.0:
            ops.gated_gelu_fwd_drop(ab, hm, M, d.d_ff, self.drop.word, self._site(L, dr.FF_INNER), self._p())
        else:
            ops.gated_gelu_fwd(ab, hm, M, d.d_ff)
        h2 = self._branch(L["wo"], hm, M, h, self._site(L, dr.FF_RES), next_ln)
        if save is not None:
            save.update(ff_x=h, ff_xn=xn, ff_ab=ab, ff_hm=hm)
        return h2

    def _ff_bwd(self, L, s, dh, M, next_site=None):
        """dh (fp32 residual-stream gradient) is updated in place."""
        d = self.d
        dhm = L["wo"].backward(self._grad_ext(dh, M, self._site(L, dr.FF_RES)), s["ff_hm"], M)          # [M, d_ff]
        dab = self._ext(M, 2 * d.d_ff)
        if self._p() > 0.0:
            ops.gated_gelu_bwd_drop(s["ff_ab"], dhm, dab, M, d.d_ff, self.drop.word, self._site(L, dr.FF_INNER), self._p())
        else:
            ops.gated_gelu_bwd(s["ff_ab"], dhm, dab, M, d.d_ff)
        dxn = L["wi"].backward(dab, s["ff_xn"], M)
        self._rmsnorm_bwd(s["ff_x"], L["ln_ff"], dxn, dh, M, next_site)

    # ------------------------------------------------------------------ encoder
    def encoder_forward(self, x, kmask, B, L, save=None):
        """x fp32 [B*L, D] inputs_embeds; -> normalised encoder output in an ext bf16 buffer [B*L, D+32]."""
        d = self.d
        M = B * L
        inner = d.t5_heads * d.d_kv
        bias = self._bias(self.enc_bias, L, L, True)
        h = x
        if self._p() > 0.0:                                  # hidden_states = dropout(inputs_embeds), modeling_t5.py:1149
            h = ops.dropout(x, torch.empty_like(x), M, d.d_model, self.drop.word, dr.site(dr.ENC, 0, dr.EMB), self._p())
        for li, layer in enumerate(self.enc):
            layer["ln_ff"] = layer["ln1"]
            s = {} if save is not None else None
            xn = self._norm_ext(h, layer["ln0"], M)
            qkv = self._ext(M, 3 * inner)                    # ext width so that dqkv can share the layout
            layer["qkv"].forward(xn, M, out=qkv[:, :3 * inner])
            ao = self._ext(M, d.d_model)
            lse = torch.empty((B, d.t5_heads, L), dtype=torch.float32, device="cuda") if save is not None else None
            self._self_attn(qkv, ao, B, L, bias, kmask, False, lse, self._site(layer, dr.SELF_P))
            h1 = self._branch(layer["o"], ao, M, h, self._site(layer, dr.SELF_RES), layer["ln_ff"])
            if s is not None:
                s.update(x=h, xn=xn, qkv=qkv, ao=ao, lse=lse)
            h = self._ff(layer, h1, M, s, self.enc[li + 1]["ln0"] if li + 1 < len(self.enc) else self.enc_final_ln)
            if save is not None:
                save.append(s)
        out = self._norm_ext(h, self.enc_final_ln, M)
        self._drop_inplace(out[:, :d.d_model], M, d.d_model, dr.site(dr.ENC, 0, dr.FINAL))      # modeling_t5.py:1258
        return out, h, bias

    def encoder_backward(self, saves, h_last, d_enc_out, kmask, B, L, bias):
        """d_enc_out fp32 [M, D]: gradient w.r.t. the normalised encoder output -> gradient w.r.t. inputs_embeds."""
        d = self.d
        M = B * L
        dh = torch.zeros((M, d.d_model), dtype=torch.float32, device="cuda")
        self._drop_inplace(d_enc_out, M, d.d_model, dr.site(dr.ENC, 0, dr.FINAL))
        self._rmsnorm_bwd(h_last, self.enc_final_ln, d_enc_out, dh, M, self._site(self.enc[-1], dr.FF_RES))
        inner = d.t5_heads * d.d_kv
        ws = torch.empty((B * d.t5_heads * L,), dtype=torch.float32, device="cuda")
        for layer, s in zip(reversed(self.enc), reversed(saves)):
            self._ff_bwd(layer, s, dh, M, self._site(layer, dr.SELF_RES))
            dao = layer["o"].backward(self._grad_ext(dh, M, self._site(layer, dr.SELF_RES)), s["ao"], M)   # [M, D] = dO of the attention
            qkv = s["qkv"]
            rs = qkv.stride(0)
            dqkv = torch.empty_like(qkv)
            st = (L * rs, rs)
            ops.attention_bwd(qkv, qkv[:, inner:], qkv[:, 2 * inner:], s["ao"], dao, dqkv, dqkv[:, inner:],
                              dqkv[:, 2 * inner:], B, d.t5_heads, L, L, d.d_kv, 1.0, st, st, st,
                              (L * s["ao"].stride(0), s["ao"].stride(0)), (L * dao.stride(0), dao.stride(0)), s["lse"], ws,
                              bias=bias, bias_zero=L - 1, kmask=kmask, causal=False, drop=self._pdrop(self._site(layer, dr.SELF_P)))
            dxn = layer["qkv"].backward(dqkv, s["xn"], M)
            self._rmsnorm_bwd(s["x"], layer["ln0"], dxn, dh, M,
                              self._site(self.enc[layer["li"] - 1], dr.FF_RES) if layer["li"] > 0 else None)
            self.side_join()                                 # this layer's weight-gradient reductions are done
            s.clear()
        return self._drop_inplace(dh, M, d.d_model, dr.site(dr.ENC, 0, dr.EMB))      # d inputs_embeds through the input dropout

    # ------------------------------------------------------------------ decoder (teacher forced)
    def decoder_forward(self, dec_ids, dmask, enc_ext, enc_kmask, B, Ld, Le, save=None):
        d = self.d
        M, Me = B * Ld, B * Le
        inner = d.t5_heads * d.d_kv
        h = torch.empty((M, d.d_model), dtype=torch.float32, device="cuda")
        ops.gather_rows(dec_ids.reshape(-1).to(torch.int32), self.emb, None, h)
        self._drop_inplace(h, M, d.d_model, dr.site(dr.DEC, 0, dr.EMB))
        bias = self._bias(self.dec_bias, Ld, Ld, False)
        ckvs, ckv_ready = [], []
        if self.overlap:
            # every layer's cross-attention K/V projection of the encoder output (the only encoder-sized GEMMs of the
            # decoder) depends on enc_ext alone: issue all of them on the side stream, next to the tiny-M main chain
            ckvs = [self._ext(Me, 2 * inner) for _ in self.dec]
            with self.side_block(hold=[enc_ext] + ckvs):
                for layer, ckv in zip(self.dec, ckvs):
                    layer["ckv"].forward(enc_ext, Me, out=ckv[:, :2 * inner])
                    ev = torch.cuda.Event()
                    ev.record()
                    ckv_ready.append(ev)
        for li, layer in enumerate(self.dec):
            layer["ln_ff"] = layer["ln2"]
            s = {} if save is not None else None
            xn = self._norm_ext(h, layer["ln0"], M)
            qkv = self._ext(M, 3 * inner)
            layer["qkv"].forward(xn, M, out=qkv[:, :3 * inner])
            ao = self._ext(M, d.d_model)
            lse = torch.empty((B, d.t5_heads, Ld), dtype=torch.float32, device="cuda") if save is not None else None
            self._self_attn(qkv, ao, B, Ld, bias, dmask, True, lse, self._site(layer, dr.SELF_P))
            h1 = self._branch(layer["o"], ao, M, h, self._site(layer, dr.SELF_RES), layer["ln1"])
            # cross attention
            xn2 = self._norm_ext(h1, layer["ln1"], M)
            cq = self._ext(M, inner)
            layer["cq"].forward(xn2, M, out=cq[:, :inner])
            if self.overlap:
                ckv = ckvs[li]
                torch.cuda.current_stream().wait_event(ckv_ready[li])
            else:
                ckv = self._ext(Me, 2 * inner)
                layer["ckv"].forward(enc_ext, Me, out=ckv[:, :2 * inner])
            co = self._ext(M, d.d_model)
            lse2 = torch.empty((B, d.t5_heads, Ld), dtype=torch.float32, device="cuda") if save is not None else None
            qs, ks = cq.stride(0), ckv.stride(0)
            ops.attention_fwd(cq, ckv, ckv[:, inner:], co, B, d.t5_heads, Ld, Le, d.d_kv, 1.0, (Ld * qs, qs),
                              (Le * ks, ks), (Le * ks, ks), (Ld * co.stride(0), co.stride(0)), kmask=enc_kmask, lse=lse2,
                              drop=self._pdrop(self._site(layer, dr.CROSS_P)))
            h2 = self._branch(layer["co"], co, M, h1, self._site(layer, dr.CROSS_RES), layer["ln_ff"])
            if s is not None:
                s.update(x=h, xn=xn, qkv=qkv, ao=ao, lse=lse, x1=h1, xn2=xn2, cq=cq, ckv=ckv, co=co, lse2=lse2)
            h = self._ff(layer, h2, M, s, self.dec[li + 1]["ln0"] if li + 1 < len(self.dec) else self.dec_final_ln)
            if save is not None:
                save.append(s)
        out = self._norm_ext(h, self.dec_final_ln, M)
        self._drop_inplace(out[:, :d.d_model], M, d.d_model, dr.site(dr.DEC, 0, dr.FINAL))
        self.side_join()
        return out, h, bias

    def decoder_backward(self, saves, h_last, d_out, dmask, enc_ext, enc_kmask, B, Ld, Le, bias):
        """d_out: gradient w.r.t. the normalised decoder output [M, D] (from lm_head.backward).
        -> gradient w.r.t. the normalised encoder output, fp32 [Me, D]."""
        d = self.d
        M, Me = B * Ld, B * Le
        inner = d.t5_heads * d.d_kv
        dh = torch.zeros((M, d.d_model), dtype=torch.float32, device="cuda")
        self._drop_inplace(d_out[:, :d.d_model], M, d.d_model, dr.site(dr.DEC, 0, dr.FINAL))
        self._rmsnorm_bwd(h_last, self.dec_final_ln, d_out, dh, M, self._site(self.dec[-1], dr.FF_RES))
        d_enc = torch.zeros((Me, d.d_model), dtype=torch.float32, device="cuda")
        ws = torch.empty((B * d.t5_heads * Ld,), dtype=torch.float32, device="cuda")
        for layer, s in zip(reversed(self.dec), reversed(saves)):
            self._ff_bwd(layer, s, dh, M, self._site(layer, dr.CROSS_RES))
            # cross attention
            dco = layer["co"].backward(self._grad_ext(dh, M, self._site(layer, dr.CROSS_RES)), s["co"], M)
            cq, ckv = s["cq"], s["ckv"]
            dcq, dckv = torch.empty_like(cq), torch.empty_like(ckv)
            qs, ks = cq.stride(0), ckv.stride(0)
            ops.attention_bwd(cq, ckv, ckv[:, inner:], s["co"], dco, dcq, dckv, dckv[:, inner:], B, d.t5_heads, Ld, Le,
                              d.d_kv, 1.0, (Ld * qs, qs), (Le * ks, ks), (Le * ks, ks),
                              (Ld * s["co"].stride(0), s["co"].stride(0)), (Ld * dco.stride(0), dco.stride(0)), s["lse2"], ws,
                              kmask=enc_kmask, drop=self._pdrop(self._site(layer, dr.CROSS_P)))
            if self.overlap:
                # encoder-sized dgrad + weight gradients of the cross K/V projection: side stream (in order, d_enc accumulates)
                with self.side_block(hold=(dckv, enc_ext, d_enc)):
                    layer["ckv"].down(enc_ext, Me)
                    g = layer["ckv"]
                    ops.down32(dckv[:, :g.N], g.B_down, dckv[:, g.N:], Me)
                    g.dgrad(dckv, Me, out=d_enc, resid=d_enc)
                    g._wgrads(dckv, enc_ext, Me)
            else:
                layer["ckv"].down(enc_ext, Me)               # recompute this layer's x.A^T columns of the shared input
                layer["ckv"].backward(dckv, enc_ext, Me, out=d_enc, resid=d_enc)     # accumulates over the 24 layers
            dxn2 = layer["cq"].backward(dcq, s["xn2"], M)
            self._rmsnorm_bwd(s["x1"], layer["ln1"], dxn2, dh, M, self._site(layer, dr.SELF_RES))
            # self attention
            dao = layer["o"].backward(self._grad_ext(dh, M, self._site(layer, dr.SELF_RES)), s["ao"], M)
            qkv = s["qkv"]
            rs = qkv.stride(0)
            dqkv = torch.empty_like(qkv)
            st = (Ld * rs, rs)
            ops.attention_bwd(qkv, qkv[:, inner:], qkv[:, 2 * inner:], s["ao"], dao, dqkv, dqkv[:, inner:],
                              dqkv[:, 2 * inner:], B, d.t5_heads, Ld, Ld, d.d_kv, 1.0, st, st, st,
                              (Ld * s["ao"].stride(0), s["ao"].stride(0)), (Ld * dao.stride(0), dao.stride(0)), s["lse"], ws,
                              bias=bias, bias_zero=Ld - 1, kmask=dmask, causal=True, drop=self._pdrop(self._site(layer, dr.SELF_P)))
            dxn = layer["qkv"].backward(dqkv, s["xn"], M)
            self._rmsnorm_bwd(s["x"], layer["ln0"], dxn, dh, M,
                              self._site(self.dec[layer["li"] - 1], dr.FF_RES) if layer["li"] > 0 else None)
            if not self.overlap:
                s.clear()
        self.side_join()                                     # d_enc is complete; saved activations may go
        return d_enc                                         # (dh = grad of the frozen embedding rows: dropped)

    # ------------------------------------------------------------------ loss (+ grads)
    def loss(self, inputs_embeds, attention_mask, labels, decoder_attention_mask, backward=True, want_logits=False):
        """T5ForConditionalGeneration.forward with labels (modeling_t5.py:1734-1893).
        inputs_embeds fp32 [B, Le, D] (cuda), attention_mask [B, Le], labels int64 [B, Ld] (-100 = ignore).
        -> dict(loss [1] fp32, logits?, d_inputs_embeds?)   and LoRA grads accumulated in the groups."""
        kmask = attention_mask.to(device="cuda", dtype=torch.int32).contiguous()
        labels = labels.to("cuda")
        dmask = (decoder_attention_mask.to(device="cuda", dtype=torch.int32).contiguous()
                 if decoder_attention_mask is not None else None)
        return self.loss_device(inputs_embeds, kmask, labels, shift_right(labels), dmask, backward, want_logits)

    def loss_device(self, inputs_embeds, kmask, labels, dec_ids, dmask, backward=True, want_logits=False, loss_out=None):
        """Device half of loss(): every argument already lives on the GPU (kmask / dmask int32, labels / dec_ids int64) and
        nothing below synchronises with the host, so the whole call can be captured into a CUDA graph.  The mean over the
        valid targets is taken on the device (mrb_cross_entropy, gscale < 0).  loss_out: optional fp32 [1] to write into."""
        d = self.d
        B, Le, D = inputs_embeds.shape
        Ld = labels.shape[1]
        x = inputs_embeds.reshape(B * Le, D).contiguous()
        enc_saves = [] if backward else None
        dec_saves = [] if backward else None
        with ops.phase("t5_encoder_fwd"):
            enc_ext, enc_h, enc_bias = self.encoder_forward(x, kmask, B, Le, enc_saves)
        with ops.phase("t5_decoder_fwd"):
            dec_ext, dec_h, dec_bias = self.decoder_forward(dec_ids, dmask, enc_ext, kmask, B, Ld, Le, dec_saves)
        M = B * Ld
        logits = self.lm_head.forward(dec_ext, M, out_dtype=torch.float32)       # [M, V] fp32
        flat = labels.reshape(-1).contiguous()
        loss = loss_out if loss_out is not None else torch.empty((1,), dtype=torch.float32, device="cuda")
        loss.zero_()
        dlogits = self._ext(M, d.vocab) if backward else None
        ops.cross_entropy(logits, flat, None, dlogits, -1.0, loss_sum=loss)
        out = {"loss": loss}
        if want_logits:
            out["logits"] = logits.view(B, Ld, d.vocab)
            out["encoder_last_hidden_state"] = enc_ext[:, :D].float().view(B, Le, D)
        if backward:
            with ops.phase("t5_decoder_bwd"):
                ddec = self.lm_head.backward(dlogits, dec_ext, M)
                d_enc = self.decoder_backward(dec_saves, dec_h, ddec, dmask, enc_ext, kmask, B, Ld, Le, dec_bias)
            with ops.phase("t5_encoder_bwd"):
                d_in = self.encoder_backward(enc_saves, enc_h, d_enc, kmask, B, Le, enc_bias)
            self.scale_grads()
            out["d_inputs_embeds"] = d_in.view(B, Le, D)
        return out

    # ------------------------------------------------------------------ generation
    def encode(self, inputs_embeds, attention_mask):
        B, Le, D = inputs_embeds.shape
        kmask = attention_mask.to(device="cuda", dtype=torch.int32).contiguous()
        enc_ext, _, _ = self.encoder_forward(inputs_embeds.reshape(B * Le, D).contiguous(), kmask, B, Le, None)
        return enc_ext, kmask

    def init_decode(self, enc_ext, B, Le, beams, max_len):
        """Project the encoder output to every decoder layer's cross K/V ONCE (the reference re-projects it at
        every step) into a persistent decode state: static cross K/V buffers, ONE self-attention K/V cache
        tensor for all layers, static token / beam-index / mask inputs and a logits output.  Static addresses are what let
        every decode step replay a captured CUDA graph (one graph per position t, captured from the second generate() call
        of a shape on; MRB_CUDA_GRAPHS=0 keeps eager launches)."""
        d = self.d
        inner = d.t5_heads * d.d_kv
        key = (B, Le, beams, max_len)
        st = self._dec_states.get(key)
        if st is None:
            NB, nl = B * beams, len(self.dec)
            st = {"key": key, "B": B, "Le": Le, "beams": beams, "max_len": max_len, "uses": 0, "graphs": {},
                  "cross": [self._ext(B * Le, 2 * inner) for _ in self.dec],
                  "kv": torch.zeros((2 * nl, NB, max_len, inner), dtype=BF, device="cuda"),
                  "tokens": torch.zeros((NB,), dtype=torch.int64, device="cuda"),
                  "beam_idx": torch.arange(NB, dtype=torch.int64, device="cuda"),
                  "kmask": torch.ones((B, Le), dtype=torch.int32, device="cuda"),
                  "logits": torch.empty((NB, d.vocab), dtype=torch.float32, device="cuda"),
                  "bias": self._bias(self.dec_bias, max_len, max_len, False)}
            st["k"] = [st["kv"][2 * i] for i in range(nl)]
            st["v"] = [st["kv"][2 * i + 1] for i in range(nl)]
            self._dec_states = {key: st}                      # one shape at a time (its graphs hold a private pool,
            self._dec_pool = None                             #  which dies with the previous shape's graphs)
        for layer, ckv in zip(self.dec, st["cross"]):
            layer["ckv"].forward(enc_ext, B * Le, out=ckv[:, :2 * inner])
        st["uses"] += 1
        return st

    def reorder_cache(self, st, beam_idx):
        """Beam reorder of the self-attention cache (modeling_t5.py:1923-1954 _reorder_cache), in place."""
        st["kv"].copy_(st["kv"].index_select(1, beam_idx))

    def decode_step(self, st, token_ids, t, enc_kmask, beam_idx=None):
        """One incremental decoder step for token position t: token_ids int64 [B*beams] -> logits fp32 [B*beams, V] (a
        static buffer, overwritten by the next step).  beam_idx (int64 [B*beams]) reorders cache rows [0, t) first."""
        st["tokens"].copy_(token_ids)
        if t == 0:
            st["kmask"].copy_(enc_kmask)
        reorder = beam_idx is not None and t > 0
        if reorder:
            st["beam_idx"].copy_(beam_idx)
        if not (self.decode_graphs and st["uses"] > 1):
            self._decode_step_device(st, t, reorder)
            return st["logits"]
        g = st["graphs"].get((t, reorder))
        if g is None:                                        # first sight of this position in graph mode: capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._dec_pool, capture_error_mode="thread_local"):
                self._decode_step_device(st, t, reorder)
            if self._dec_pool is None:
                self._dec_pool = g.pool()
            st["graphs"][(t, reorder)] = g
        g.replay()
        return st["logits"]

    def _decode_step_device(self, st, t, reorder):
        d = self.d
        token_ids = st["tokens"]
        NB = token_ids.shape[0]
        inner = d.t5_heads * d.d_kv
        Lm = st["max_len"]
        enc_kmask = st["kmask"]
        if reorder:                                          # rows [0, t) of every layer's K and V follow their beams
            live = st["kv"][:, :, :t]
            live.copy_(live.index_select(1, st["beam_idx"]))
        h = torch.empty((NB, d.d_model), dtype=torch.float32, device="cuda")
        ops.gather_rows(token_ids.to(torch.int32), self.emb, None, h)
        for li, layer in enumerate(self.dec):
            layer["ln_ff"] = layer["ln2"]
            xn = self._ext(NB, d.d_model)
            ops.norm(h, layer["ln0"], None, d.t5_ln_eps, 1, out_h=xn)
            qkv = layer["qkv"].forward(xn, NB)
            st["k"][li][:, t] = qkv[:, inner:2 * inner]
            st["v"][li][:, t] = qkv[:, 2 * inner:]
            ao = self._ext(NB, d.d_model)
            ops.attention_fwd(qkv, st["k"][li], st["v"][li], ao, NB, d.t5_heads, 1, t + 1, d.d_kv, 1.0,
                              (3 * inner, 3 * inner), (Lm * inner, inner), (Lm * inner, inner), (ao.stride(0), ao.stride(0)),
                              bias=st["bias"], bias_zero=Lm - 1, q_pos0=t)
            h1 = torch.empty_like(h)
            layer["o"].forward(ao, NB, out=h1, resid=h)
            xn2 = self._ext(NB, d.d_model)
            ops.norm(h1, layer["ln1"], None, d.t5_ln_eps, 1, out_h=xn2)
            cq = layer["cq"].forward(xn2, NB)
            ckv = st["cross"][li]
            co = self._ext(NB, d.d_model)
            Le = st["Le"]
            ks = ckv.stride(0)
            # the beams of one clip are the query rows of ONE attention problem over that clip's cross K/V (read once per
            # clip, not once per beam): batch = clips, Lq = beams
            nbm = st["beams"]
            qs, cs = cq.stride(0), co.stride(0)
            ops.attention_fwd(cq, ckv, ckv[:, inner:], co, NB // nbm, d.t5_heads, nbm, Le, d.d_kv, 1.0, (nbm * qs, qs),
                              (Le * ks, ks), (Le * ks, ks), (nbm * cs, cs), kmask=enc_kmask)
            h2 = torch.empty_like(h)
            layer["co"].forward(co, NB, out=h2, resid=h1)
            h = self._ff(layer, h2, NB, None)
        out = self._ext(NB, d.d_model)
        ops.norm(h, self.dec_final_ln, None, d.t5_ln_eps, 1, out_h=out)
        self.lm_head.forward(out, NB, out=st["logits"])
