"""Frame encoder of the hot path on the GPU: EVA ViT-g (lavis/models/eva_vit.py:324-340), ln_vision
(blip2.py:113-119), the query-only Q-Former (blip2_models/Qformer.py:804-965) and t5_proj
(blip2_mr.py:491), driven kernel by kernel through the C ABI (mr_blip_b200/ops.py).

Precision follows the reference's GPU regime: fp16 GEMM operands with fp32
accumulation, fp32 residual stream, fp32 LayerNorm and softmax statistics.  Everything here is
frozen under task=qformer_freeze_lora except t5_proj, so only forward kernels exist for it.
"""
import torch

from . import dropout as dr
from . import ops
from .dims import Dims, qf_has_cross

H16 = torch.float16


def _h(t):
    return t.detach().to(device="cuda", dtype=H16).contiguous()


def _f(t):
    return t.detach().to(device="cuda", dtype=torch.float32).contiguous()


class VitEngine:
    """Packed fp16 weights + forward of the EVA ViT."""

    def __init__(self, d: Dims, get, prefix="visual_encoder."):
        self.d = d
        W, P = d.vit_width, d.patch
        self.k_patch = ((3 * P * P + 7) // 8) * 8           # 588 -> 592 (16-byte TMA row pitch)
        pw = torch.zeros((W, self.k_patch), dtype=H16, device="cuda")
        pw[:, :3 * P * P] = _h(get(prefix + "patch_embed.proj.weight")).reshape(W, -1)
        self.patch_w, self.patch_b = pw, _f(get(prefix + "patch_embed.proj.bias"))
        self.cls = _f(get(prefix + "cls_token")).reshape(W)
        self.pos = _f(get(prefix + "pos_embed")).reshape(d.vit_tokens, W)
        # 257 = 2 x 128 + 1 tokens: the 257th query row runs in a small kernel on a side stream (forked / joined inside the
        # captured step) next to the tcgen05 kernel that handles the two full tiles.  MRB_OVERLAP=0 disables.
        import os
        self.overlap = os.environ.get("MRB_OVERLAP", "1") != "0"
        self.split = self.overlap and os.environ.get("MRB_VIT_SPLIT", "1") == "1"
        self.side = torch.cuda.Stream()
        ops.splitk_register(self.side)
        self.blocks = []
        for i in range(d.vit_depth):
            b = f"{prefix}blocks.{i}."
            qb, vb = _f(get(b + "attn.q_bias")), _f(get(b + "attn.v_bias"))
            self.blocks.append(dict(
                ln1_w=_f(get(b + "norm1.weight")), ln1_b=_f(get(b + "norm1.bias")),
                qkv_w=_h(get(b + "attn.qkv.weight")), qkv_b=torch.cat([qb, torch.zeros_like(vb), vb]),   # eva_vit.py:120-122
                proj_w=_h(get(b + "attn.proj.weight")), proj_b=_f(get(b + "attn.proj.bias")),
                ln2_w=_f(get(b + "norm2.weight")), ln2_b=_f(get(b + "norm2.bias")),
                fc1_w=_h(get(b + "mlp.fc1.weight")), fc1_b=_f(get(b + "mlp.fc1.bias")),
                fc2_w=_h(get(b + "mlp.fc2.weight")), fc2_b=_f(get(b + "mlp.fc2.bias"))))

    # CLIP normalisation of the video processors (lavis/processors/blip_processors.py:61-70), used when raw uint8 frames come in
    PIXEL_MEAN = (0.48145466, 0.4578275, 0.40821073)
    PIXEL_STD = (0.26862954, 0.26130258, 0.27577711)
    SPLIT_MIN_FRAMES = 32          # two-stream block loop only when each half still fills the GPU (>= 32 frames = 8224 rows)

    def forward(self, image, return_all=False):
        """image fp32 [F,3,S,S] (cuda, already normalised as the dataset yields it) or raw uint8 [F,3,S,S] (normalisation
        fused into the patch extraction) -> residual stream fp32 [F*257, W]."""
        d = self.d
        F_, W, T = image.shape[0], d.vit_width, d.vit_tokens
        M = F_ * T
        A = torch.empty((F_ * d.n_patches, self.k_patch), dtype=H16, device="cuda")
        if image.dtype == torch.uint8:
            ops.patchify_u8(image.contiguous(), A, d.img_size, d.patch, self.PIXEL_MEAN, self.PIXEL_STD)
        else:
            ops.patchify(image.contiguous(), A, d.img_size, d.patch)
        x = torch.empty((M, W), dtype=torch.float32, device="cuda")
        ops.cls_pos(self.cls, self.pos, x, F_, T, W)
        ops.gemm(A, self.patch_w, out=x, bias=self.patch_b, resid=self.pos, row_group=d.n_patches)
        del A
        xn = torch.empty((M, W), dtype=H16, device="cuda")
        qkv = torch.empty((M, 3 * W), dtype=H16, device="cuda")
        ao = torch.empty((M, W), dtype=H16, device="cuda")
        hid = torch.empty((M, d.vit_mlp), dtype=H16, device="cuda")
        hd, Hh = d.vit_head_dim, d.vit_heads
        outs = [x.clone()] if return_all else None
        if self.split and not return_all and F_ >= 2 * self.SPLIT_MIN_FRAMES and ops.attention_vit_ok(T, hd):
            # Frames are independent in the ViT: the two halves of the batch run the 39 blocks on two streams (forked / joined
            # inside the captured step).  The persistent GEMM / attention kernels of the two halves cannot share an SM (shared
            # memory), so they alternate -- but the HBM-bound LayerNorms of one half run under the other half's tensor-core
            # kernels, and a kernel's last partial wave is filled by the other stream's next kernel.  B200, alternating runs on
            # one box: 255.0 / 251.7 -> 253.6 / 249.9 ms per step (call 25), 251.2 / 253.3 -> 249.9 (call 27).  MRB_VIT_SPLIT=0
            # disables; per-launch CUDA-event timing (bench.py's roofline pass) sets self.split = False, overlapping kernels of
            # two streams would be timed into each other.
            Fa = F_ // 2
            main = torch.cuda.current_stream()
            self.side.wait_stream(main)
            for f0, f1, st in ((0, Fa, main), (Fa, F_, self.side)):
                r0, r1 = f0 * T, f1 * T
                with torch.cuda.stream(st):
                    xs, xns, qkvs, aos, hids = x[r0:r1], xn[r0:r1], qkv[r0:r1], ao[r0:r1], hid[r0:r1]
                    rs = 3 * W
                    for blk in self.blocks:
                        ops.norm(xs, blk["ln1_w"], blk["ln1_b"], d.vit_ln_eps, 0, out_h=xns)
                        ops.gemm(xns, blk["qkv_w"], out=qkvs, bias=blk["qkv_b"])
                        ops.attention_vit(qkvs, qkvs[:, W:], qkvs[:, 2 * W:], aos, f1 - f0, Hh, T, hd, hd ** -0.5,
                                          (T * rs, rs), (T * rs, rs), (T * rs, rs), (T * W, W))
                        self._tail(blk, xs, xns, aos, hids)
            main.wait_stream(self.side)
            return x
        for blk in self.blocks:
            ops.norm(x, blk["ln1_w"], blk["ln1_b"], d.vit_ln_eps, 0, out_h=xn)
            ops.gemm(xn, blk["qkv_w"], out=qkv, bias=blk["qkv_b"])
            rs = 3 * W
            if ops.attention_vit_ok(T, hd):
                # CLS + 256 patches: one persistent tcgen05 launch does all 257 query rows of every (frame, head)
                ops.attention_vit(qkv, qkv[:, W:], qkv[:, 2 * W:], ao, F_, Hh, T, hd, hd ** -0.5,
                                  (T * rs, rs), (T * rs, rs), (T * rs, rs), (T * W, W))
                self._tail(blk, x, xn, ao, hid)
                if return_all:
                    outs.append(x.clone())
                continue
            # other token counts: the first floor(T / 128) * 128 query rows as full tcgen05 tiles, the rest in the small kernels
            Tq = (T // 128) * 128 if ops.USE_TC_ATTENTION else 0
            forked = self.overlap and Tq and Tq == T - 1
            if forked:
                main = torch.cuda.current_stream()
                self.side.wait_stream(main)
                with torch.cuda.stream(self.side):
                    ops.attention_row(qkv[Tq:], qkv[:, W:], qkv[:, 2 * W:], ao[Tq:], F_, Hh, T, hd, hd ** -0.5,
                                      T * rs, (T * rs, rs), (T * rs, rs), T * W)
            if Tq:
                ops.attention_fwd(qkv, qkv[:, W:], qkv[:, 2 * W:], ao, F_, Hh, Tq, T, hd, hd ** -0.5,
                                  (T * rs, rs), (T * rs, rs), (T * rs, rs), (T * W, W), impl="tc")
            if forked:
                main.wait_stream(self.side)
            elif Tq == T - 1:
                ops.attention_row(qkv[Tq:], qkv[:, W:], qkv[:, 2 * W:], ao[Tq:], F_, Hh, T, hd, hd ** -0.5,
                                  T * rs, (T * rs, rs), (T * rs, rs), T * W)
            elif Tq < T:
                ops.attention_fwd(qkv[Tq:], qkv[:, W:], qkv[:, 2 * W:], ao[Tq:], F_, Hh, T - Tq, T, hd, hd ** -0.5,
                                  (T * rs, rs), (T * rs, rs), (T * rs, rs), (T * W, W), impl="mma")
            self._tail(blk, x, xn, ao, hid)
            if return_all:
                outs.append(x.clone())
        return outs if return_all else x

    def _tail(self, blk, x, xn, ao, hid):
        """x += proj(attention output); x += fc2(gelu(fc1(LN(x))))  (eva_vit.py:173-176)."""
        d = self.d
        ops.gemm(ao, blk["proj_w"], out=x, bias=blk["proj_b"], resid=x)
        ops.norm(x, blk["ln2_w"], blk["ln2_b"], d.vit_ln_eps, 0, out_h=xn)
        ops.gemm(xn, blk["fc1_w"], out=hid, bias=blk["fc1_b"], gelu=True)
        ops.gemm(hid, blk["fc2_w"], out=x, bias=blk["fc2_b"], resid=x)


class QFormerEngine:
    """ln_vision + query-only Q-Former + t5_proj."""

    def __init__(self, d: Dims, get, prefix="Qformer.bert."):
        self.d = d
        Hq = d.qf_hidden
        self.lnv_w, self.lnv_b = _f(get("ln_vision.weight")), _f(get("ln_vision.bias"))
        self.query_tokens = _f(get("query_tokens")).reshape(d.num_query, Hq)
        self.emb_ln = (_f(get(prefix + "embeddings.LayerNorm.weight")), _f(get(prefix + "embeddings.LayerNorm.bias")))
        self.layers, kv_w, kv_b = [], [], []
        for i in range(d.qf_layers):
            b = f"{prefix}encoder.layer.{i}."

            def lin(n):
                return _h(get(b + n + ".weight")), _f(get(b + n + ".bias"))

            def ln(n):
                return _f(get(b + n + ".weight")), _f(get(b + n + ".bias"))

            sq, sk, sv = lin("attention.self.query"), lin("attention.self.key"), lin("attention.self.value")
            L = dict(self_qkv_w=torch.cat([sq[0], sk[0], sv[0]]).contiguous(), self_qkv_b=torch.cat([sq[1], sk[1], sv[1]]),
                     self_o=lin("attention.output.dense"), self_ln=ln("attention.output.LayerNorm"),
                     ffn_i=lin("intermediate_query.dense"), ffn_o=lin("output_query.dense"), ffn_ln=ln("output_query.LayerNorm"),
                     cross=None)
            if qf_has_cross(d, i):
                ck, cv = lin("crossattention.self.key"), lin("crossattention.self.value")
                L["cross"] = dict(idx=len(kv_w), q=lin("crossattention.self.query"), o=lin("crossattention.output.dense"),
                                  ln=ln("crossattention.output.LayerNorm"))
                kv_w.append(torch.cat([ck[0], cv[0]]))
                kv_b.append(torch.cat([ck[1], cv[1]]))
            self.layers.append(L)
        # every cross-attention layer reads the same image_embeds: one [n_cross*2*H, 1408] K/V projection
        self.kv_w = torch.cat(kv_w).contiguous()
        self.kv_b = torch.cat(kv_b).contiguous()
        self.n_cross = len(kv_w)
        self.proj_w16 = None
        self.proj_b = None
        self.xattn_events = None     # set to a list to time the cross-attention path (K/V projection GEMM + attention cores)
        self.drop = None             # dropout.DropState while a train-mode step with dropout runs (the frozen Q-Former stays in
                                     # train mode in the reference: hidden / attention-probability dropout 0.1, Qformer.py:107,258,287,373)

    def set_t5_proj(self, weight, bias):
        """t5_proj is trainable (blip2_mr.py:291 only sets an attribute on the Module): refresh per step."""
        if self.proj_w16 is None:
            self.proj_w16 = torch.empty(weight.shape, dtype=H16, device="cuda")
        ops.cast_to(weight.detach().contiguous(), self.proj_w16)
        self.proj_b = bias.detach()

    def _attn_block(self, ctx, o, ln, h, h16, M, site=None):
        """LayerNorm(dropout(dense(ctx)) + h) (BertSelfOutput / BertOutput, Qformer.py:278-289,366-375) -> h, h16 in place."""
        d = self.d
        t = torch.empty((M, d.qf_hidden), dtype=torch.float32, device="cuda")
        p = self.drop.qformer if self.drop is not None else 0.0
        if p > 0.0:
            br = ops.gemm(ctx, o[0], bias=o[1], out_dtype=torch.float32)
            ops.dropout_add(h, br, t, self.drop.word, site, p)
        else:
            ops.gemm(ctx, o[0], out=t, bias=o[1], resid=h)
        ops.norm(t, ln[0], ln[1], d.qf_ln_eps, 0, out_f32=h, out_h=h16)

    def _pdrop(self, site):
        return self.drop.attn(site, self.drop.qformer) if self.drop is not None else None

    def forward(self, vit_out, frames, return_all=False):
        """vit_out fp32 [F*257, 1408] -> (last_hidden fp32 [F*32, 768], fp16 copy, image_embeds fp16)."""
        d = self.d
        Hq, nq, T, heads = d.qf_hidden, d.num_query, d.vit_tokens, d.qf_heads
        hd = Hq // heads
        M = frames * nq
        ie16 = torch.empty((frames * T, d.vit_width), dtype=H16, device="cuda")
        ie32 = torch.empty((frames * T, d.vit_width), dtype=torch.float32, device="cuda") if return_all else None
        ops.norm(vit_out, self.lnv_w, self.lnv_b, 1e-5, 0, out_h=ie16, out_f32=ie32)
        ev = self.xattn_events
        if ev is not None:
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
        kv = ops.gemm(ie16, self.kv_w, bias=self.kv_b)                     # [F*257, n_cross*1536] fp16
        if ev is not None:
            e1 = torch.cuda.Event(enable_timing=True); e1.record(); ev.append((e0, e1))
        kv_rs = self.n_cross * 2 * Hq
        # BertEmbeddings on the query tokens is frame independent (Qformer.py:104-108): LN once, broadcast
        q0 = torch.empty((nq, Hq), dtype=torch.float32, device="cuda")
        ops.norm(self.query_tokens, self.emb_ln[0], self.emb_ln[1], d.qf_ln_eps, 0, out_f32=q0)
        h = q0.unsqueeze(0).expand(frames, nq, Hq).reshape(M, Hq).contiguous()
        if self.drop is not None and self.drop.qformer > 0.0:        # every frame draws its own mask over the shared embeddings
            ops.dropout(h, h, M, Hq, self.drop.word, dr.site(dr.QF, 0, dr.EMB), self.drop.qformer)
        h16 = torch.empty((M, Hq), dtype=H16, device="cuda")
        ops.cast_to(h, h16)
        qkv = torch.empty((M, 3 * Hq), dtype=H16, device="cuda")
        ctx = torch.empty((M, Hq), dtype=H16, device="cuda")
        qc = torch.empty((M, Hq), dtype=H16, device="cuda")
        inter = torch.empty((M, d.qf_inter), dtype=H16, device="cuda")
        outs = [h.clone()] if return_all else None
        for li, L in enumerate(self.layers):
            ops.gemm(h16, L["self_qkv_w"], out=qkv, bias=L["self_qkv_b"])
            rs = 3 * Hq
            ops.attention_fwd(qkv, qkv[:, Hq:], qkv[:, 2 * Hq:], ctx, frames, heads, nq, nq, hd, hd ** -0.5,
                              (nq * rs, rs), (nq * rs, rs), (nq * rs, rs), (nq * Hq, Hq), drop=self._pdrop(dr.site(dr.QF, li, dr.SELF_P)))
            self._attn_block(ctx, L["self_o"], L["self_ln"], h, h16, M, dr.site(dr.QF, li, dr.SELF_RES))
            c = L["cross"]
            if c is not None:
                ops.gemm(h16, c["q"][0], out=qc, bias=c["q"][1])
                kbase = kv[:, c["idx"] * 2 * Hq:]
                if ev is not None:
                    e0 = torch.cuda.Event(enable_timing=True); e0.record()
                ops.attention_fwd(qc, kbase, kbase[:, Hq:], ctx, frames, heads, nq, T, hd, hd ** -0.5,
                                  (nq * Hq, Hq), (T * kv_rs, kv_rs), (T * kv_rs, kv_rs), (nq * Hq, Hq),
                                  drop=self._pdrop(dr.site(dr.QF, li, dr.CROSS_P)))
                if ev is not None:
                    e1 = torch.cuda.Event(enable_timing=True); e1.record(); ev.append((e0, e1))
                self._attn_block(ctx, c["o"], c["ln"], h, h16, M, dr.site(dr.QF, li, dr.CROSS_RES))
            ops.gemm(h16, L["ffn_i"][0], out=inter, bias=L["ffn_i"][1], gelu=True)
            self._attn_block(inter, L["ffn_o"], L["ffn_ln"], h, h16, M, dr.site(dr.QF, li, dr.FF_RES))
            if return_all:
                outs.append(h.clone())
        if return_all:
            return h, h16, ie32, outs
        return h, h16

    def time_xattn_path(self, frames, train=True, reps=10):
        """The kernels of the cross-attention path ALONE -- the batched K/V projection GEMM of the 6 cross-attention layers and the
        6 attention cores, same shapes, strides and dropout sites as forward() -- back to back in one CUDA graph, replayed `reps`
        times between two events: ms per pass.  (CUDA events around the eager calls of forward() also count the launch gap in
        front of every small kernel, the host being the slower side there.)  Inputs are random: the kernels' time does not
        depend on the values."""
        from .dropout import DropState
        d = self.d
        Hq, nq, T, heads = d.qf_hidden, d.num_query, d.vit_tokens, d.qf_heads
        hd, M = Hq // heads, frames * nq
        g = torch.Generator(device="cuda").manual_seed(0)
        ie16 = torch.randn((frames * T, d.vit_width), device="cuda", generator=g).to(H16)
        qc = (torch.randn((M, Hq), device="cuda", generator=g) * 0.5).to(H16)
        ctx = torch.empty((M, Hq), dtype=H16, device="cuda")
        kv = torch.empty((frames * T, self.n_cross * 2 * Hq), dtype=H16, device="cuda")
        kv_rs = kv.shape[1]
        drop = DropState() if train else None
        cross = [(li, L["cross"]) for li, L in enumerate(self.layers) if L["cross"] is not None]

        def path():
            ops.gemm(ie16, self.kv_w, out=kv, bias=self.kv_b)
            for li, c in cross:
                kbase = kv[:, c["idx"] * 2 * Hq:]
                ops.attention_fwd(qc, kbase, kbase[:, Hq:], ctx, frames, heads, nq, T, hd, hd ** -0.5,
                                  (nq * Hq, Hq), (T * kv_rs, kv_rs), (T * kv_rs, kv_rs), (nq * Hq, Hq),
                                  drop=drop.attn(dr.site(dr.QF, li, dr.CROSS_P), drop.qformer) if drop is not None else None)

        path()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, capture_error_mode="thread_local"):
            path()
        for _ in range(2):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def project(self, h16):
        """t5_proj (blip2_mr.py:491): fp32 [F*32, 2048]."""
        return ops.gemm(h16, self.proj_w16, bias=self.proj_b, out_dtype=torch.float32)
