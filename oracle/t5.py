"""fp32 restatement of lavis/models/blip2_models/modeling_t5.py (T5ForConditionalGeneration as
BLIP2_MR drives it) with peft-style LoRA on every Linear.  Test infrastructure only.

LoRA (peft==0.13.0 `lora.Linear.forward`, third-party, not vendored -- parity unpinned):
    y = base(x) + lora_B(lora_A(dropout(x))) * (alpha / r)        [eval: dropout = identity]

Every function takes an optional `drop` (oracle/dropout.py Dropper): None = eval mode; otherwise the train-mode dropout of
modeling_t5.py:327,346,600,652,690,1149,1258 and of the LoRA inputs is applied at the reference's places.
"""
import math

import torch
import torch.nn.functional as F

from . import dropout as D_

PREFIX = "t5_model.base_model.model."


def lora_linear(sd, d, name, x, drop=None):
    w = sd[name + ".base_layer.weight"] if (name + ".base_layer.weight") in sd else sd[name + ".weight"]
    y = F.linear(x, w)
    a = sd.get(name + ".lora_A.default.weight")
    if a is not None:
        b = sd[name + ".lora_B.default.weight"]
        xa = drop(x, D_.lora_site(name), drop.lora) if drop is not None else x
        y = y + F.linear(F.linear(xa, a), b) * (d.lora_alpha / d.lora_r)
    return y


def rmsnorm(x, w, eps):
    """T5LayerNorm, modeling_t5.py:263-277."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def relative_position_bucket(rel, bidirectional, num_buckets, max_distance):
    """modeling_t5.py:393-445."""
    buckets = torch.zeros_like(rel)
    if bidirectional:
        num_buckets //= 2
        buckets = buckets + (rel > 0).to(torch.long) * num_buckets
        rel = torch.abs(rel)
    else:
        rel = -torch.min(rel, torch.zeros_like(rel))
    max_exact = num_buckets // 2
    is_small = rel < max_exact
    large = max_exact + (
        torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rel, large)


def compute_bias(table, q_len, k_len, bidirectional, d):
    """modeling_t5.py:447-472 -> [1, H, q_len, k_len]."""
    ctx = torch.arange(q_len, dtype=torch.long, device=table.device)[:, None]
    mem = torch.arange(k_len, dtype=torch.long, device=table.device)[None, :]
    bucket = relative_position_bucket(mem - ctx, bidirectional, d.rel_buckets, d.rel_max_dist)
    return table[bucket].permute(2, 0, 1).unsqueeze(0)


def t5_attention(sd, d, name, hidden, kv, position_bias, drop=None, p_site=None):
    """T5Attention.forward, modeling_t5.py:474-620: no 1/sqrt(d) scaling, additive bias(+mask),
    fp32 softmax."""
    B, L, _ = hidden.shape
    H = d.t5_heads

    def shape(x):
        return x.view(B, -1, H, d.d_kv).transpose(1, 2)

    q = shape(lora_linear(sd, d, name + ".q", hidden, drop))
    k = shape(lora_linear(sd, d, name + ".k", kv, drop))
    v = shape(lora_linear(sd, d, name + ".v", kv, drop))
    scores = torch.matmul(q, k.transpose(3, 2))
    if scores.dtype == position_bias.dtype:
        scores = scores + position_bias
    else:                                                     # under autocast the reference adds in place, in the scores' dtype
        scores += position_bias                               # (modeling_t5.py:597: `scores += position_bias_masked`)
    w = F.softmax(scores.float(), dim=-1).type_as(scores)
    if drop is not None:
        w = drop(w, p_site, drop.t5)                          # modeling_t5.py:600
    out = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, -1, H * d.d_kv)
    return lora_linear(sd, d, name + ".o", out, drop)


def t5_ff(sd, d, name, x, drop=None, inner_site=None):
    """T5DenseGatedActDense with dense_act_fn='gelu' (exact erf GELU; blip2_mr.py:145),
    modeling_t5.py:314-329."""
    g = F.gelu(lora_linear(sd, d, name + ".wi_0", x, drop))
    lin = lora_linear(sd, d, name + ".wi_1", x, drop)
    h = g * lin
    if drop is not None:
        h = drop(h, inner_site, drop.t5)                      # modeling_t5.py:327
    return lora_linear(sd, d, name + ".wo", h, drop)


def extended_mask(mask, dtype=torch.float32):
    """get_extended_attention_mask as used at modeling_t5.py:1114-1116: (1 - m) * finfo.min."""
    return (1.0 - mask[:, None, None, :].to(dtype)) * torch.finfo(dtype).min


def _res(h, branch, drop, stack, layer, slot):
    """hidden + dropout(branch): modeling_t5.py:346,652,690."""
    return h + (drop(branch, D_.site(stack, layer, slot), drop.t5) if drop is not None else branch)


def t5_encoder(sd, d, inputs_embeds, attention_mask, prefix=PREFIX, return_all=False, drop=None):
    """T5Stack.forward (encoder), modeling_t5.py:1021-1282."""
    h = inputs_embeds
    if drop is not None:
        h = drop(h, D_.site(D_.ENC, 0, D_.EMB), drop.t5)      # modeling_t5.py:1149
    L = h.shape[1]
    table = sd[prefix + "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    bias = compute_bias(table, L, L, True, d) + extended_mask(attention_mask)
    outs = [h]
    for i in range(d.t5_layers):
        b = f"{prefix}encoder.block.{i}."
        n = rmsnorm(h, sd[b + "layer.0.layer_norm.weight"], d.t5_ln_eps)
        a = t5_attention(sd, d, b + "layer.0.SelfAttention", n, n, bias, drop, D_.site(D_.ENC, i, D_.SELF_P))
        h = _res(h, a, drop, D_.ENC, i, D_.SELF_RES)
        n = rmsnorm(h, sd[b + "layer.1.layer_norm.weight"], d.t5_ln_eps)
        f = t5_ff(sd, d, b + "layer.1.DenseReluDense", n, drop, D_.site(D_.ENC, i, D_.FF_INNER))
        h = _res(h, f, drop, D_.ENC, i, D_.FF_RES)
        outs.append(h)
    h = rmsnorm(h, sd[prefix + "encoder.final_layer_norm.weight"], d.t5_ln_eps)
    if drop is not None:
        h = drop(h, D_.site(D_.ENC, 0, D_.FINAL), drop.t5)    # modeling_t5.py:1258
    return (h, outs) if return_all else h


def shift_right(labels, start_id=0, pad_id=0):
    """modeling_t5.py:919-948."""
    s = labels.new_zeros(labels.shape)
    s[..., 1:] = labels[..., :-1].clone()
    s[..., 0] = start_id
    return s.masked_fill(s == -100, pad_id)


def t5_decoder(sd, d, decoder_input_ids, enc_out, enc_mask, decoder_attention_mask=None, prefix=PREFIX, drop=None):
    """T5Stack.forward (decoder, no cache): causal self-attention with unidirectional buckets,
    cross-attention with zero position bias + encoder padding mask (modeling_t5.py:575-598)."""
    h = sd[prefix + "shared.weight"][decoder_input_ids]
    if drop is not None:
        h = drop(h, D_.site(D_.DEC, 0, D_.EMB), drop.t5)
    B, L = decoder_input_ids.shape
    if decoder_attention_mask is None:
        decoder_attention_mask = torch.ones(B, L, device=h.device)
    table = sd[prefix + "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    causal = torch.tril(torch.ones(L, L, device=h.device))[None, None] * decoder_attention_mask[:, None, None, :].float()
    self_bias = compute_bias(table, L, L, False, d) + (1.0 - causal) * torch.finfo(torch.float32).min
    cross_bias = extended_mask(enc_mask)
    for i in range(d.t5_dec_layers):
        b = f"{prefix}decoder.block.{i}."
        n = rmsnorm(h, sd[b + "layer.0.layer_norm.weight"], d.t5_ln_eps)
        a = t5_attention(sd, d, b + "layer.0.SelfAttention", n, n, self_bias, drop, D_.site(D_.DEC, i, D_.SELF_P))
        h = _res(h, a, drop, D_.DEC, i, D_.SELF_RES)
        n = rmsnorm(h, sd[b + "layer.1.layer_norm.weight"], d.t5_ln_eps)
        a = t5_attention(sd, d, b + "layer.1.EncDecAttention", n, enc_out, cross_bias, drop, D_.site(D_.DEC, i, D_.CROSS_P))
        h = _res(h, a, drop, D_.DEC, i, D_.CROSS_RES)
        n = rmsnorm(h, sd[b + "layer.2.layer_norm.weight"], d.t5_ln_eps)
        f = t5_ff(sd, d, b + "layer.2.DenseReluDense", n, drop, D_.site(D_.DEC, i, D_.FF_INNER))
        h = _res(h, f, drop, D_.DEC, i, D_.FF_RES)
    h = rmsnorm(h, sd[prefix + "decoder.final_layer_norm.weight"], d.t5_ln_eps)
    if drop is not None:
        h = drop(h, D_.site(D_.DEC, 0, D_.FINAL), drop.t5)
    return h


def t5_logits(sd, d, dec_out, prefix=PREFIX, drop=None):
    """modeling_t5.py:1862-1870: no d_model**-0.5 rescale because FlanT5 unties lm_head."""
    return lora_linear(sd, d, prefix + "lm_head", dec_out, drop)


def t5_forward(sd, d, inputs_embeds, attention_mask, labels, decoder_attention_mask=None, prefix=PREFIX, drop=None):
    """T5ForConditionalGeneration.forward with labels, modeling_t5.py:1734-1893.
    -> dict(loss, logits, encoder_last_hidden_state)."""
    enc = t5_encoder(sd, d, inputs_embeds, attention_mask, prefix, drop=drop)
    dec_in = shift_right(labels)
    dec = t5_decoder(sd, d, dec_in, enc, attention_mask, decoder_attention_mask, prefix, drop=drop)
    logits = t5_logits(sd, d, dec, prefix, drop=drop)
    loss = F.cross_entropy(logits.view(-1, logits.size(-1)), labels.view(-1), ignore_index=-100)
    return {"loss": loss, "logits": logits, "encoder_last_hidden_state": enc}
