"""fp32 restatement of the query-only branch of lavis/models/blip2_models/Qformer.py as BLIP2_MR
runs it (blip2_mr.py:483-489: query_embeds + encoder_hidden_states, no text).
Test infrastructure only (see oracle/__init__.py)."""
import math

import torch
import torch.nn.functional as F

from . import dropout as D_


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd, d, name, x):
    return F.layer_norm(x, (d.qf_hidden,), sd[name + ".weight"], sd[name + ".bias"], d.qf_ln_eps)


def _heads(d, x):
    B, L, _ = x.shape
    return x.view(B, L, d.qf_heads, -1).permute(0, 2, 1, 3)      # transpose_for_scores, Qformer.py:160-166


def bert_attention(sd, d, name, hidden, kv_source, drop=None, p_site=None, res_site=None):
    """BertSelfAttention + BertSelfOutput (Qformer.py:169-289).  The additive masks are all zero on
    this path: query self-attention mask is all-ones (Qformer.py:881-886 with attention_mask=None)
    and image_atts is all-ones (blip2_mr.py:448-450), so they are omitted."""
    q = _heads(d, _lin(sd, name + ".self.query", hidden))
    k = _heads(d, _lin(sd, name + ".self.key", kv_source))
    v = _heads(d, _lin(sd, name + ".self.value", kv_source))
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d.qf_hidden // d.qf_heads)
    probs = scores.softmax(dim=-1)
    if drop is not None:
        probs = drop(probs, p_site, drop.qformer)                  # Qformer.py:258
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
    ctx = ctx.view(ctx.shape[0], ctx.shape[1], -1)
    out = _lin(sd, name + ".output.dense", ctx)
    if drop is not None:
        out = drop(out, res_site, drop.qformer)                    # BertSelfOutput, Qformer.py:287
    return _ln(sd, d, name + ".output.LayerNorm", out + hidden)


def qformer_forward(sd, d, image_embeds, prefix="Qformer.bert.", return_all=False, drop=None):
    """BertModel.forward (Qformer.py:804-965) -> BertEncoder (495-589) -> BertLayer (402-474); drop=None is eval mode,
    a Dropper (oracle/dropout.py) the train mode the frozen Q-Former runs in.
    image_embeds [BT,257,1408] -> last_hidden_state [BT,32,768]."""
    BT = image_embeds.shape[0]
    h = sd["query_tokens"].expand(BT, -1, -1)
    h = _ln(sd, d, prefix + "embeddings.LayerNorm", h)             # BertEmbeddings, Qformer.py:104-108
    if drop is not None:
        h = drop(h, D_.site(D_.QF, 0, D_.EMB), drop.qformer)
    outs = [h]
    for i in range(d.qf_layers):
        b = f"{prefix}encoder.layer.{i}."
        h = bert_attention(sd, d, b + "attention", h, h, drop, D_.site(D_.QF, i, D_.SELF_P), D_.site(D_.QF, i, D_.SELF_RES))
        if i % d.qf_cross_freq == 0:                               # Qformer.py:386-389
            h = bert_attention(sd, d, b + "crossattention", h, image_embeds, drop, D_.site(D_.QF, i, D_.CROSS_P),
                               D_.site(D_.QF, i, D_.CROSS_RES))
        inter = F.gelu(_lin(sd, b + "intermediate_query.dense", h))   # hidden_act "gelu" (bert-base)
        ff = _lin(sd, b + "output_query.dense", inter)
        if drop is not None:
            ff = drop(ff, D_.site(D_.QF, i, D_.FF_RES), drop.qformer)   # BertOutput, Qformer.py:373
        h = _ln(sd, d, b + "output_query.LayerNorm", ff + h)
        outs.append(h)
    return outs if return_all else h
