"""Shapes of the BLIP2_MR hot path (EVA ViT-g -> Q-Former -> FlanT5-XL) and seeded synthetic
weights in the reference's state-dict key names.

Reference shapes: ViT-g  lavis/models/eva_vit.py:415-428; Q-Former  blip2_models/blip2.py:46-61
(bert-base + cross-attention every 2nd layer, encoder_width 1408, 32 queries); FlanT5-XL public
config (d_model 2048, d_kv 64, 32 heads, d_ff 5120, 24+24 layers, gated-gelu, 32 buckets / 128,
vocab 32128, untied lm_head); LoRA r=8 alpha=8 on every T5 Linear incl. lm_head
(blip2_mr_models/blip2_mr.py:183-200).

Initialisers restate the reference's own (eva_vit.py:300-315, Qformer.py:664-674,
modeling_t5.py:853-913, peft LoRA: A kaiming-uniform(a=sqrt 5), B zeros).  No network access
exists in the build/bench environment, so every parity and throughput run uses these seeded
weights; real checkpoints load through BLIP2_MR.load_checkpoint when present.
"""
import math
from dataclasses import dataclass, asdict, replace

import torch


@dataclass(frozen=True)
class Dims:
    # EVA ViT-g/14
    img_size: int = 224
    patch: int = 14
    vit_width: int = 1408
    vit_depth: int = 39
    vit_heads: int = 16
    vit_mlp: int = 6144          # int(1408 * 4.3637)
    vit_ln_eps: float = 1e-6
    # Q-Former
    num_query: int = 32
    qf_hidden: int = 768
    qf_layers: int = 12
    qf_heads: int = 12
    qf_inter: int = 3072
    qf_cross_freq: int = 2
    qf_ln_eps: float = 1e-12
    # FlanT5-XL
    d_model: int = 2048
    d_kv: int = 64
    t5_heads: int = 32
    d_ff: int = 5120
    t5_layers: int = 24
    t5_dec_layers: int = 24
    vocab: int = 32128
    rel_buckets: int = 32
    rel_max_dist: int = 128
    t5_ln_eps: float = 1e-6
    lora_r: int = 8
    lora_alpha: int = 8

    @property
    def n_patches(self):
        return (self.img_size // self.patch) ** 2

    @property
    def vit_tokens(self):
        return self.n_patches + 1

    @property
    def vit_head_dim(self):
        return self.vit_width // self.vit_heads

    def as_dict(self):
        return asdict(self)


FULL = Dims()
# true widths, shallow stacks: what the parity tests and golden vectors use
TINY = replace(FULL, vit_depth=2, qf_layers=2, t5_layers=2, t5_dec_layers=2)

T5_PREFIX = "t5_model.base_model.model."
LORA_TARGETS = ("q", "k", "v", "o", "wi_0", "wi_1", "wo", "lm_head")


def _trunc_normal(shape, std, gen):
    t = torch.empty(shape, dtype=torch.float32, device=gen.device)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0, generator=gen)
    return t


def _normal(shape, std, gen):
    return torch.empty(shape, dtype=torch.float32, device=gen.device).normal_(0.0, std, generator=gen)


def _rand(shape, gen):
    return torch.rand(shape, generator=gen, device=gen.device)


def _ones(n, gen):
    return torch.ones(n, device=gen.device)


def _zeros(*shape, gen):
    return torch.zeros(*shape, device=gen.device)


def init_vit(d: Dims, gen, sd, prefix="visual_encoder."):
    """eva_vit.py:252-315 (trunc_normal .02, zero bias, LN 1/0, proj/fc2 rescale by sqrt(2*layer))."""
    W, P = d.vit_width, d.patch
    sd[prefix + "cls_token"] = _trunc_normal((1, 1, W), 0.02, gen)
    sd[prefix + "pos_embed"] = _trunc_normal((1, d.vit_tokens, W), 0.02, gen)
    fan_in = 3 * P * P
    bound = 1.0 / math.sqrt(fan_in)  # nn.Conv2d default init (kaiming_uniform a=sqrt5)
    sd[prefix + "patch_embed.proj.weight"] = (_rand((W, 3, P, P), gen) * 2 - 1) * bound
    sd[prefix + "patch_embed.proj.bias"] = (_rand((W,), gen) * 2 - 1) * bound
    for i in range(d.vit_depth):
        b = f"{prefix}blocks.{i}."
        sd[b + "norm1.weight"] = _ones(W, gen)
        sd[b + "norm1.bias"] = _zeros(W, gen=gen)
        sd[b + "attn.q_bias"] = _normal((W,), 0.02, gen)   # zeros in the reference init; random here so
        sd[b + "attn.v_bias"] = _normal((W,), 0.02, gen)   # the bias path is exercised
        sd[b + "attn.qkv.weight"] = _trunc_normal((3 * W, W), 0.02, gen)
        sd[b + "attn.proj.weight"] = _trunc_normal((W, W), 0.02, gen) / math.sqrt(2.0 * (i + 1))
        sd[b + "attn.proj.bias"] = _normal((W,), 0.02, gen)
        sd[b + "norm2.weight"] = _ones(W, gen)
        sd[b + "norm2.bias"] = _zeros(W, gen=gen)
        sd[b + "mlp.fc1.weight"] = _trunc_normal((d.vit_mlp, W), 0.02, gen)
        sd[b + "mlp.fc1.bias"] = _normal((d.vit_mlp,), 0.02, gen)
        sd[b + "mlp.fc2.weight"] = _trunc_normal((W, d.vit_mlp), 0.02, gen) / math.sqrt(2.0 * (i + 1))
        sd[b + "mlp.fc2.bias"] = _normal((W,), 0.02, gen)


def qf_has_cross(d: Dims, layer):
    return layer % d.qf_cross_freq == 0   # Qformer.py:386-389


def init_qformer(d: Dims, gen, sd, prefix="Qformer.bert."):
    """Qformer.py:664-674: Linear/Embedding normal(0, .02), LN 1/0; query_tokens normal(0, .02)
    (blip2.py:57-60); ln_vision is blip2.py:113-119's LayerNorm(1408)."""
    H, E = d.qf_hidden, d.vit_width
    sd["ln_vision.weight"] = _ones(E, gen) + _normal((E,), 0.02, gen)
    sd["ln_vision.bias"] = _normal((E,), 0.02, gen)
    sd["query_tokens"] = _normal((1, d.num_query, H), 0.02, gen)
    sd[prefix + "embeddings.LayerNorm.weight"] = _ones(H, gen)
    sd[prefix + "embeddings.LayerNorm.bias"] = _zeros(H, gen=gen)

    def lin(name, out_f, in_f):
        sd[name + ".weight"] = _normal((out_f, in_f), 0.02, gen)
        sd[name + ".bias"] = _normal((out_f,), 0.02, gen)

    def ln(name):
        sd[name + ".weight"] = _ones(H, gen)
        sd[name + ".bias"] = _zeros(H, gen=gen)

    for i in range(d.qf_layers):
        b = f"{prefix}encoder.layer.{i}."
        for w in ("query", "key", "value"):
            lin(b + "attention.self." + w, H, H)
        lin(b + "attention.output.dense", H, H)
        ln(b + "attention.output.LayerNorm")
        if qf_has_cross(d, i):
            lin(b + "crossattention.self.query", H, H)
            lin(b + "crossattention.self.key", H, E)
            lin(b + "crossattention.self.value", H, E)
            lin(b + "crossattention.output.dense", H, H)
            ln(b + "crossattention.output.LayerNorm")
        lin(b + "intermediate_query.dense", d.qf_inter, H)
        lin(b + "output_query.dense", H, d.qf_inter)
        ln(b + "output_query.LayerNorm")


def init_t5(d: Dims, gen, sd, prefix=T5_PREFIX, lora_b_std=0.0, lm_head_std=None):
    """modeling_t5.py:853-913 (Mesh-TF init, factor 1.0) + peft-style LoRA adapters.
    lm_head_std defaults to d_model**-0.5 instead of the reference's 1.0 so that random-init logits
    have O(1) scale (a softmax over N(0, 45^2) logits is degenerate and hides errors)."""
    D, KV, H, F = d.d_model, d.d_kv, d.t5_heads, d.d_ff
    inner = KV * H
    r = d.lora_r
    if lm_head_std is None:
        lm_head_std = D ** -0.5

    def lora_linear(name, out_f, in_f, std):
        sd[name + ".base_layer.weight"] = _normal((out_f, in_f), std, gen)
        bound = 1.0 / math.sqrt(in_f)  # kaiming_uniform(a=sqrt 5) on [r, in_f]
        sd[name + ".lora_A.default.weight"] = (_rand((r, in_f), gen) * 2 - 1) * bound
        if lora_b_std > 0:
            sd[name + ".lora_B.default.weight"] = _normal((out_f, r), lora_b_std, gen)
        else:
            sd[name + ".lora_B.default.weight"] = _zeros(out_f, r, gen=gen)

    def attn(name, has_bias):
        lora_linear(name + ".q", inner, D, (D * KV) ** -0.5)
        lora_linear(name + ".k", inner, D, D ** -0.5)
        lora_linear(name + ".v", inner, D, D ** -0.5)
        lora_linear(name + ".o", D, inner, inner ** -0.5)
        if has_bias:
            sd[name + ".relative_attention_bias.weight"] = _normal((d.rel_buckets, H), 1.0, gen)

    def ff(name):
        lora_linear(name + ".wi_0", F, D, D ** -0.5)
        lora_linear(name + ".wi_1", F, D, D ** -0.5)
        lora_linear(name + ".wo", D, F, F ** -0.5)

    shared = _normal((d.vocab, D), 1.0, gen)
    sd[prefix + "shared.weight"] = shared
    sd[prefix + "encoder.embed_tokens.weight"] = shared     # tied storage, as in HF
    sd[prefix + "decoder.embed_tokens.weight"] = shared
    for i in range(d.t5_layers):
        b = f"{prefix}encoder.block.{i}."
        attn(b + "layer.0.SelfAttention", i == 0)
        sd[b + "layer.0.layer_norm.weight"] = _ones(D, gen) + _normal((D,), 0.02, gen)
        ff(b + "layer.1.DenseReluDense")
        sd[b + "layer.1.layer_norm.weight"] = _ones(D, gen) + _normal((D,), 0.02, gen)
    sd[prefix + "encoder.final_layer_norm.weight"] = _ones(D, gen)
    for i in range(d.t5_dec_layers):
        b = f"{prefix}decoder.block.{i}."
        attn(b + "layer.0.SelfAttention", i == 0)
        sd[b + "layer.0.layer_norm.weight"] = _ones(D, gen) + _normal((D,), 0.02, gen)
        attn(b + "layer.1.EncDecAttention", False)
        sd[b + "layer.1.layer_norm.weight"] = _ones(D, gen) + _normal((D,), 0.02, gen)
        ff(b + "layer.2.DenseReluDense")
        sd[b + "layer.2.layer_norm.weight"] = _ones(D, gen) + _normal((D,), 0.02, gen)
    sd[prefix + "decoder.final_layer_norm.weight"] = _ones(D, gen)
    lora_linear(prefix + "lm_head", d.vocab, D, lm_head_std)


def init_t5_proj(d: Dims, gen, sd):
    """nn.Linear(768, 2048) default init (blip2_mr.py:267-269)."""
    bound = 1.0 / math.sqrt(d.qf_hidden)
    sd["t5_proj.weight"] = (_rand((d.d_model, d.qf_hidden), gen) * 2 - 1) * bound
    sd["t5_proj.bias"] = (_rand((d.d_model,), gen) * 2 - 1) * bound


ANSWERER_PREFIX = "answerer_model.base_model.model."


def add_answerer(sd, d: Dims = FULL, seed: int = 1234, lora_b_std: float = 0.0, device="cpu"):
    """The QA branch's second LoRA T5 (blip2_mr.py:148-156,199-235: `answerer_model`, loaded from the SAME pretrained FlanT5 as
    the localizer and wrapped by its own peft adapters): frozen weights = the localizer's (shared storage), fresh LoRA A / B."""
    gen = torch.Generator(device=device).manual_seed(seed + 4)
    for k in [k for k in sd if k.startswith(T5_PREFIX)]:
        nk = ANSWERER_PREFIX + k[len(T5_PREFIX):]
        v = sd[k]
        if ".lora_A." in k:
            sd[nk] = (_rand(tuple(v.shape), gen) * 2 - 1) * (1.0 / math.sqrt(v.shape[1]))
        elif ".lora_B." in k:
            sd[nk] = _normal(tuple(v.shape), lora_b_std, gen) if lora_b_std > 0 else torch.zeros_like(v)
        else:
            sd[nk] = v
    return sd


def init_state_dict(d: Dims = FULL, seed: int = 1234, lora_b_std: float = 0.0, parts=("vit", "qformer", "t5"),
                    device="cpu"):
    """Seeded fp32 state dict with the reference's key names (SURVEY.md §5 checkpoint row).  device="cpu" is
    what parity tests and golden vectors use (bit-reproducible); device="cuda" draws the full-size 4 B
    parameters on the GPU in seconds for throughput runs (different random stream, same distributions)."""
    sd = {}

    def gen(k):
        return torch.Generator(device=device).manual_seed(seed + k)

    # one generator per part so that a part's weights do not depend on which other parts are built
    if "vit" in parts:
        init_vit(d, gen(0), sd)
    if "qformer" in parts:
        init_qformer(d, gen(1), sd)
        init_t5_proj(d, gen(2), sd)
    if "t5" in parts:
        init_t5(d, gen(3), sd, lora_b_std=lora_b_std)
    return sd
