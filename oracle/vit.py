"""fp32 restatement of lavis/models/eva_vit.py (VisionTransformer.forward_features) and of
blip2.py:113-119 (ln_vision).  Test infrastructure only (see oracle/__init__.py)."""
import torch
import torch.nn.functional as F


def patch_embed(sd, d, image, prefix="visual_encoder."):
    """eva_vit.py:198-204 Conv2d(3, W, k=14, s=14) -> flatten(2).transpose(1, 2); then cls concat and
    +pos_embed (eva_vit.py:325-332)."""
    x = F.conv2d(image, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"],
                 stride=d.patch)
    x = x.flatten(2).transpose(1, 2)
    cls = sd[prefix + "cls_token"].expand(x.shape[0], -1, -1)
    x = torch.cat((cls, x), dim=1)
    return x + sd[prefix + "pos_embed"]


def attention(sd, d, x, b):
    """eva_vit.py:118-148: qkv Linear with bias (q_bias, 0, v_bias); q * hd**-0.5; softmax(q k^T) v; proj."""
    B, N, C = x.shape
    qkv_bias = torch.cat((sd[b + "attn.q_bias"], torch.zeros_like(sd[b + "attn.v_bias"]), sd[b + "attn.v_bias"]))
    qkv = F.linear(x, sd[b + "attn.qkv.weight"], qkv_bias)
    qkv = qkv.reshape(B, N, 3, d.vit_heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (d.vit_head_dim ** -0.5)
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    return F.linear(x, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])


def mlp(sd, x, b):
    """eva_vit.py:54-61: fc1 -> exact (erf) GELU -> fc2."""
    x = F.gelu(F.linear(x, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]))
    return F.linear(x, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])


def block(sd, d, x, b):
    """eva_vit.py:173-176 (gamma_1 is None: init_values=None at eva_vit.py:416-428)."""
    W = (d.vit_width,)
    x = x + attention(sd, d, F.layer_norm(x, W, sd[b + "norm1.weight"], sd[b + "norm1.bias"], d.vit_ln_eps), b)
    x = x + mlp(sd, F.layer_norm(x, W, sd[b + "norm2.weight"], sd[b + "norm2.bias"], d.vit_ln_eps), b)
    return x


def vit_forward(sd, d, image, prefix="visual_encoder.", return_all=False):
    """eva_vit.py:324-340 forward_features. image [BT,3,224,224] fp32 -> [BT,257,1408]."""
    x = patch_embed(sd, d, image, prefix)
    outs = [x]
    for i in range(d.vit_depth):
        x = block(sd, d, x, f"{prefix}blocks.{i}.")
        outs.append(x)
    return outs if return_all else x


def ln_vision(sd, d, x):
    """blip2.py:113-119 LayerNorm (fp32 compute), default eps 1e-5."""
    return F.layer_norm(x.float(), (d.vit_width,), sd["ln_vision.weight"], sd["ln_vision.bias"], 1e-5)
