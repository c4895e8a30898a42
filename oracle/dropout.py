"""Train-mode dropout of the oracle.  Test infrastructure only.

The reference draws its masks from torch's RNG (nn.Dropout / F.dropout: modeling_t5.py:309,327,346,600,652,690,1149,1258;
Qformer.py:107,258,287,373; peft lora_dropout=0.05, blip2_mr.py:197).  That stream cannot be reproduced by any other
implementation, so train-mode parity is defined on a counter-hash mask that is a pure function of (step seed, site, row, column);
this file restates mr_blip_b200/csrc/dropmask.cuh in numpy (tests/test_host_logic.py compiles that header into a host harness
and compares the two bit for bit).  Where the masks are applied -- which tensors, before/after which op, scaled by 1/(1-p) --
follows the reference lines above; `Dropper.torch_rng=True` swaps the hash for torch's own F.dropout at the same places, which is
how the placement is pinned against the reference modules run in train mode (tests/test_oracle_golden.py).

Sites (one id per nn.Dropout call of a step):  site = stack << 12 | layer << 5 | slot
"""
import numpy as np
import torch

ENC, DEC, QF, HEAD = 0, 1, 2, 3
# T5 block / stack slots (modeling_t5.py line of the dropout call)
EMB, SELF_P, SELF_RES, CROSS_P, CROSS_RES, FF_INNER, FF_RES, FINAL = 0, 1, 2, 3, 4, 5, 6, 7
#        :1149  :600    :652      :600     :690       :327      :346    :1258
# LoRA input dropout, one per adapted Linear (peft lora.Linear.forward: lora_B(lora_A(dropout(x))))
LORA_SLOT = {"SelfAttention.q": 8, "SelfAttention.k": 9, "SelfAttention.v": 10, "SelfAttention.o": 11,
             "EncDecAttention.q": 12, "EncDecAttention.k": 13, "EncDecAttention.v": 14, "EncDecAttention.o": 15,
             "DenseReluDense.wi_0": 16, "DenseReluDense.wi_1": 17, "DenseReluDense.wo": 18, "lm_head": 19}
# Q-Former slots: EMB (Qformer.py:107), SELF_P / CROSS_P (:258), SELF_RES / CROSS_RES (BertSelfOutput :287), FF_RES (BertOutput :373)

M32 = np.uint64(0xFFFFFFFF)


def site(stack, layer, slot):
    return (stack << 12) | (layer << 5) | slot


def lora_site(name):
    """'...encoder.block.3.layer.0.SelfAttention.q' -> site id; '...lm_head' -> HEAD."""
    if name.endswith("lm_head"):
        return site(HEAD, 0, LORA_SLOT["lm_head"])
    parts = name.split(".")
    i = parts.index("block")
    stack = ENC if parts[i - 1] == "encoder" else DEC
    return site(stack, int(parts[i + 1]), LORA_SLOT[".".join(parts[-2:])])


def mix(x):
    """lowbias32 on uint32 arrays (computed in uint64, masked)."""
    x = np.asarray(x, dtype=np.uint64) & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x21f0aaad)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x735a2d97)) & M32
    x ^= x >> np.uint64(15)
    return x


def key(seed, site_id):
    return mix((np.uint64(seed) + np.uint64(0x9E3779B9) * np.uint64(site_id + 1)) & M32)


def thr_of(p):
    return min(255, max(0, int(256.0 * p + 0.5)))


def scale_of(p):
    return np.float32(256.0) / np.float32(256 - thr_of(p))


def draws(seed, site_id, rows, cols):
    """uint8 [rows, cols]: the 8-bit draw of every element of a [rows, cols] site."""
    ng = (cols + 3) // 4
    r = np.arange(rows, dtype=np.uint64)[:, None]
    g = np.arange(ng, dtype=np.uint64)[None, :]
    w = mix(((r * np.uint64(ng) + g) & M32) ^ key(seed, site_id))                    # [rows, ng]
    b = (w[:, :, None] >> (np.uint64(8) * np.arange(4, dtype=np.uint64))[None, None, :]) & np.uint64(0xFF)
    return b.reshape(rows, ng * 4)[:, :cols].astype(np.uint8)


def keep_mask(seed, site_id, rows, cols, p):
    return draws(seed, site_id, rows, cols) >= thr_of(p)


def _mix_t(x):
    """lowbias32 on int64 tensors holding uint32 values (products stay below 2^63)."""
    m = 0xFFFFFFFF
    x = x & m
    x = x ^ (x >> 16)
    x = (x * 0x21f0aaad) & m
    x = x ^ (x >> 15)
    x = (x * 0x735a2d97) & m
    return x ^ (x >> 15)


def keep_mask_torch(seed, site_id, rows, cols, p, device):
    """keep_mask() evaluated with torch integer ops on `device` (the full-depth GPU parity runs need masks of 10^8+ elements);
    bit-identical to the numpy restatement (tests/test_oracle_golden.py::test_dropout_mask_torch_matches_numpy)."""
    ng = (cols + 3) // 4
    k = int(key(seed, site_id))
    r = torch.arange(rows, dtype=torch.int64, device=device)[:, None]
    g = torch.arange(ng, dtype=torch.int64, device=device)[None, :]
    w = _mix_t(((r * ng + g) & 0xFFFFFFFF) ^ k)
    b = (w[:, :, None] >> (8 * torch.arange(4, dtype=torch.int64, device=device))[None, None, :]) & 0xFF
    return b.reshape(rows, ng * 4)[:, :cols] >= thr_of(p)


class Dropper:
    """drop(x, site_id, p): x [..., cols] -> x * keep * scale with rows = the flattened leading dimensions."""

    def __init__(self, seed, t5=0.1, lora=0.05, qformer=0.1, torch_rng=False):
        self.seed, self.t5, self.lora, self.qformer, self.torch_rng = int(seed) & 0xFFFFFFFF, t5, lora, qformer, torch_rng

    def __call__(self, x, site_id, p):
        if p <= 0.0:
            return x
        if self.torch_rng:
            return torch.nn.functional.dropout(x, p=p, training=True)
        cols = x.shape[-1]
        rows = x.numel() // cols
        if x.is_cuda:
            m = keep_mask_torch(self.seed, site_id, rows, cols, p, x.device).view(x.shape)
        else:
            m = torch.from_numpy(keep_mask(self.seed, site_id, rows, cols, p)).view(x.shape)
        return x * (m.to(x.dtype) * float(scale_of(p)))
