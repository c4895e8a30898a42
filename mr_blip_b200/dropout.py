"""Train-mode dropout state of the product path (kernels: csrc/dropmask.cuh, csrc/dropout.cu and the attention kernels).

The reference's train() step draws dropout masks in the frozen Q-Former (0.1: Qformer.py:107,258,287,373), in T5 (0.1:
modeling_t5.py:327,346,600,652,690,1149,1258) and on every LoRA input (0.05: blip2_mr.py:197); the ViT stays in eval mode
(blip2_mr.py:136-137).  Here a mask is the pure function keep(step seed, site, row, column) of csrc/dropmask.cuh, recomputed by
the backward kernels; this module owns the two host-side pieces: the site numbering (one id per nn.Dropout call of a step) and the
step seed, ONE device word rewritten before every step so that a replayed CUDA graph draws fresh masks.
"""
import torch

ENC, DEC, QF, HEAD = 0, 1, 2, 3
EMB, SELF_P, SELF_RES, CROSS_P, CROSS_RES, FF_INNER, FF_RES, FINAL = range(8)
LORA_SLOT = {"SelfAttention.q": 8, "SelfAttention.k": 9, "SelfAttention.v": 10, "SelfAttention.o": 11,
             "EncDecAttention.q": 12, "EncDecAttention.k": 13, "EncDecAttention.v": 14, "EncDecAttention.o": 15,
             "DenseReluDense.wi_0": 16, "DenseReluDense.wi_1": 17, "DenseReluDense.wo": 18, "lm_head": 19}


def site(stack, layer, slot):
    return (stack << 12) | (layer << 5) | slot


def lora_site(name):
    """Site of the LoRA input dropout of the Linear called `name` ('...decoder.block.3.layer.1.EncDecAttention.k', '...lm_head')."""
    if name.endswith("lm_head"):
        return site(HEAD, 0, LORA_SLOT["lm_head"])
    parts = name.split(".")
    i = parts.index("block")
    return site(ENC if parts[i - 1] == "encoder" else DEC, int(parts[i + 1]), LORA_SLOT[".".join(parts[-2:])])


def _mix(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x21f0aaad) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x735a2d97) & 0xFFFFFFFF
    x ^= x >> 15
    return x


class DropState:
    """Probabilities + the device word that holds the step seed.  `advance()` before every training step (outside any graph
    capture); `set_seed(v)` pins the word for a parity test against oracle.dropout.Dropper(seed=v)."""

    def __init__(self, t5=0.1, lora=0.05, qformer=0.1, base_seed=0, attention=True, device="cuda"):
        self.t5, self.lora, self.qformer = float(t5), float(lora), float(qformer)
        self.attention = attention               # False: skip the attention-probability sites (A/B runs)
        self.base_seed, self.step = int(base_seed), 0
        self.word = torch.zeros(1, dtype=torch.int32, device=device)
        self.seed = 0
        self.set_seed(_mix(self.base_seed))

    def set_seed(self, v):
        self.seed = int(v) & 0xFFFFFFFF
        self.word.fill_(self.seed - (1 << 32) if self.seed >= (1 << 31) else self.seed)      # the kernels read it as uint32
        return self.seed

    def advance(self):
        self.step += 1
        return self.set_seed(_mix(self.base_seed ^ _mix(self.step)))

    def attn(self, site_id, p):
        """(seed word, site, p) for an attention kernel, or None when the attention sites are switched off."""
        return (self.word, site_id, p) if (self.attention and p > 0.0) else None
