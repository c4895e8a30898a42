"""Golden vectors that pin WHERE the train-mode dropout masks are applied: the reference's own Q-Former and T5
(Qformer.py / modeling_t5.py, unmodified, through ref_shim.py) run in .train() mode right after torch.manual_seed(S); the
oracle with Dropper(torch_rng=True) calls torch's F.dropout at its own dropout sites, so it consumes the same RNG stream and
reproduces these outputs only if every mask sits at the reference's place, on a tensor of the reference's shape, in the
reference's order (tests/test_oracle_golden.py::test_train_mode_dropout_placement).  The LoRA input dropout is peft's
(third-party, absent: parity unpinned) and is switched off on both sides here (merged weights, Dropper(lora=0)).

Run here:   python tests/golden/make_golden_train_mode.py      -> tests/golden/train_mode_tiny.npz
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402
from make_golden import SEED, build_ref_qformer, build_ref_t5  # noqa: E402
from mr_blip_b200.dims import TINY, init_state_dict  # noqa: E402

RNG_SEED = 99


def inputs(d):
    g = torch.Generator().manual_seed(21)
    image_embeds = torch.randn(3, d.vit_tokens, d.vit_width, generator=g)
    emb = torch.randn(2, 40, d.d_model, generator=g) * 2.0
    mask = torch.ones(2, 40, dtype=torch.long)
    mask[1, 33:] = 0
    labels = torch.randint(2, 1000, (2, 7), generator=g)
    labels[:, -1] = 1
    labels[1, 5:] = -100
    labels[1, 4] = 1
    return image_embeds, emb, mask, labels


def main():
    torch.set_num_threads(1)                       # the bernoulli fill of large tensors is split over threads
    _, qf, t5 = ref_shim.load_reference_modules()
    d = TINY
    sd = init_state_dict(d, seed=SEED, lora_b_std=0.02)
    image_embeds, emb, mask, labels = inputs(d)
    qformer = build_ref_qformer(qf, d, sd).train()
    ref_t5 = build_ref_t5(t5, d, sd).train()
    with torch.no_grad():
        torch.manual_seed(RNG_SEED)
        q_out = qformer.bert(query_embeds=sd["query_tokens"].expand(3, -1, -1), encoder_hidden_states=image_embeds,
                             encoder_attention_mask=torch.ones(3, d.vit_tokens, dtype=torch.long),
                             return_dict=True).last_hidden_state
        torch.manual_seed(RNG_SEED)
        out = ref_t5(inputs_embeds=emb, attention_mask=mask, labels=labels, decoder_attention_mask=(labels != -100).long(),
                     return_dict=True)
    np.savez_compressed(os.path.join(HERE, "train_mode_tiny.npz"), rng_seed=np.int64(RNG_SEED),
                        qformer_out=q_out.numpy(), loss=np.float64(out.loss.item()),
                        logits_lse=torch.logsumexp(out.logits, -1).numpy(), logits_head=out.logits[:, :, :64].numpy(),
                        enc_out=out.encoder_last_hidden_state[:, ::4, ::8].numpy())
    print("train-mode loss", out.loss.item(), "qformer absmean", q_out.abs().mean().item())


if __name__ == "__main__":
    main()
