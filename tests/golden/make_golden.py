"""Generate the committed golden vectors by running the REFERENCE's own modules
(/root/reference/lavis/models/{eva_vit.py, blip2_models/Qformer.py, blip2_models/modeling_t5.py,
blip2_mr_models/utils.py}, executed unmodified through ref_shim.py) on seeded synthetic inputs and
seeded weights (mr_blip_b200.dims.init_state_dict, TINY = true widths, 2-layer stacks).

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
Outputs: tests/golden/*.npz, tests/golden/mr_utils_golden.json.  The GPU box has no /root/reference;
tests there read only these files.

LoRA note: peft is absent, so the reference T5 runs with merged weights W + B.A (identical in exact
arithmetic to peft's base(x) + B(A(x)) at alpha/r = 1); LoRA gradients are derived from the
reference autograd's dense dL/dW as dB = G.A^T, dA = B^T.G.
"""
import json
import os
import sys
from functools import partial

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402
from mr_blip_b200.dims import TINY, T5_PREFIX, init_state_dict  # noqa: E402
from mr_blip_b200.tokenizer import SyntheticT5Tokenizer  # noqa: E402
from oracle import blip2_mr as ob, synth  # noqa: E402

SEED = 1234
VIT_TOKENS = [0, 1, 100, 256]
LORA_PROBES = ["encoder.block.0.layer.0.SelfAttention.q", "encoder.block.1.layer.0.SelfAttention.v",
               "encoder.block.0.layer.1.DenseReluDense.wi_0", "encoder.block.1.layer.1.DenseReluDense.wo",
               "decoder.block.0.layer.0.SelfAttention.k", "decoder.block.1.layer.1.EncDecAttention.k",
               "decoder.block.1.layer.1.EncDecAttention.o", "decoder.block.0.layer.2.DenseReluDense.wi_1",
               "lm_head"]


def build_ref_vit(eva, d, sd):
    m = eva.VisionTransformer(img_size=d.img_size, patch_size=d.patch, use_mean_pooling=False,
                              embed_dim=d.vit_width, depth=d.vit_depth, num_heads=d.vit_heads,
                              mlp_ratio=4.3637, qkv_bias=True, drop_path_rate=0.0,
                              norm_layer=partial(nn.LayerNorm, eps=1e-6)).eval()
    m.load_state_dict({k[len("visual_encoder."):]: v for k, v in sd.items() if k.startswith("visual_encoder.")},
                      strict=True)
    return m


def build_ref_qformer(qf, d, sd):
    from transformers import BertConfig
    c = BertConfig()
    c.encoder_width, c.add_cross_attention, c.cross_attention_freq = d.vit_width, True, d.qf_cross_freq
    c.query_length, c.num_hidden_layers = d.num_query, d.qf_layers
    m = qf.BertLMHeadModel(c).eval()
    m.cls = None                                   # blip2_mr.py:259-265
    m.bert.embeddings.word_embeddings = None
    m.bert.embeddings.position_embeddings = None
    for layer in m.bert.encoder.layer:
        layer.output = None
        layer.intermediate = None
    res = m.load_state_dict({k[len("Qformer."):]: v for k, v in sd.items() if k.startswith("Qformer.")}, strict=False)
    assert not res.unexpected_keys and res.missing_keys == ["bert.embeddings.position_ids"], res
    return m


def build_ref_t5(t5, d, sd):
    from transformers import T5Config
    tc = T5Config(vocab_size=d.vocab, d_model=d.d_model, d_kv=d.d_kv, d_ff=d.d_ff, num_layers=d.t5_layers,
                  num_decoder_layers=d.t5_dec_layers, num_heads=d.t5_heads,
                  relative_attention_num_buckets=d.rel_buckets, relative_attention_max_distance=d.rel_max_dist,
                  dropout_rate=0.1, layer_norm_epsilon=d.t5_ln_eps, feed_forward_proj="gated-gelu",
                  pad_token_id=0, eos_token_id=1, decoder_start_token_id=0)
    tc.dense_act_fn = "gelu"            # blip2_mr.py:145
    tc.tie_word_embeddings = False      # FlanT5 unties lm_head
    m = t5.T5ForConditionalGeneration(tc).eval()
    msd = {}
    for k, v in sd.items():
        if not k.startswith(T5_PREFIX):
            continue
        kk = k[len(T5_PREFIX):]
        if kk.endswith(".base_layer.weight"):
            n = k[:-len(".base_layer.weight")]
            msd[kk.replace(".base_layer.weight", ".weight")] = (
                v + sd[n + ".lora_B.default.weight"] @ sd[n + ".lora_A.default.weight"])
        elif "lora_" not in kk:
            msd[kk] = v
    m.load_state_dict(msd, strict=True)
    return m


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    eva, qf, t5 = ref_shim.load_reference_modules()
    d = TINY
    sd = init_state_dict(d, seed=SEED, lora_b_std=0.02)
    tok = SyntheticT5Tokenizer()

    # ---- 1. ViT + ln_vision + Q-Former + t5_proj on 2 frames --------------------------------------
    vit, qformer = build_ref_vit(eva, d, sd), build_ref_qformer(qf, d, sd)
    g = torch.Generator().manual_seed(7)
    frames = torch.randn(2, 3, d.img_size, d.img_size, generator=g)
    with torch.no_grad():
        vit_out = vit(frames)
        ln = nn.LayerNorm(d.vit_width)
        ln.weight.data, ln.bias.data = sd["ln_vision.weight"], sd["ln_vision.bias"]
        image_embeds = ln(vit_out)
        q_out = qformer.bert(query_embeds=sd["query_tokens"].expand(2, -1, -1), encoder_hidden_states=image_embeds,
                             encoder_attention_mask=torch.ones(2, d.vit_tokens, dtype=torch.long),
                             return_dict=True).last_hidden_state
        proj = torch.nn.functional.linear(q_out, sd["t5_proj.weight"], sd["t5_proj.bias"])
    np.savez_compressed(os.path.join(HERE, "vision_tiny.npz"),
                        frames_checksum=np.float64(frames.double().sum().item()),
                        vit_tokens=np.array(VIT_TOKENS), vit_out=vit_out[:, VIT_TOKENS].numpy(),
                        vit_out_absmean=np.float64(vit_out.abs().mean().item()),
                        image_embeds=image_embeds[:, VIT_TOKENS].numpy(),
                        qformer_out=q_out.numpy(), t5_proj_out=proj[:, :, ::8].numpy())

    # ---- 2. T5 (merged LoRA): loss, logits, grads -------------------------------------------------
    ref_t5 = build_ref_t5(t5, d, sd)
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 72, d.d_model, generator=g) * 2.0
    mask = torch.ones(2, 72, dtype=torch.long)
    mask[1, 60:] = 0
    labels = torch.randint(2, 1000, (2, 9), generator=g)
    labels[:, -1] = 1
    labels[1, 6:] = -100
    labels[1, 5] = 1
    dmask = (labels != -100).long()
    emb.requires_grad_(True)
    out = ref_t5(inputs_embeds=emb, attention_mask=mask, labels=labels, decoder_attention_mask=dmask,
                 return_dict=True)
    out.loss.backward()
    grads = {}
    for name in LORA_PROBES:
        mod = ref_t5.get_submodule(name)
        G = mod.weight.grad
        A, B = sd[T5_PREFIX + name + ".lora_A.default.weight"], sd[T5_PREFIX + name + ".lora_B.default.weight"]
        grads["gA." + name] = (B.t() @ G).numpy()
        gB = G @ A.t()
        grads["gB." + name] = (gB[::16] if name == "lm_head" else gB).numpy()
    np.savez_compressed(os.path.join(HERE, "t5_tiny.npz"),
                        emb_checksum=np.float64(emb.detach().double().sum().item()),
                        labels=labels.numpy(), mask=mask.numpy(), loss=np.float64(out.loss.item()),
                        logits_head=out.logits[:, :, :256].detach().numpy(),
                        logits_lse=torch.logsumexp(out.logits.detach(), -1).numpy(),
                        enc_out=out.encoder_last_hidden_state[:, ::8, ::4].detach().numpy(),
                        d_emb=emb.grad[:, ::4, ::4].numpy(), **grads)

    # ---- 3. whole forward_mr: reference sub-modules + restated prompt_concatenation ---------------
    samples = synth.make_samples(batch=2, frames=3, seed=3)
    with torch.no_grad():
        b, t = samples["video"].shape[:2]
        ie = ln(vit(samples["video"].reshape(-1, 3, d.img_size, d.img_size)))
        qo = qformer.bert(query_embeds=sd["query_tokens"].expand(b * t, -1, -1), encoder_hidden_states=ie,
                          encoder_attention_mask=torch.ones(b * t, d.vit_tokens, dtype=torch.long),
                          return_dict=True).last_hidden_state
        f = torch.nn.functional.linear(qo, sd["t5_proj.weight"], sd["t5_proj.bias"]).reshape(b, -1, d.d_model)
        inputs, atts = ob.prompt_concatenation(sd, d, tok, samples["timestamps"], samples["duration"], f,
                                               samples["video_prompt_end"], samples["query_prompt"],
                                               samples["task_prompt"], d.num_query)
        ans = tok(samples["relevant_windows"], padding="longest", truncation=True, max_length=200,
                  return_tensors="pt")
        lab = ans.input_ids.masked_fill(ans.input_ids == 0, -100)
        o = ref_t5(inputs_embeds=inputs, attention_mask=atts, labels=lab,
                   decoder_attention_mask=ans.attention_mask, return_dict=True)
    np.savez_compressed(os.path.join(HERE, "forward_mr_tiny.npz"),
                        video_checksum=np.float64(samples["video"].double().sum().item()),
                        loss=np.float64(o.loss.item()), labels=lab.numpy(), L_enc=np.int64(inputs.shape[1]),
                        inputs_embeds=inputs[:, ::16, ::8].numpy(), logits_head=o.logits[:, :, :256].numpy(),
                        logits_lse=torch.logsumexp(o.logits, -1).numpy())

    # ---- 4. string helpers ------------------------------------------------------------------------
    ru = ref_shim.load_reference_utils()
    corpus = ["[[0, 1], [4, 7]]", "[[0, 1] [4, 7]]", "[[5 2],, [4,, 7]]</s>junk", "garbage", "[[10, 3]]",
              "[[1, 2, 3]]", "[[0, 1],, [4, 7],]", "[[12, 40], [52, 60]]</s>", "[]", "[[a, b]]", "[[3 9] [1 2]]",
              "[[0, 150]]", "[[ 7, 3 ]]", "[[1, 2]] trailing", "[[1,2],[3,4]]", "[[40, 12], [60, 52]]", ""]
    ts = [torch.tensor([1.25, 3.75, 6.5, 149.6]), torch.tensor([0.4, 10.5, 11.5, 29.9])]
    du = torch.tensor([150.0, 30.2])
    table = {3: 4, 150: 151}
    si = ru.get_timestamps_as_seconds_integers(ts, du, table)
    with open(os.path.join(HERE, "mr_utils_golden.json"), "w") as fjson:
        json.dump({"post_process": [[c, ru.post_process(c)] for c in corpus],
                   "moment_str_to_list": [[c, ru.moment_str_to_list(ru.post_process(c))] for c in corpus],
                   "seconds_integers": {"timestamps": [t.tolist() for t in ts], "durations": du.tolist(),
                                        "table": {str(k): v for k, v in table.items()},
                                        "out_ts": [t.tolist() for t in si[0]], "out_dur": si[1], "out_prompt": si[2]}},
                  fjson, indent=1)
    print("golden vectors written to", HERE)
    for fn in sorted(os.listdir(HERE)):
        print("  %-28s %8d bytes" % (fn, os.path.getsize(os.path.join(HERE, fn))))


if __name__ == "__main__":
    main()
