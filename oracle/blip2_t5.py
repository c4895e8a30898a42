"""fp32 restatement of Blip2T5.forward / generate (lavis/models/blip2_models/blip2_t5.py:99-151, 152-255, image branch),
composed around the pinned sub-module oracles.  The frozen T5 has no adapters: the peft-named state dict is used with
lora_B = 0, which is the same function.  `max_text_length` (:122, an AttributeError as shipped) is read as max_txt_len.
Test infrastructure only (see oracle/__init__.py)."""
import torch

from . import t5 as _t5
from .beam_search import beam_search
from .blip2_mr import frame_tokens


def _inputs(sd, d, tok, image, text):
    f, _ = frame_tokens(sd, d, image.unsqueeze(1))                       # [B, 32, D]: one "frame" per image
    emb = sd[_t5.PREFIX + "shared.weight"]
    inputs = torch.cat([f, emb[text.input_ids]], dim=1)
    atts = torch.cat([torch.ones(f.shape[:2], dtype=torch.long), text.attention_mask], dim=1)
    return inputs, atts


def forward(sd, d, tok, samples, max_txt_len=32):
    text = tok(samples["text_input"], padding="longest", truncation=True, max_length=max_txt_len, return_tensors="pt")
    out_t = tok(samples["text_output"], padding="longest", truncation=True, max_length=max_txt_len, return_tensors="pt")
    inputs, atts = _inputs(sd, d, tok, samples["image"], text)
    labels = out_t.input_ids.masked_fill(out_t.input_ids == tok.pad_token_id, -100)
    return _t5.t5_forward(sd, d, inputs, atts, labels, out_t.attention_mask)


@torch.no_grad()
def generate(sd, d, tok, samples, num_beams=5, max_length=30, min_length=1, length_penalty=1.0):
    image = samples["image"]
    prompt = samples.get("prompt", "")
    if isinstance(prompt, str):
        prompt = [prompt] * image.size(0)
    text = tok(prompt, padding="longest", return_tensors="pt")
    inputs, atts = _inputs(sd, d, tok, image, text)
    enc = _t5.t5_encoder(sd, d, inputs, atts)
    enc_b = enc.repeat_interleave(num_beams, dim=0)
    atts_b = atts.repeat_interleave(num_beams, dim=0)

    def step(ids):
        dec = _t5.t5_decoder(sd, d, ids, enc_b, atts_b)
        return _t5.t5_logits(sd, d, dec[:, -1])

    seqs = beam_search(step, image.size(0), num_beams, max_length, min_length, length_penalty,
                       eos_id=tok.eos_token_id, pad_id=tok.pad_token_id)
    return tok.batch_decode(seqs, skip_special_tokens=True), seqs
