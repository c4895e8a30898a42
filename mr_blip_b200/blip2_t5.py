"""Blip2T5 (image BLIP-2 with a frozen FlanT5) behind the LAVIS surface, on the same sm_100a kernels as BLIP2_MR.

Thin sibling of mr_blip_b200/blip2_mr.py for lavis/models/blip2_models/blip2_t5.py: registry name "blip2_t5",
from_config keys (:365-396), forward(samples) -> {"loss"} (:99-151), generate(samples, ...) -> list[str] (:152-255),
state-dict names visual_encoder.*, ln_vision.*, Qformer.bert.*, query_tokens, t5_proj.*, t5_model.* (plain HF T5
names: no peft wrapper here, the T5 is frozen bf16, :81-85).  It shares every kernel of the hot path: ViT ->
ln_vision -> Q-Former -> t5_proj -> [32 query tokens | prompt embeddings] -> T5.

Scope: forward computes the loss value only.  In the reference this class trains the Q-Former (stage-2 BLIP-2
pre-training); the Q-Former / ViT backward is not part of the Mr. BLIP hot path (every mr_BLIP recipe freezes both,
SURVEY.md §3.1), so the returned loss carries no autograd graph.  As shipped the reference forward raises AttributeError
(self.max_text_length at :122/:129 vs self.max_txt_len at :93); the evident intent max_txt_len is used.
"""
import logging

import numpy as np
import torch

from . import ops
from .base_model import attach, disabled_train
from .blip2_mr import Blip2Base
from .dims import Dims, FULL, T5_PREFIX, init_state_dict
from .registry import registry
from .t5 import T5Engine, shift_right
from .tokenizer import load_t5_tokenizer
from .vision import VitEngine, QFormerEngine

PLAIN_PREFIX = "t5_model."


def plain_t5_state_dict(sd):
    """peft-style names of init_state_dict (t5_model.base_model.model.X.base_layer.weight + lora_A/B) -> the plain HF names
    Blip2T5 checkpoints carry (t5_model.X.weight), LoRA adapters dropped."""
    out = {}
    for k, v in sd.items():
        if not k.startswith(T5_PREFIX):
            out[k] = v
        elif "lora_" in k:
            continue
        else:
            out[PLAIN_PREFIX + k[len(T5_PREFIX):].replace(".base_layer.", ".")] = v
    return out


@registry.register_model("blip2_t5")
class Blip2T5(Blip2Base):
    PRETRAINED_MODEL_CONFIG_DICT = {
        "pretrain_flant5xl": "configs/models/blip2/blip2_pretrain_flant5xl.yaml",
        "pretrain_flant5xxl": "configs/models/blip2/blip2_pretrain_flant5xxl.yaml",
        "caption_coco_flant5xl": "configs/models/blip2/blip2_caption_flant5xl.yaml",
    }

    def __init__(self, img_size=224, drop_path_rate=0, use_grad_checkpoint=False, vit_precision="fp16", freeze_vit=True,
                 num_query_token=32, t5_model="google/flan-t5-xl", prompt="", max_txt_len=32, apply_lemmatizer=False,
                 dims: Dims = None, init_seed=1234, state_dict=None, tokenizer=None):
        super().__init__()
        self.dims = d = dims or FULL
        assert img_size == d.img_size and num_query_token == d.num_query
        sd = state_dict if state_dict is not None else plain_t5_state_dict(init_state_dict(d, seed=init_seed))
        shared = None
        for name, t in sd.items():
            if name.startswith(PLAIN_PREFIX) and name.endswith("embed_tokens.weight"):
                continue
            t = t.clone()
            if name.startswith("visual_encoder.") and vit_precision == "fp16" and "pos_embed" not in name and "cls_token" not in name \
                    and (t.ndim >= 2 or name.endswith((".fc1.bias", ".fc2.bias", ".proj.bias"))):
                t = t.half()                                 # convert_weights_to_fp16, eva_vit.py:397-412
            # reference: Q-Former, query_tokens and t5_proj are the trainable part (stage 2); ViT and T5 are frozen
            train = name.startswith(("Qformer.", "t5_proj.")) or name == "query_tokens"
            p = attach(self, name, t, requires_grad=train and t.is_floating_point())
            if name == PLAIN_PREFIX + "shared.weight":
                shared = p
        for side in ("encoder", "decoder"):
            attach(self, PLAIN_PREFIX + side + ".embed_tokens.weight", shared)
        if freeze_vit:
            self.visual_encoder.eval()
            self.visual_encoder.train = disabled_train.__get__(self.visual_encoder)
            logging.info("freeze vision encoder")
        self.t5_tokenizer = tokenizer or load_t5_tokenizer(t5_model)
        self.max_txt_len = max_txt_len
        self.prompt = prompt
        self._apply_lemmatizer = apply_lemmatizer
        self._engines = None
        self._zeros = {}

    @classmethod
    def from_config(cls, cfg):
        get = cfg.get
        model = cls(img_size=get("image_size", 224), num_query_token=get("num_query_token", 32),
                    t5_model=get("t5_model", "google/flan-t5-xl"), drop_path_rate=get("drop_path_rate", 0),
                    use_grad_checkpoint=get("use_grad_checkpoint", False), vit_precision=get("vit_precision", "fp16"),
                    freeze_vit=get("freeze_vit", True), prompt=get("prompt", ""), max_txt_len=get("max_txt_len", 32),
                    apply_lemmatizer=get("apply_lemmatizer", False), dims=get("dims", None), init_seed=get("init_seed", 1234))
        import os
        from . import weights
        if isinstance(get("t5_model", None), str) and os.path.isdir(get("t5_model")):      # local transformers directory
            weights.load_hf_t5(model, get("t5_model"), prefix=PLAIN_PREFIX)
        if get("vit_weights", None):
            weights.load_eva_vit(model, get("vit_weights"))
        for key in ("pretrained", "finetuned") if get("load_finetuned", True) else ("pretrained",):
            path = get(key, None)
            if path and os.path.isfile(path):
                model.load_checkpoint(path)
        return model

    def _weights_changed(self):
        self._engines = None

    # ---------------------------------------------------------------------------------------------
    def _get(self, name):
        obj = self
        for p in name.split("."):
            obj = obj._modules[p] if p in obj._modules else obj._parameters[p]
        return obj

    def _get_t5(self, name):
        """The T5 engine asks for peft-style names; this frozen T5 has none: base weights under the plain names, zero adapters."""
        if ".lora_A.default.weight" in name or ".lora_B.default.weight" in name:
            base = self._get(name.split(".lora_")[0] + ".weight")
            shape = (self.dims.lora_r, base.shape[1]) if ".lora_A." in name else (base.shape[0], self.dims.lora_r)
            if name not in self._zeros:
                self._zeros[name] = torch.zeros(shape, dtype=torch.float32, device="cuda")
            return self._zeros[name]
        return self._get(name.replace(".base_layer.", "."))

    def engines(self):
        if self._engines is None:
            if not torch.cuda.is_available() or self.device.type != "cuda":
                raise RuntimeError("Blip2T5 runs only on a CUDA device through libmrblip_b200.so "
                                   "(no CPU / eager fallback); move the model with .cuda()")
            d = self.dims
            vit, qf = VitEngine(d, self._get), QFormerEngine(d, self._get)
            t5 = T5Engine(d, self._get_t5, prefix=PLAIN_PREFIX)
            t5.overlap = False
            qf.set_t5_proj(self.t5_proj.weight, self.t5_proj.bias)
            self._engines = (vit, qf, t5)
        return self._engines

    # ---------------------------------------------------------------------------------------------
    def _inputs(self, image, text):
        """ViT -> ln_vision -> Q-Former -> t5_proj, then [32 query tokens | text embeddings] (blip2_t5.py:100-140).
        -> inputs_embeds fp32 [B, 32 + Lt, D] (cuda), attention mask int32 [B, 32 + Lt]."""
        vit, qf, t5 = self.engines()
        d = self.dims
        B, n = image.shape[0], d.num_query
        img = image.to(device="cuda", non_blocking=True) if image.dtype == torch.uint8 else \
            image.to(device="cuda", dtype=torch.float32, non_blocking=True)
        qf.set_t5_proj(self.t5_proj.weight, self.t5_proj.bias)
        x = vit.forward(img)
        _, h16 = qf.forward(x, B)
        frames = qf.project(h16)                                        # [B*32, D] fp32
        Lt = text.input_ids.shape[1]
        table = np.empty((B, n + Lt), dtype=np.int32)
        for b in range(B):
            table[b, :n] = -(b * n + np.arange(n)) - 1                  # frame-token rows, as in the BLIP2_MR row table
        table[:, n:] = text.input_ids.numpy()
        idx = torch.from_numpy(table.reshape(-1)).to("cuda", non_blocking=True)
        inputs = torch.empty((B * (n + Lt), d.d_model), dtype=torch.float32, device="cuda")
        ops.gather_rows(idx, t5.emb, frames, inputs)
        atts = torch.cat([torch.ones((B, n), dtype=torch.long), text.attention_mask], dim=1)
        return inputs.view(B, n + Lt, d.d_model), atts.to(device="cuda", dtype=torch.int32)

    @torch.no_grad()
    def forward(self, samples):
        """blip2_t5.py:99-151: loss of text_output given the image and text_input (value only, see the module docstring)."""
        _, _, t5 = self.engines()
        tok = self.t5_tokenizer
        text = tok(samples["text_input"], padding="longest", truncation=True, max_length=self.max_txt_len, return_tensors="pt")
        out_t = tok(samples["text_output"], padding="longest", truncation=True, max_length=self.max_txt_len, return_tensors="pt")
        inputs, atts = self._inputs(samples["image"], text)
        labels = out_t.input_ids.masked_fill(out_t.input_ids == tok.pad_token_id, -100).to("cuda")
        dmask = out_t.attention_mask.to(device="cuda", dtype=torch.int32)
        res = t5.loss_device(inputs, atts, labels, shift_right(labels), dmask, backward=False, want_logits=True)
        self._last_logits = res["logits"]
        return {"loss": res["loss"].reshape(())}

    @torch.no_grad()
    def generate(self, samples, use_nucleus_sampling=False, num_beams=5, max_length=30, min_length=1, top_p=0.9,
                 repetition_penalty=1.0, length_penalty=1.0, num_captions=1, temperature=1):
        """blip2_t5.py:152-255 (image branch): beam search over [query tokens | prompt] -> list[str]."""
        from .generation import beam_search
        if use_nucleus_sampling or num_captions != 1 or repetition_penalty != 1.0:
            raise NotImplementedError("only deterministic beam search with one caption per image is implemented")
        _, _, t5 = self.engines()
        image = samples["image"]
        prompt = samples["prompt"] if "prompt" in samples else self.prompt
        if isinstance(prompt, str):
            prompt = [prompt] * image.size(0)
        else:
            assert len(prompt) == image.size(0), "The number of prompts must be equal to the batch size."
        text = self.t5_tokenizer(prompt, padding="longest", return_tensors="pt")
        inputs, atts = self._inputs(image, text)
        seqs = beam_search(t5, inputs, atts, num_beams=num_beams, max_new_tokens=max_length, min_length=min_length,
                           length_penalty=length_penalty, eos_id=self.t5_tokenizer.eos_token_id,
                           pad_id=self.t5_tokenizer.pad_token_id)
        self._last_sequences = seqs
        return self.t5_tokenizer.batch_decode(seqs, skip_special_tokens=True)
