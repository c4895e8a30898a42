"""Loading the third-party weights the reference pulls from hubs -- FlanT5 (transformers `from_pretrained`,
blip2_mr.py:140-147) and EVA ViT-g (`create_eva_vit_g`, eva_vit.py:415-442) -- from LOCAL files into the model's
parameter names.  The BLIP-2 / Mr. BLIP checkpoints themselves ({"model": ...} with Q-Former, t5_proj, LoRA keys) already
carry the model's names and go through BaseModel.load_checkpoint."""
import glob
import json
import os

import torch

from .dims import T5_PREFIX


def read_state_dict(path):
    """A .safetensors / .bin / .pth / .pt file, or a directory of (sharded) transformers weights -> one flat dict.
    A {"model": ...} or {"state_dict": ...} wrapper is unwrapped."""
    if os.path.isdir(path):
        files = []
        for index in ("model.safetensors.index.json", "pytorch_model.bin.index.json"):
            ip = os.path.join(path, index)
            if os.path.isfile(ip):
                files = sorted({os.path.join(path, f) for f in json.load(open(ip))["weight_map"].values()})
                break
        if not files:
            files = sorted(glob.glob(os.path.join(path, "*.safetensors"))) or sorted(glob.glob(os.path.join(path, "pytorch_model*.bin")))
        if not files:
            raise RuntimeError("no weight files under %s" % path)
        out = {}
        for f in files:
            out.update(read_state_dict(f))
        return out
    if not os.path.isfile(path):
        raise RuntimeError("weights path is invalid: %s" % path)
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    sd = torch.load(path, map_location="cpu")
    for key in ("model", "state_dict", "module"):
        if isinstance(sd, dict) and key in sd and isinstance(sd[key], dict):
            sd = sd[key]
    return sd


def hf_t5_to_model_keys(hf_sd, model_keys, prefix=T5_PREFIX):
    """transformers T5 names -> this model's: everything moves under the peft prefix, and the weight of a Linear that carries a
    LoRA adapter here sits one level down, at `.base_layer.weight` (what peft's wrapper does to q, k, v, o, wi_0, wi_1, wo
    and lm_head).  Keys the model does not have (e.g. a tied copy the file repeats) are returned separately."""
    model_keys = set(model_keys)
    mapped, unknown = {}, []
    for k, v in hf_sd.items():
        name = prefix + k
        if name not in model_keys and k.endswith(".weight"):
            wrapped = prefix + k[:-len(".weight")] + ".base_layer.weight"
            if wrapped in model_keys:
                name = wrapped
        if name in model_keys:
            mapped[name] = v
        else:
            unknown.append(k)
    return mapped, unknown


def _apply(model, mapped, what):
    want = {k: v.shape for k, v in model.state_dict().items() if k in mapped}
    bad = [k for k, v in mapped.items() if tuple(v.shape) != tuple(want[k])]
    if bad:
        raise RuntimeError("%s: shape mismatch for %s (file %s, model %s)" % (what, bad[0], tuple(mapped[bad[0]].shape), tuple(want[bad[0]])))
    msg = model.load_state_dict(mapped, strict=False)
    assert not msg.unexpected_keys
    if hasattr(model, "_weights_changed"):
        model._weights_changed()
    return msg


def load_hf_t5(model, path_or_sd, prefix=T5_PREFIX):
    """FlanT5 weights (a local transformers directory / file, or an already loaded state dict) into model.t5_model; the LoRA
    adapters keep their values.  Returns (number of tensors loaded, file keys without a home)."""
    sd = read_state_dict(path_or_sd) if isinstance(path_or_sd, str) else path_or_sd
    mapped, unknown = hf_t5_to_model_keys(sd, model.state_dict().keys(), prefix)
    need = [k for k in model.state_dict() if k.startswith(prefix) and "lora_" not in k and "embed_tokens" not in k and k not in mapped]
    if need:
        raise RuntimeError("T5 weights are incomplete: no tensor for %s (+%d more)" % (need[0], len(need) - 1))
    _apply(model, mapped, "T5")
    return len(mapped), unknown


def load_eva_vit(model, path_or_sd):
    """eva_vit_g.pth into model.visual_encoder (eva_vit.py:432-441: strict=False -- the file has one block more than the 39
    the model runs, a final norm and a head).  16-bit parameters are filled with the rounded values
    (convert_weights_to_fp16).  Returns (number loaded, skipped file keys)."""
    sd = read_state_dict(path_or_sd) if isinstance(path_or_sd, str) else path_or_sd
    have = model.state_dict()
    mapped = {"visual_encoder." + k: v for k, v in sd.items() if "visual_encoder." + k in have}
    skipped = [k for k in sd if "visual_encoder." + k not in have]
    need = [k for k in have if k.startswith("visual_encoder.") and k not in mapped]
    if need:
        raise RuntimeError("ViT weights are incomplete: no tensor for %s (+%d more)" % (need[0], len(need) - 1))
    _apply(model, mapped, "ViT")
    return len(mapped), skipped
