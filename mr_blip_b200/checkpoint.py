"""Checkpoint files in the wire format the reference's runner writes and reads (lavis/runners/runner_base.py:572-644):
{"model": state dict WITHOUT the frozen parameters, "optimizer", "config", "scaler", "epoch"}.  A file written here
resumes under the LAVIS runner and the other way round; the key names are the model's (peft-style LoRA keys included)."""
import os

import torch


REFERENCE_ONLY_BUFFERS = (".position_ids",)


def trainable_state_dict(model):
    """state_dict() minus every parameter with requires_grad False (runner_base.py:577-587): for Mr. BLIP that leaves the
    LoRA factors and t5_proj (19.5 M values).  Buffers and anything else named_parameters() does not list stay -- which
    includes the tied aliases encoder/decoder.embed_tokens.weight of the frozen T5 embedding, exactly as in the files the
    reference writes."""
    model = getattr(model, "module", model)
    frozen = {k for k, p in model.named_parameters() if not p.requires_grad}
    return {k: v for k, v in model.state_dict().items() if k not in frozen}


def save_checkpoint(model, optimizer, output_dir, cur_epoch, is_best=False, config=None, scaler=None):
    """Write checkpoint_{epoch|best}.pth under output_dir (runner_base.py:588-600); returns the path."""
    obj = {"model": trainable_state_dict(model),
           "optimizer": optimizer.state_dict() if optimizer is not None else None,
           "config": config.to_dict() if hasattr(config, "to_dict") else config,
           "scaler": scaler.state_dict() if scaler else None,
           "epoch": cur_epoch}
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, "checkpoint_{}.pth".format("best" if is_best else cur_epoch))
    torch.save(obj, path)
    return path


def resume_checkpoint(model, optimizer, path, scaler=None, map_location="cpu"):
    """Resume as runner_base.py:621-644 does -- model (partial files load non-strictly, as _reload_best_model :602-619
    falls back to), optimizer and scaler state -- and return the epoch to start from."""
    if not os.path.isfile(path):
        raise RuntimeError("checkpoint url or path is invalid")
    ckpt = torch.load(path, map_location=map_location)
    model = getattr(model, "module", model)
    # persistent buffers the reference's modules carry and this parameter tree does not (Qformer.py:69 position_ids): files written
    # by the reference runner hold them, they carry no information
    state = {k: v for k, v in ckpt["model"].items() if not k.endswith(REFERENCE_ONLY_BUFFERS)}
    msg = model.load_state_dict(state, strict=False)
    if msg.unexpected_keys:
        raise RuntimeError("unexpected keys in checkpoint: %s" % msg.unexpected_keys[:8])
    if hasattr(model, "_weights_changed"):
        model._weights_changed()
    if optimizer is not None and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    if scaler and ckpt.get("scaler") is not None:
        scaler.load_state_dict(ckpt["scaler"])
    return ckpt["epoch"] + 1
