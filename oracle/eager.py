"""The reference's GPU regime applied to the oracle restatement: what `lavis` does when it trains Mr. BLIP on a CUDA device.
Test / baseline infrastructure only (tests/test_full_depth_gpu.py, bench.py's `eager_gpu` leg) -- never the product path.

  * ViT weights in fp16 (convert_weights_to_fp16, eva_vit.py:397-412: Conv / Linear weights and biases; `vit_precision: fp16`)
  * ViT + ln_vision under torch.amp.autocast fp16 (blip2_mr.py:446), Q-Former + t5_proj under the training loop's fp16 autocast
    (moment_retrieval.py:217), prompt assembly + T5 under autocast bf16 (blip2_mr.py:512)
  * one optimiser step as the task's train loop does it: GradScaler.scale(loss).backward(), scaler.step, scaler.update,
    zero_grad (moment_retrieval.py:215-238 / base_task.py:229-247), AdamW over the trainable tensors (runner_base.py:103-131)
  * train mode: torch's own F.dropout at the reference's sites (Dropper(torch_rng=True)); eval mode: none
"""
import torch

from . import blip2_mr as _mr
from .dropout import Dropper


def reference_gpu_state_dict(sd):
    """fp16 copies of the ViT's Conv / Linear weights and biases (what convert_weights_to_fp16 touches); the rest is shared."""
    out = dict(sd)
    for k, v in sd.items():
        if k.startswith("visual_encoder.") and (k.endswith((".weight", ".bias")) and ".norm" not in k):
            if ".qkv." in k or ".proj." in k or ".fc1." in k or ".fc2." in k or "patch_embed.proj" in k:
                out[k] = v.half()
    return out


def trainable_leaves(sd):
    """The tensors the reference trains under task=qformer_freeze_lora: LoRA A/B and t5_proj (blip2_mr.py:183-235,291)."""
    return {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if "lora_" in k or k.startswith("t5_proj.")}


class EagerTrainer:
    """One eager training step of the oracle in the reference's GPU regime."""

    def __init__(self, sd, d, tok, train_dropout=True, lr=1e-5, weight_decay=0.05):
        self.d, self.tok = d, tok
        self.sd = reference_gpu_state_dict(sd)
        self.leaves = trainable_leaves(sd)
        self.sd.update(self.leaves)
        self.drop = Dropper(0, torch_rng=True) if train_dropout else None
        self.opt = torch.optim.AdamW(list(self.leaves.values()), lr=lr, weight_decay=weight_decay)
        self.scaler = torch.amp.GradScaler("cuda")

    def step(self, samples, frame_token_aggregation=None):
        out = _mr.forward_mr(self.sd, self.d, self.tok, samples, frame_token_aggregation=frame_token_aggregation,
                             drop=self.drop, amp=True)
        loss = out["loss"]
        self.scaler.scale(loss).backward()
        self.scaler.step(self.opt)
        self.scaler.update()
        self.opt.zero_grad(set_to_none=True)
        return loss.detach()
