"""Optimiser parameter groups and the learning-rate schedules the mr_BLIP recipes name, for stand-alone use of the
package (inside LAVIS the runner keeps its own: lavis/runners/runner_base.py:103-131, lavis/common/optims.py:14-119).
Host-side only; the AdamW update itself is torch's fused kernel."""
import math

import torch

from .registry import registry


def param_groups(model, weight_decay):
    """Two groups as lavis/runners/runner_base.py:108-124 splits them: trainable tensors of rank >= 2 whose name holds
    neither "bias", "ln" nor "bn" decay; everything else trainable does not.  Returns (groups, n_trainable_elements)."""
    wd, no_wd, n = [], [], 0
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if p.ndim < 2 or "bias" in name or "ln" in name or "bn" in name:
            no_wd.append(p)
        else:
            wd.append(p)
        n += p.data.nelement()
    return [{"params": wd, "weight_decay": float(weight_decay)}, {"params": no_wd, "weight_decay": 0}], n


def build_optimizer(model, init_lr, weight_decay, beta2=0.999, fused=None):
    """AdamW over `param_groups` with betas (0.9, beta2) (runner_base.py:125-131).  `fused` defaults to True when the
    parameters live on a GPU."""
    groups, _ = param_groups(model, weight_decay)
    if fused is None:
        fused = any(p.is_cuda for g in groups for p in g["params"])
    return torch.optim.AdamW(groups, lr=float(init_lr), weight_decay=float(weight_decay), betas=(0.9, beta2),
                             fused=bool(fused))


def _set_lr(optimizer, lr):
    for g in optimizer.param_groups:
        g["lr"] = lr


def cosine_lr_schedule(optimizer, epoch, max_epoch, init_lr, min_lr):
    """Half-cosine from init_lr (epoch 0) to min_lr (epoch max_epoch) -- optims.py:104-110."""
    _set_lr(optimizer, (init_lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * epoch / max_epoch)) + min_lr)


def warmup_lr_schedule(optimizer, step, max_step, init_lr, max_lr):
    """Linear ramp init_lr -> max_lr over max_step steps, clipped at max_lr -- optims.py:113-117."""
    _set_lr(optimizer, min(max_lr, init_lr + (max_lr - init_lr) * step / max(max_step, 1)))


def step_lr_schedule(optimizer, epoch, init_lr, min_lr, decay_rate):
    """Geometric decay per epoch, floored at min_lr -- optims.py:120-124."""
    _set_lr(optimizer, max(min_lr, init_lr * (decay_rate ** epoch)))


@registry.register_lr_scheduler("linear_warmup_step_lr")
class LinearWarmupStepLRScheduler:
    """optims.py:14-53: per-step linear warm-up during epoch 0, then a per-epoch geometric decay."""

    def __init__(self, optimizer, max_epoch, min_lr, init_lr, decay_rate=1, warmup_start_lr=-1, warmup_steps=0,
                 **kwargs):
        self.optimizer, self.max_epoch, self.min_lr, self.init_lr = optimizer, max_epoch, min_lr, init_lr
        self.decay_rate, self.warmup_steps = decay_rate, warmup_steps
        self.warmup_start_lr = warmup_start_lr if warmup_start_lr >= 0 else init_lr

    def step(self, cur_epoch, cur_step):
        if cur_epoch == 0:
            warmup_lr_schedule(self.optimizer, cur_step, self.warmup_steps, self.warmup_start_lr, self.init_lr)
        else:
            step_lr_schedule(self.optimizer, cur_epoch, self.init_lr, self.min_lr, self.decay_rate)


@registry.register_lr_scheduler("linear_warmup_cosine_lr")
class LinearWarmupCosineLRScheduler:
    """optims.py:56-101, the schedule every mr_BLIP recipe uses: linear warm-up counted in GLOBAL steps (epoch x the
    largest step index seen so far + step), then a cosine that moves once per epoch."""

    def __init__(self, optimizer, max_epoch, min_lr, init_lr, warmup_steps=0, warmup_start_lr=-1, **kwargs):
        self.optimizer, self.max_epoch, self.min_lr, self.init_lr = optimizer, max_epoch, min_lr, init_lr
        self.warmup_steps = warmup_steps
        self.warmup_start_lr = warmup_start_lr if warmup_start_lr >= 0 else init_lr
        self.max_iters_per_epoch = 0

    def step(self, cur_epoch, cur_step):
        self.max_iters_per_epoch = max(self.max_iters_per_epoch, cur_step)
        done = cur_epoch * self.max_iters_per_epoch + cur_step
        if done < self.warmup_steps:
            warmup_lr_schedule(self.optimizer, done, self.warmup_steps, self.warmup_start_lr, self.init_lr)
        else:
            cosine_lr_schedule(self.optimizer, cur_epoch, self.max_epoch, self.init_lr, self.min_lr)


def build_lr_scheduler(optimizer, run_cfg):
    """Scheduler from a recipe's `run` section (runner_base.py:152-188: lr_sched, max_epoch, min_lr, init_lr,
    lr_decay_rate, warmup_lr, warmup_steps)."""
    get = run_cfg.get if hasattr(run_cfg, "get") else (lambda k, d=None: getattr(run_cfg, k, d))
    cls = registry.get_lr_scheduler_class(get("lr_sched"))
    if cls is None:
        raise KeyError("unknown lr_sched %r" % (get("lr_sched"),))
    return cls(optimizer=optimizer, max_epoch=int(get("max_epoch")), min_lr=float(get("min_lr")),
               init_lr=float(get("init_lr")), decay_rate=get("lr_decay_rate", None),
               warmup_start_lr=float(get("warmup_lr", -1)), warmup_steps=int(get("warmup_steps", 0)))
