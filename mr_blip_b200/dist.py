"""Data-parallel plumbing: one process per GPU, NCCL over NVLink (reference: lavis/common/dist_utils.py:57-90
init_distributed_mode and the DDP wrap at lavis/runners/runner_base.py:89-96).

The path shards by clip with no data-path collective; the only exchange is the all-reduce of the 19.5 M
trainable gradients (LoRA A/B + t5_proj, 78 MB fp32) once per optimiser step.  The backward kernels write them into
one flat buffer, so GradAllReducer issues one in-place NCCL all-reduce (AVG) instead of DDP's 25 MB buckets.  The model is
also compatible with torch DDP itself (gradients are handed to autograd, see blip2_mr._HandOverGrads)."""
import datetime
import os

import torch
import torch.distributed as dist


def init_distributed_mode(backend=None):
    """env:// rendezvous from RANK / WORLD_SIZE / LOCAL_RANK (dist_utils.py:57-90). -> (rank, world, local_rank)."""
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        return 0, 1, 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, init_method="env://", world_size=world, rank=rank,
                                timeout=datetime.timedelta(minutes=30))
        dist.barrier()
    return rank, world, local


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def barrier():
    if is_dist():
        dist.barrier()


def cleanup():
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


class GradAllReducer:
    """Gradient averaging over the trainable parameters with ONE all-reduce.  When the model keeps its gradients in one
    flat buffer (BLIP2_MR.flat_grads(): every .grad is a view of it) that buffer is reduced in place; otherwise the
    gradients are packed into a flat staging buffer, reduced and copied back."""

    def __init__(self, params, flat_fn=None):
        self.params = [p for p in params if p.requires_grad]
        self.flat_fn = flat_fn
        self.flat, self.views = None, []

    def _staging(self):
        if self.flat is None:
            n = sum(p.numel() for p in self.params)
            self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
            off = 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        return self.flat

    def __call__(self):
        if not is_dist() or not self.params:
            return
        world = dist.get_world_size()
        flat = self.flat_fn() if self.flat_fn is not None else None
        if flat is not None:                                 # zero-copy: the .grad tensors alias `flat`
            if dist.get_backend() == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.div_(world)
            return
        flat = self._staging()
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
