"""Permuted batch, full depth: does the TRAIN-mode forward (activations saved, lse written, backward run) produce the same
logits / loss as the no-grad forward?  Mirrors the order of tests/test_full_size_gpu.py."""
import sys
import torch
sys.path.insert(0, ".")
from mr_blip_b200.blip2_mr import BLIP2_MR
from mr_blip_b200.dims import FULL, init_state_dict
from oracle import synth

sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
model = BLIP2_MR(dims=FULL, state_dict=sd, cuda_graphs=False).cuda().train()
del sd
s = synth.make_samples(batch=4, frames=60, query_words=32, seed=100)
s["video"] = s["video"].cuda()
perm = [2, 0, 3, 1]


def train(x, want_logits):
    for p in model.parameters():
        p.grad = None
    r = model.forward_mr(x, want_logits=want_logits)
    v0 = r["loss"].item()                     # read before the backward kernels are even launched? no: backward ran inside forward
    r["loss"].backward()
    torch.cuda.synchronize()
    return v0, r["loss"].item(), r.get("logits")


a = train(s, False)
g = model.flat_grads().clone()
sp = {k: (v[perm] if torch.is_tensor(v) else [v[i] for i in perm]) for k, v in s.items()}
b = train(sp, False)
c = train(sp, True)
with torch.no_grad():
    d = model.forward_mr(sp, want_logits=True)
e = train(s, False)
print("train s        loss %.7f" % a[1])
print("train sp       loss %.7f (before backward() call %.7f)" % (b[1], b[0]))
print("train sp +lgts loss %.7f" % c[1])
print("no_grad sp     loss %.7f   logits equal to train-mode logits: %s  max|d| %.3e" %
      (d["loss"].item(), torch.equal(d["logits"], c[2]), (d["logits"].float() - c[2].float()).abs().max().item()))
print("train s again  loss %.7f" % e[1])
