"""Run-to-run and batch-permutation determinism of the forward pass at full size (eager launches): prints the loss of
repeated identical steps, of a permuted batch, and with MRB_OVERLAP-style side streams disabled."""
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
sp = {k: (v[perm] if torch.is_tensor(v) else [v[i] for i in perm]) for k, v in s.items()}


def fwd(x, per_clip=False):
    with torch.no_grad():
        r = model.forward_mr(x, want_logits=True)
    torch.cuda.synchronize()
    return r["loss"].item(), r["logits"].float().cpu(), r["inputs_embeds"].float().cpu()


a = [fwd(s) for _ in range(3)]
print("same batch, 3 runs: loss", [x[0] for x in a])
print("  logits bitwise equal run0/run1:", torch.equal(a[0][1], a[1][1]), " inputs_embeds equal:", torch.equal(a[0][2], a[1][2]))
b = fwd(sp)
print("permuted batch: loss", b[0])
print("  inputs_embeds (un-permuted) equal:", torch.equal(b[2], a[0][2][perm]), " max |d| %.3e" % (b[2] - a[0][2][perm]).abs().max().item())
print("  logits equal:", torch.equal(b[1], a[0][1][perm]), " max |d| %.3e" % (b[1] - a[0][1][perm]).abs().max().item())
vit, qf, t5 = model.engines()
vit.overlap = False
t5.overlap = False
c = [fwd(s) for _ in range(2)]
print("side streams off: loss", [x[0] for x in c], " logits equal to run0:", torch.equal(c[0][1], a[0][1]))
