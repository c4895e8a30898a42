"""First-step effect? Train-mode eager steps (fwd+bwd) at full size: loss and logits of consecutive identical steps."""
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
outs = []
for i in range(4):
    for p in model.parameters():
        p.grad = None
    r = model.forward_mr(s, want_logits=(i % 2 == 1))
    r["loss"].backward()
    torch.cuda.synchronize()
    g = model.flat_grads()
    outs.append((r["loss"].item(), r.get("logits"), g.clone() if g is not None else None))
    print("train step %d want_logits=%s loss %.7f  |g| %.6e" % (i, i % 2 == 1, outs[-1][0], outs[-1][2].norm().item() if outs[-1][2] is not None else -1), flush=True)
print("logits step1 == step3:", torch.equal(outs[1][1], outs[3][1]))
for a, b in ((0, 1), (1, 2), (2, 3)):
    print("grad rel diff %d/%d: %.3e" % (a, b, ((outs[a][2] - outs[b][2]).norm() / outs[b][2].norm()).item()))
with torch.no_grad():
    print("no_grad loss %.7f" % model.forward_mr(s)["loss"].item())
