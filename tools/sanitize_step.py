"""One eager training step (fwd+bwd) of a 2-layer-deep, full-width model at the QVH shape (batch 4, 60 frames, unpadded
L_enc 2033 / L_dec 14, clips ordered [2, 0, 3, 1]) -- target for `compute-sanitizer --tool memcheck`."""
import sys
import torch
sys.path.insert(0, ".")
from mr_blip_b200.blip2_mr import BLIP2_MR
from mr_blip_b200.dims import TINY, init_state_dict
from oracle import synth

sd = init_state_dict(TINY, seed=1234, lora_b_std=0.02, device="cuda")
model = BLIP2_MR(dims=TINY, state_dict=sd, cuda_graphs=False).cuda().train()
del sd
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
s = synth.make_samples(batch=4, frames=frames, query_words=32, seed=100)
s["video"] = s["video"].cuda()
perm = [2, 0, 3, 1]
sp = {k: (v[perm] if torch.is_tensor(v) else [v[i] for i in perm]) for k, v in s.items()}
for x in (s, sp):
    for p in model.parameters():
        p.grad = None
    loss = model(x)["loss"]
    loss.backward()
    torch.cuda.synchronize()
    print("loss %.7f" % loss.item(), flush=True)
print("done")
