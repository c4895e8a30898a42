"""Bisect the first-step effect seen by tests/test_full_size_gpu.py: variants run in fresh processes."""
import subprocess
import sys

BODY = '''
import sys, torch
sys.path.insert(0, ".")
from mr_blip_b200.blip2_mr import BLIP2_MR
from mr_blip_b200.dims import FULL, init_state_dict
from oracle import synth
sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
model = BLIP2_MR(dims=FULL, state_dict=sd).cuda().train()
del sd
s = synth.make_samples(batch=4, frames=60, query_words=32, seed=100)
s["video"] = s["video"].cuda()
model.cuda_graphs = False
VARIANT = "%s"
vals = []
for i in range(3):
    for p in model.parameters():
        p.grad = None
    if VARIANT == "call":
        loss = model(s)["loss"]
        (loss * 1.0).backward()
        vals.append(loss.item())
    elif VARIANT == "call_item_first":
        loss = model(s)["loss"]
        v = loss.item()
        (loss * 1.0).backward()
        vals.append(v)
    elif VARIANT == "logits":
        r = model.forward_mr(s, want_logits=True)
        r["loss"].backward()
        vals.append((r["loss"].item(), r["logits"].clone(), r["inputs_embeds"].clone()))
if VARIANT == "logits":
    print(VARIANT, [v[0] for v in vals], "logits 0==1", torch.equal(vals[0][1], vals[1][1]), "emb 0==1", torch.equal(vals[0][2], vals[1][2]),
          "max dlogit %%.3e" %% (vals[0][1] - vals[1][1]).abs().max().item())
else:
    print(VARIANT, vals)
'''
for v in ("call", "call_item_first", "logits"):
    out = subprocess.run([sys.executable, "-c", BODY % v], capture_output=True, text=True, timeout=250)
    print(out.stdout.strip()[-400:] or out.stderr[-600:], flush=True)
