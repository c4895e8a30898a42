"""Run-to-run and batch-permutation determinism at full size (eager launches, 4 QVH clips); one script, four checks:

python tools/determinism_check.py forward   no-grad forward: 3 identical runs, a permuted batch, side streams off (logits bitwise)
python tools/determinism_check.py train     4 train steps (fwd + bwd): loss, logits of steps 1 / 3, gradient differences step to step
python tools/determinism_check.py permuted  train-mode forward of a permuted batch vs the no-grad forward (logits bitwise)
python tools/determinism_check.py first     first-step effect: three call styles, each in a FRESH process
(default: forward train permuted).  Result of round 1: profiles/determinism_r01b.log."""
import subprocess
import sys

import torch

sys.path.insert(0, ".")
PERM = [2, 0, 3, 1]


def setup(cuda_graphs=False):
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from mr_blip_b200.dims import FULL, init_state_dict
    from oracle import synth
    sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
    model = BLIP2_MR(dims=FULL, state_dict=sd, cuda_graphs=cuda_graphs).cuda().train()
    del sd
    s = synth.make_samples(batch=4, frames=60, query_words=32, seed=100)
    s["video"] = s["video"].cuda()
    sp = {k: (v[PERM] if torch.is_tensor(v) else [v[i] for i in PERM]) for k, v in s.items()}
    return model, s, sp


def zero_grads(model):
    for p in model.parameters():
        p.grad = None


def check_forward():
    model, s, sp = setup()

    def fwd(x):
        with torch.no_grad():
            r = model.forward_mr(x, want_logits=True)
        torch.cuda.synchronize()
        return r["loss"].item(), r["logits"].float().cpu(), r["inputs_embeds"].float().cpu()

    a = [fwd(s) for _ in range(3)]
    print("same batch, 3 runs: loss", [x[0] for x in a])
    print("  logits bitwise equal run0/run1:", torch.equal(a[0][1], a[1][1]), " inputs_embeds equal:", torch.equal(a[0][2], a[1][2]))
    b = fwd(sp)
    print("permuted batch: loss", b[0])
    print("  inputs_embeds (un-permuted) equal:", torch.equal(b[2], a[0][2][PERM]), " max |d| %.3e" % (b[2] - a[0][2][PERM]).abs().max().item())
    print("  logits equal:", torch.equal(b[1], a[0][1][PERM]), " max |d| %.3e" % (b[1] - a[0][1][PERM]).abs().max().item())
    vit, qf, t5 = model.engines()
    vit.overlap = False
    t5.overlap = False
    c = [fwd(s) for _ in range(2)]
    print("side streams off: loss", [x[0] for x in c], " logits equal to run0:", torch.equal(c[0][1], a[0][1]))


def check_train():
    model, s, _ = setup()
    outs = []
    for i in range(4):
        zero_grads(model)
        r = model.forward_mr(s, want_logits=(i % 2 == 1))
        r["loss"].backward()
        torch.cuda.synchronize()
        outs.append((r["loss"].item(), r.get("logits"), model.flat_grads().clone()))
        print("train step %d want_logits=%s loss %.7f  |g| %.6e" % (i, i % 2 == 1, outs[-1][0], outs[-1][2].norm().item()), flush=True)
    print("logits step1 == step3:", torch.equal(outs[1][1], outs[3][1]))
    for a, b in ((0, 1), (1, 2), (2, 3)):
        print("grad rel diff %d/%d: %.3e" % (a, b, ((outs[a][2] - outs[b][2]).norm() / outs[b][2].norm()).item()))
    with torch.no_grad():
        print("no_grad loss %.7f" % model.forward_mr(s)["loss"].item())


def check_permuted():
    model, s, sp = setup()

    def train(x, want_logits):
        zero_grads(model)
        r = model.forward_mr(x, want_logits=want_logits)
        v0 = r["loss"].item()
        r["loss"].backward()
        torch.cuda.synchronize()
        return v0, r["loss"].item(), r.get("logits")

    a, b, c = train(s, False), train(sp, False), train(sp, True)
    with torch.no_grad():
        d = model.forward_mr(sp, want_logits=True)
    e = train(s, False)
    print("train s        loss %.7f" % a[1])
    print("train sp       loss %.7f (before backward() call %.7f)" % (b[1], b[0]))
    print("train sp +lgts loss %.7f" % c[1])
    print("no_grad sp     loss %.7f   logits equal to train-mode logits: %s  max|d| %.3e" %
          (d["loss"].item(), torch.equal(d["logits"], c[2]), (d["logits"].float() - c[2].float()).abs().max().item()))
    print("train s again  loss %.7f" % e[1])


def first_step_variant(variant):
    model, s, _ = setup(cuda_graphs=True)
    model.cuda_graphs = False
    vals = []
    for _ in range(3):
        zero_grads(model)
        if variant == "logits":
            r = model.forward_mr(s, want_logits=True)
            r["loss"].backward()
            vals.append((r["loss"].item(), r["logits"].clone(), r["inputs_embeds"].clone()))
            continue
        loss = model(s)["loss"]
        v = loss.item() if variant == "call_item_first" else None
        (loss * 1.0).backward()
        vals.append(v if v is not None else loss.item())
    if variant == "logits":
        print(variant, [v[0] for v in vals], "logits 0==1", torch.equal(vals[0][1], vals[1][1]), "emb 0==1", torch.equal(vals[0][2], vals[1][2]),
              "max dlogit %.3e" % (vals[0][1] - vals[1][1]).abs().max().item())
    else:
        print(variant, vals)


def check_first():
    for v in ("call", "call_item_first", "logits"):
        out = subprocess.run([sys.executable, __file__, "first:" + v], capture_output=True, text=True, timeout=250)
        print(out.stdout.strip()[-400:] or out.stderr[-600:], flush=True)


if __name__ == "__main__":
    CHECKS = {"forward": check_forward, "train": check_train, "permuted": check_permuted, "first": check_first}
    for name in (sys.argv[1:] or ["forward", "train", "permuted"]):
        if name.startswith("first:"):
            first_step_variant(name.split(":", 1)[1])
            continue
        print("==== %s" % name, flush=True)
        CHECKS[name]()
