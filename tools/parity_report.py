"""Print error metrics of the CUDA path against the golden vectors (reference modules) and the CPU
oracle on the TINY configuration.  Used to calibrate the tolerances written in tests/test_model_gpu.py."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from mr_blip_b200.blip2_mr import BLIP2_MR  # noqa: E402
from mr_blip_b200.dims import TINY, T5_PREFIX, init_state_dict  # noqa: E402
from mr_blip_b200.mr_utils import post_process  # noqa: E402
from oracle import blip2_mr as ob, synth, t5 as ot5  # noqa: E402

G = "tests/golden"


def rel(got, want):
    got, want = torch.as_tensor(got).float().cpu(), torch.as_tensor(want).float().cpu()
    return "max|err| %.3e  rel-fro %.3e  max|want| %.3e" % ((got - want).abs().max().item(),
                                                          ((got - want).norm() / want.norm().clamp_min(1e-30)).item(),
                                                          want.abs().max().item())


def main():
    torch.manual_seed(0)
    d = TINY
    sd = init_state_dict(d, seed=1234, lora_b_std=0.02)
    model = BLIP2_MR(dims=d, state_dict=sd).cuda()
    vit, qf, t5 = model.engines()

    gold = np.load(os.path.join(G, "vision_tiny.npz"))
    frames = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(7))
    x = vit.forward(frames.cuda())
    h, h16, ie, _ = qf.forward(x, 2, return_all=True)
    p = qf.project(h16)
    tk = gold["vit_tokens"].tolist()
    print("vit_out      ", rel(x.view(2, 257, -1)[:, tk], gold["vit_out"]))
    print("image_embeds ", rel(ie.view(2, 257, -1)[:, tk], gold["image_embeds"]))
    print("qformer_out  ", rel(h.view(2, 32, -1), gold["qformer_out"]))
    print("t5_proj_out  ", rel(p.view(2, 32, -1)[:, :, ::8], gold["t5_proj_out"]))

    gold = np.load(os.path.join(G, "t5_tiny.npz"))
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 72, d.d_model, generator=g) * 2.0
    mask = torch.ones(2, 72, dtype=torch.long)
    mask[1, 60:] = 0
    labels = torch.randint(2, 1000, (2, 9), generator=g)
    labels[:, -1] = 1
    labels[1, 6:] = -100
    labels[1, 5] = 1
    t5.zero_grads()
    out = t5.loss(emb.cuda(), mask, labels, (labels != -100).long(), backward=True, want_logits=True)
    print("t5 loss       got %.5f want %.5f" % (out["loss"].item(), float(gold["loss"])))
    print("t5 logits    ", rel(out["logits"][:, :, :256], gold["logits_head"]))
    print("t5 lse       ", rel(torch.logsumexp(out["logits"], -1), gold["logits_lse"]))
    print("t5 enc_out   ", rel(out["encoder_last_hidden_state"][:, ::8, ::4], gold["enc_out"]))
    print("t5 d_emb     ", rel(out["d_inputs_embeds"][:, ::4, ::4], gold["d_emb"]))
    grads = {id(p_): g_ for p_, g_ in t5.param_grads()}
    for k in gold.files:
        if k.startswith("gA.") or k.startswith("gB."):
            name = k[3:]
            ab = "lora_A" if k.startswith("gA.") else "lora_B"
            par = model._get(f"{T5_PREFIX}{name}.{ab}.default.weight")
            got = grads[id(par)]
            if name == "lm_head" and ab == "lora_B":
                got = got[::16]
            print("%-58s %s" % (k, rel(got, gold[k])))

    # whole model vs oracle (fp32 CPU) incl. gradients
    samples = synth.make_samples(batch=2, frames=3, seed=3)
    gold = np.load(os.path.join(G, "forward_mr_tiny.npz"))
    model.train()
    t0 = time.time()
    res = model.forward_mr(samples, want_logits=True)
    res["loss"].backward()
    torch.cuda.synchronize()
    print("forward_mr loss got %.5f want(reference) %.5f   (%.2fs)" % (res["loss"].item(), float(gold["loss"]), time.time() - t0))
    print("inputs_embeds", rel(res["inputs_embeds"][:, ::16, ::8], gold["inputs_embeds"]))
    print("logits       ", rel(res["logits"][:, :, :256], gold["logits_head"]))
    osd = dict(sd)
    leaves = {}
    for k in list(osd):
        if "lora_" in k or k.startswith("t5_proj."):
            leaves[k] = osd[k].clone().requires_grad_(True)
            osd[k] = leaves[k]
    o = ob.forward_mr(osd, d, model.t5_tokenizer, samples)
    o["loss"].backward()
    worst = []
    for k, leaf in leaves.items():
        got = model._get(k).grad
        if got is None:
            print("NO GRAD", k)
            continue
        e = ((got.cpu() - leaf.grad).norm() / leaf.grad.norm().clamp_min(1e-30)).item()
        worst.append((e, k, leaf.grad.norm().item()))
    worst.sort(reverse=True)
    print("param grads vs oracle autograd: %d tensors, rel-fro worst 5:" % len(worst))
    for e, k, nrm in worst[:5]:
        print("   %.3e  |g|=%.3e  %s" % (e, nrm, k))
    print("   median %.3e" % sorted(w[0] for w in worst)[len(worst) // 2])

    model.eval()
    t0 = time.time()
    gen = model.generate(samples, num_beams=5, max_length=8)
    torch.cuda.synchronize()
    og = ob.generate(sd, d, model.t5_tokenizer, samples, post_process, num_beams=5, max_length=8)
    print("generate (%.2fs) cuda  :" % (time.time() - t0), gen["sequences"].tolist())
    print("generate oracle        :", og["sequences"].tolist())


if __name__ == "__main__":
    main()
