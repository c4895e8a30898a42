"""FULL-depth parity (ViT-g 39 blocks, Q-Former 12 layers, FlanT5-XL 24 + 24 layers, full widths) on the GPU box.

Three runs of the same step on the same seeded weights and inputs:
  ref    the fp32 oracle (oracle/, the CPU restatement of the reference) evaluated on the GPU in true fp32 (TF32 off)
  eager  the same oracle in the reference's own GPU regime (oracle/eager.py: fp16 ViT weights, fp16 autocast for ViT /
         Q-Former, bf16 autocast for T5 -- eva_vit.py:397-412, blip2_mr.py:446,512, moment_retrieval.py:217), i.e. what `lavis`
         computes on this device
  cuda   the product (BLIP2_MR.forward_mr through libmrblip_b200.so, rate-0 dropout so that all three evaluate one function)

BASELINE.json's north_star asks for "logits within 1e-3 rel of reference"; no 16-bit path -- the reference's own included -- gets
there through 48 T5 layers, so the gate is the one SURVEY.md section 7 defines:  err(cuda vs ref) <= 1.5 x err(eager vs ref)
for the Q-Former query embeddings, the T5 logits, the loss and the gradients of t5_proj and five LoRA adapters spread over the
stacks.  Cases: BASELINE.json configs[0] (1 clip, 4 frames, 8-word query) and one clip of the QVH config (60 frames, L_enc 2033).
The measured errors are written to gpurun_out/full_depth_parity.json (copied to profiles/ for the record).
"""
import json
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]

from mr_blip_b200.dims import FULL, T5_PREFIX, init_state_dict  # noqa: E402

GRADS = ["t5_proj.weight",
         T5_PREFIX + "encoder.block.0.layer.0.SelfAttention.q.lora_A.default.weight",
         T5_PREFIX + "encoder.block.23.layer.1.DenseReluDense.wo.lora_B.default.weight",
         T5_PREFIX + "decoder.block.0.layer.1.EncDecAttention.k.lora_A.default.weight",
         T5_PREFIX + "decoder.block.23.layer.2.DenseReluDense.wi_0.lora_B.default.weight",
         T5_PREFIX + "lm_head.lora_A.default.weight"]
GATE = 1.5
RESULTS = {}


def _rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def full():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mr_blip_b200.blip2_mr import BLIP2_MR
    sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
    model = BLIP2_MR(dims=FULL, state_dict=sd, train_dropout=False, cuda_graphs=False).cuda().train()
    yield sd, model
    del model, sd
    torch.cuda.empty_cache()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if RESULTS and os.path.isdir(out):
        json.dump(RESULTS, open(os.path.join(out, "full_depth_parity.json"), "w"), indent=1)


def _oracle_run(sd, tok, samples, amp):
    """loss / logits / qformer / grads of the oracle on the GPU: amp=False true fp32, amp=True the reference's autocast regime."""
    from oracle import blip2_mr as ob, eager
    osd = eager.reference_gpu_state_dict(sd) if amp else dict(sd)
    leaves = eager.trainable_leaves(sd)
    osd.update(leaves)
    scale = 65536.0 if amp else 1.0                          # GradScaler's initial scale (fp16 t5_proj gradients)
    out = ob.forward_mr(osd, FULL, tok, samples, amp=amp)
    (out["loss"] * scale).backward()
    res = {"loss": out["loss"].item(), "logits": out["logits"].detach().float(), "qformer": out["qformer"].detach().float(),
           "grads": {k: leaves[k].grad.float() / scale for k in GRADS},
           "all_grads": {k: v.grad.float() / scale for k, v in leaves.items()}}
    del out, leaves, osd
    torch.cuda.empty_cache()
    return res


@pytest.mark.parametrize("name,frames,qwords", [("config1_1clip_4frames_8word_query", 4, 8), ("qvh_1clip_60frames", 60, 32)])
def test_full_depth_parity_vs_fp32_oracle_and_eager_autocast(full, name, frames, qwords):
    from oracle import synth
    sd, model = full
    tok = model.t5_tokenizer
    samples = synth.make_samples(batch=1, frames=frames, query_words=qwords, seed=4)
    samples["video"] = samples["video"].cuda()
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = _oracle_run(sd, tok, samples, amp=False)
        eag = _oracle_run(sd, tok, samples, amp=True)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    for p in model.parameters():
        p.grad = None
    res = model.forward_mr(samples, want_logits=True)
    res["loss"].backward()
    torch.cuda.synchronize()
    cuda = {"loss": res["loss"].item(), "logits": res["logits"].float(), "qformer": res["qformer"].float().reshape(ref["qformer"].shape),
            "grads": {k: model._get(k).grad.float() for k in GRADS}}
    assert cuda["logits"].shape == ref["logits"].shape

    def errs(x):
        e = {"qformer": _rel(x["qformer"], ref["qformer"]), "logits": _rel(x["logits"], ref["logits"]),
             "loss": abs(x["loss"] - ref["loss"]) / abs(ref["loss"])}
        e.update({"grad " + k.replace(T5_PREFIX, "").replace(".default.weight", ""): _rel(x["grads"][k], ref["grads"][k]) for k in GRADS})
        return e

    e_cuda, e_eager = errs(cuda), errs(eag)
    # every trainable gradient, for the record (not gated one by one): the ten with the largest cuda / eager error ratio
    allg = []
    for k, g_ref in ref["all_grads"].items():
        ec, ee = _rel(model._get(k).grad.float(), g_ref), _rel(eag["all_grads"][k], g_ref)
        allg.append((ec / max(ee, 1e-12), k.replace(T5_PREFIX, "").replace(".default.weight", ""), ec, ee))
    allg.sort(reverse=True)
    import statistics
    worst = [{"param": k, "err_cuda": ec, "err_eager": ee, "ratio": r} for r, k, ec, ee in allg[:10]]
    summary = {"n": len(allg), "median_ratio": statistics.median(r for r, *_ in allg),
               "median_err_cuda": statistics.median(ec for _, _, ec, _ in allg), "median_err_eager": statistics.median(ee for *_, ee in allg)}
    RESULTS[name + ".all_gradients"] = {"summary": summary, "worst_ratio": worst}
    RESULTS[name] = {"L_enc": int(res["inputs_embeds"].shape[1]), "loss_fp32_oracle": ref["loss"], "loss_cuda": cuda["loss"],
                     "loss_eager_autocast": eag["loss"], "err_cuda_vs_fp32": e_cuda, "err_eager_autocast_vs_fp32": e_eager,
                     "gate": "err_cuda <= %.1f x err_eager" % GATE}
    print(json.dumps(RESULTS[name], indent=1))
    # The eager regime returns its loss as an fp16 number (resolution 2^-7 at |loss| ~ 11: 10.796875, 11.1953125), so its own "error"
    # of 1.4e-4 - 3.2e-4 is wherever the rounding happened to land inside half a unit (3.6e-4 relative); the product's fp32 loss moves
    # by +-1.5e-4 with nothing but the summation order of a norm kernel (measured: 1.47e-4 <-> 2.48e-4, 2.97e-4 <-> 3.5e-5).  The
    # scalar is therefore gated against that resolution; tensors (embeddings, logits, gradients) are gated against the measured error.
    floor = {"loss": 0.5 * 2.0 ** -7 / abs(ref["loss"])}
    bad = {k: (e_cuda[k], e_eager[k]) for k in e_cuda if e_cuda[k] > GATE * max(e_eager[k], floor.get(k, 0.0)) + 1e-6}
    assert not bad, "product further from the fp32 oracle than %.1f x the reference's own autocast regime: %s" % (GATE, bad)
    # absolute sanity at full depth (bf16 operands through 24 + 24 layers); the 2-layer tolerances of test_model_gpu.py do not apply
    assert e_cuda["qformer"] < 5e-3 and e_cuda["logits"] < 5e-2 and e_cuda["loss"] < 2e-3, e_cuda
