"""In-graph time of the four T5 phases of the QVH training step (B 4, L_enc 2037 -> bucket 2048, 16 target tokens), each
captured into its OWN CUDA graph and replayed: encoder forward, decoder chain (decoder forward + lm_head + cross-entropy +
lm_head / decoder backward, incl. the side-stream cross K/V work), encoder backward.  The ncu launch list serialises kernels and
flushes caches between them, which overstates the ~2 000 decoder-sized kernels; this is the number they cost inside a graph.
python tools/t5_phase_bench.py [out.json]      (MRB_TRAIN_DROPOUT=0: eval arithmetic; MRB_T5_PHASES=dec_chain: that phase only)"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from mr_blip_b200 import ops  # noqa: E402
from mr_blip_b200.blip2_mr import BLIP2_MR  # noqa: E402
from mr_blip_b200.dims import FULL, init_state_dict  # noqa: E402
from mr_blip_b200.dropout import DropState  # noqa: E402
from mr_blip_b200.t5 import shift_right  # noqa: E402


def main():
    B, Le, Ld = 4, int(os.environ.get("MRB_T5_LE", "2048")), 16
    sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
    model = BLIP2_MR(dims=FULL, state_dict=sd).cuda().train()
    del sd
    t5 = model.engines()[2]
    d = FULL
    if os.environ.get("MRB_TRAIN_DROPOUT", "1") != "0":
        t5.drop = DropState()
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, Le, d.d_model, device="cuda", generator=g) * 0.5
    kmask = torch.ones(B, Le, dtype=torch.int32, device="cuda")
    labels = torch.randint(3, 1000, (B, Ld), device="cuda", generator=g)
    dec_ids = shift_right(labels)
    dmask = torch.ones(B, Ld, dtype=torch.int32, device="cuda")
    flat = torch.zeros(t5.n_grad_elems(), dtype=torch.float32, device="cuda")
    t5.bind_grads(flat)
    st = {}

    def enc_fwd():
        st["enc_saves"] = []
        st["enc_ext"], st["enc_h"], st["enc_bias"] = t5.encoder_forward(x.reshape(B * Le, -1), kmask, B, Le, st["enc_saves"])

    def dec_chain():
        saves = []
        dec_ext, dec_h, dec_bias = t5.decoder_forward(dec_ids, dmask, st["enc_ext"], kmask, B, Ld, Le, saves)
        M = B * Ld
        logits = t5.lm_head.forward(dec_ext, M, out_dtype=torch.float32)
        loss = torch.zeros(1, dtype=torch.float32, device="cuda")
        dlogits = t5._ext(M, d.vocab)
        ops.cross_entropy(logits, labels.reshape(-1).contiguous(), None, dlogits, -1.0, loss_sum=loss)
        ddec = t5.lm_head.backward(dlogits, dec_ext, M)
        st["d_enc"] = t5.decoder_backward(saves, dec_h, ddec, dmask, st["enc_ext"], kmask, B, Ld, Le, dec_bias)
        t5.side_join()

    def enc_bwd():
        # the saves are consumed (cleared) by the backward: re-run needs fresh saves, so this graph holds forward + backward and
        # the forward's time is subtracted
        enc_fwd()
        t5.encoder_backward(st["enc_saves"], st["enc_h"], st["d_enc"].clone(), kmask, B, Le, st["enc_bias"])
        t5.side_join()

    res = {}
    only = os.environ.get("MRB_T5_PHASES", "")          # e.g. "dec_chain": time that phase alone (the encoder forward runs once, untimed)
    if only:
        enc_fwd()
    for name, fn in (("enc_fwd", enc_fwd), ("dec_chain", dec_chain), ("enc_fwd_bwd", enc_bwd)):
        if only and name not in only.split(","):
            continue
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, capture_error_mode="thread_local"):
            fn()
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        res[name] = round(e0.elapsed_time(e1) / reps, 3)
        print(name, res[name], "ms", flush=True)
    if "enc_fwd_bwd" in res and "enc_fwd" in res:
        res["enc_bwd"] = round(res["enc_fwd_bwd"] - res["enc_fwd"], 3)
    res["dropout"] = t5.drop is not None
    print(json.dumps(res))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
