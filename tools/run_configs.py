"""The other BASELINE.json configs at FULL size on one GPU (parity cases at tiny size live in tests/; this checks that the
full-size shapes run and how long they take): Charades-STA (batch 8, 20 frames, frame-token aggregation "mean"),
ActivityNet stress (batch 2, 120 frames -> L_enc ~ 4017), generate (60 frames, beam 4).  Usage: python tools/run_configs.py [out.json]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from mr_blip_b200.blip2_mr import BLIP2_MR  # noqa: E402
from mr_blip_b200.dims import FULL, init_state_dict  # noqa: E402
from oracle import synth  # noqa: E402


def timed(fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, r


def main():
    sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
    model = BLIP2_MR(dims=FULL, state_dict=sd).cuda().train()
    del sd
    res = {}

    def train_cfg(name, batch, frames, agg, duration):
        model.frame_token_aggregation = agg
        s = synth.make_samples(batch=batch, frames=frames, query_words=32, seed=7, duration=duration)
        s["video"] = s["video"].cuda()

        def step():
            for p in model.parameters():
                p.grad = None
            loss = model(s)["loss"]
            loss.backward()
            return loss
        for _ in range(3):                       # eager, capture, replay
            step()
        dt, loss = timed(step, 3)
        host = model._host_phase(s, bucket=model.graph_bucket)
        res[name] = {"batch": batch, "frames": frames, "aggregation": agg, "L_enc": host["Le"], "L_dec": host["Ld"],
                     "ms_per_step": round(dt * 1e3, 2), "clips_per_s": round(batch / dt, 2), "loss": round(loss.item(), 4),
                     "loss_finite": bool(torch.isfinite(loss))}
        print(name, res[name], flush=True)
        model.frame_token_aggregation = None
        model.reset_graphs()
        torch.cuda.empty_cache()

    train_cfg("charades_sta_b8_t20_mean", 8, 20, "mean", 120.0)
    train_cfg("activitynet_b2_t120", 2, 120, None, 180.0)
    model.eval()
    for batch in (4, 16):
        s = synth.make_samples(batch=batch, frames=60, query_words=32, seed=9)
        s["video"] = s["video"].cuda()
        for _ in range(2):                       # eager, then graph capture of every decode position
            model.generate(s, num_beams=4, max_length=50)
        dt, out = timed(lambda: model.generate(s, num_beams=4, max_length=50), 1)
        res["generate_b%d_t60_beam4" % batch] = {"batch": batch, "s_per_call": round(dt, 3), "clips_per_s": round(batch / dt, 2),
                                                "new_tokens": int(out["sequences"].shape[1]), "sample": out["raw_prediction"][0][:40]}
        print("generate", res["generate_b%d_t60_beam4" % batch], flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
