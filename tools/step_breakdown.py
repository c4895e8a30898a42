"""Per-C-ABI-call time breakdown of one full-size QVH training step (CUDA events around every call, warm).
Usage: python tools/step_breakdown.py [out.json]"""
import collections
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from mr_blip_b200 import _lib  # noqa: E402
from mr_blip_b200.blip2_mr import BLIP2_MR  # noqa: E402
from mr_blip_b200.dims import FULL, init_state_dict  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
    model = BLIP2_MR(dims=FULL, state_dict=sd).cuda().train()
    del sd
    samples = synth.make_samples(batch=4, frames=60, query_words=32, seed=100)
    samples["video"] = samples["video"].cuda()
    # ---- the product path first: captured CUDA graph of the device half of the step
    for _ in range(3):
        model(samples)["loss"].backward()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(4):
        model(samples)["loss"].backward()
    g1.record()
    torch.cuda.synchronize()
    graph_ms = g0.elapsed_time(g1) / 4
    t0 = time.perf_counter()
    model(samples)["loss"].backward()
    graph_host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    # ---- eager launches of the same kernels for the per-call breakdown
    model.cuda_graphs = False
    for _ in range(2):
        model(samples)["loss"].backward()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model(samples)["loss"].backward()
    t_host = time.perf_counter() - t0           # host time to ENQUEUE a step
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    _lib.PROFILE = []
    from mr_blip_b200 import ops
    ops.GEMM_PROFILE = []
    # phase markers
    marks = {}
    vit, qf, t5 = model.engines()
    orig = {"vit": vit.forward, "qf": qf.forward, "enc_f": t5.encoder_forward, "dec_f": t5.decoder_forward,
            "dec_b": t5.decoder_backward, "enc_b": t5.encoder_backward}

    def wrap(name, fn):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            marks[name] = (e0, e1)
            return r
        return inner

    vit.forward, qf.forward = wrap("vit", orig["vit"]), wrap("qf", orig["qf"])
    t5.encoder_forward, t5.decoder_forward = wrap("enc_f", orig["enc_f"]), wrap("dec_f", orig["dec_f"])
    t5.decoder_backward, t5.encoder_backward = wrap("dec_b", orig["dec_b"]), wrap("enc_b", orig["enc_b"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model(samples)["loss"].backward()
    e1.record()
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    gprof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
    shapes = collections.defaultdict(lambda: [0, 0.0])
    for m, n, k, a, b in gprof:
        shapes[(m, n, k)][0] += 1
        shapes[(m, n, k)][1] += a.elapsed_time(b)
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for name, a, b in prof:
        tot[name] += a.elapsed_time(b)
        cnt[name] += 1
    step_ms = e0.elapsed_time(e1)
    res = {"graph_step_ms": graph_ms, "graph_host_ms": graph_host_ms, "step_ms_with_events": step_ms, "host_enqueue_ms": t_host * 1e3, "wall_ms_no_events": t_wall * 1e3,
           "phases_ms": {k: a.elapsed_time(b) for k, (a, b) in marks.items()},
           "gemm_shapes": [{"M": m, "N": n, "K": k, "calls": c, "ms": round(t, 3), "tflops": round(2.0 * m * n * k * c / t / 1e9, 1)}
                           for (m, n, k), (c, t) in sorted(shapes.items(), key=lambda x: -x[1][1])],
           "ops": {k: {"ms": round(v, 3), "calls": cnt[k]} for k, v in sorted(tot.items(), key=lambda x: -x[1])}}
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
