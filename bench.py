#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: video-clips/sec, fwd+bwd (+AdamW step), 60-frame QVH config
(batch 4 per GPU, 60 frames, 32 Q-Former queries, ViT-g + FlanT5-XL + LoRA r=8), synthetic data and
seeded random-init weights (no network).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --config {qvh,charades,anet,generate}    the other BASELINE.json configs, same JSON schema
  python bench.py --impl reference ...                     the reference's CPU path (oracle port) on host cores
  python bench.py --impl eager ...                         the oracle in the reference's autocast regime, eagerly on cuda:0
                                                           (the "eager PyTorch on a B200" bar of BASELINE.md section 4)

One JSON line on stdout (rank 0).  value = clips/s with the frame tensor resident in HBM; e2e = the
same step through the public API (BLIP2_MR.forward / generate on a samples dict holding the video in PINNED
HOST memory: the H2D copy of [B,T,3,224,224] fp32 and the D2H read of the loss / token ids are inside the timed region).
The training step runs the reference's train() arithmetic: dropout 0.1 in the Q-Former and T5, 0.05 on the LoRA inputs.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

UNIT = "clips/s"
QUERY_WORDS = 32
# BASELINE.json configs[1..4]; FLOPs per clip from SURVEY.md section 8(d) (GEMM + attention, 2 FLOPs per MAC)
CONFIGS = {
    "qvh": dict(kind="train", batch=4, frames=60, agg=None, duration=150.0, spread=7.0, flops_per_clip=44.83e12,
                metric="video_clips_per_sec_fwd_bwd_qvh60",
                what="QVH: batch 4 per GPU, 60 frames, 32 Q-Former queries"),
    "charades": dict(kind="train", batch=8, frames=20, agg="mean", duration=30.0, spread=0.5, flops_per_clip=11.30e12,
                     metric="video_clips_per_sec_fwd_bwd_charades20",
                     what="Charades-STA: batch 8 per GPU, 20 frames, 32 Q-Former queries, frame-token aggregation 'mean'"),
    "anet": dict(kind="train", batch=2, frames=120, agg=None, duration=180.0, spread=7.0, flops_per_clip=92.34e12,
                 metric="video_clips_per_sec_fwd_bwd_anet120",
                 what="ActivityNet stress: batch 2 per GPU, 120 frames (L_enc ~ 4017), 32 Q-Former queries"),
    "generate": dict(kind="generate", batch=16, frames=60, agg=None, duration=150.0, spread=7.0, flops_per_clip=38.4e12,
                     beams=4, max_new_tokens=50, metric="video_clips_per_sec_generate_beam4_t60",
                     what="generate: batch 16 per GPU, 60 frames, T5 beam search (4 beams, 50 new tokens max)"),
}
NCU_FC1_TRAFFIC = 924.1e6          # dram__bytes_read.sum + dram__bytes_write.sum, one fc1 launch (profiles/ncu_gemm2_fc1_r02f.csv)

_JSON_FD = None


def _reserve_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 (NCCL prints its version banner there on the
    first communicator) goes to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback"


def make_samples(cfg, seed, batch=None, frames=None):
    from oracle import synth        # synthetic samples generator only (input construction, not compute)
    return synth.make_samples(batch=batch or cfg["batch"], frames=frames or cfg["frames"], query_words=QUERY_WORDS, seed=seed,
                              duration=cfg["duration"], spread=cfg["spread"])


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU reference arm
CPU_FRAMES, CPU_T5_LAYERS = 4, 2


def cpu_reference_sample(cfg, threads):
    """A bounded sample of ONE clip of `cfg` through the oracle (CPU fp32 restatement of the reference path), full widths:
      A. frame encoder (ViT-g 39 blocks -> ln_vision -> Q-Former 12 layers -> t5_proj) on CPU_FRAMES of the clip's frames
         -- every frame costs the same, so the clip costs frames / CPU_FRAMES times that;
      B. T5 at the clip's FULL encoder length (the L^2 attention term included) on CPU_T5_LAYERS + CPU_T5_LAYERS of the 24 + 24
         layers, loss + backward to the LoRA leaves and inputs_embeds (train configs) or encoder only + a few no-cache decode
         steps (generate) -- every layer costs the same, so the stack costs 24 / CPU_T5_LAYERS times that.
    Layer weights of part A are one set of tensors aliased across layers (timing only).  -> callable returning seconds per clip."""
    import dataclasses
    from mr_blip_b200.dims import FULL, T5_PREFIX, init_state_dict
    from mr_blip_b200.tokenizer import SyntheticT5Tokenizer
    from oracle import blip2_mr as ob, t5 as ot5
    torch.set_num_threads(threads)
    d = FULL
    small = dataclasses.replace(d, vit_depth=1, qf_layers=2, t5_layers=CPU_T5_LAYERS, t5_dec_layers=CPU_T5_LAYERS)
    one = init_state_dict(small, seed=1, lora_b_std=0.02)
    sd = dict(one)
    for fmt, n, src in (("visual_encoder.blocks.%d.", d.vit_depth, lambda i: 0),
                        ("Qformer.bert.encoder.layer.%d.", d.qf_layers, lambda i: i % 2)):
        for i in range(n):
            s_, d_ = fmt % src(i), fmt % i
            for k in list(one):
                if k.startswith(s_):
                    sd[d_ + k[len(s_):]] = one[k]
    tok = SyntheticT5Tokenizer()
    samples = make_samples(cfg, seed=0, batch=1)
    video = samples["video"][:, :CPU_FRAMES]
    agg = cfg["agg"]
    n_tok = 1 if agg else d.num_query
    # the clip's real prompt (all frames) gives the encoder length; the frame rows of part B are random stand-ins of the
    # right shape (same arithmetic)
    with torch.no_grad():
        f4, _ = ob.frame_tokens(sd, d, video, agg)
        fake = torch.randn(1, cfg["frames"] * n_tok, d.d_model) * f4.std()
        inputs, atts = ob.prompt_concatenation(sd, small, tok, samples["timestamps"], samples["duration"], fake,
                                               samples["video_prompt_end"], samples["query_prompt"], samples["task_prompt"], n_tok)
    ans = tok(samples["relevant_windows"], padding="longest", truncation=True, max_length=200, return_tensors="pt")
    labels = ans.input_ids.masked_fill(ans.input_ids == tok.pad_token_id, -100)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "lora_" in k or k.startswith("t5_proj.")}
    sd.update(leaves)
    scale_a, scale_b = cfg["frames"] / CPU_FRAMES, d.t5_layers / CPU_T5_LAYERS
    info = {"L_enc": int(inputs.shape[1]), "frames_timed": CPU_FRAMES, "t5_layers_timed": CPU_T5_LAYERS}

    def step():
        t0 = time.perf_counter()
        with torch.no_grad():
            ob.frame_tokens(sd, d, video, agg)
        ta = time.perf_counter() - t0
        t0 = time.perf_counter()
        if cfg["kind"] == "train":
            x = inputs.clone().requires_grad_(True)
            out = ot5.t5_forward(sd, small, x, atts, labels, ans.attention_mask)
            out["loss"].backward()
            for v in leaves.values():
                v.grad = None
            tb = time.perf_counter() - t0
            return scale_a * ta + scale_b * tb
        with torch.no_grad():                                # generate: encoder once, then decode steps of `beams` rows
            enc = ot5.t5_encoder(sd, small, inputs, atts)
            t_enc = time.perf_counter() - t0
            nb, steps = cfg["beams"], 3
            enc_b, atts_b = enc.repeat_interleave(nb, 0), atts.repeat_interleave(nb, 0)
            ids = torch.zeros(nb, 1, dtype=torch.long)
            t0 = time.perf_counter()
            for _ in range(steps):                           # the oracle's decoder re-runs the whole prefix (no cache)
                dec = ot5.t5_decoder(sd, small, ids, enc_b, atts_b)
                nxt = ot5.t5_logits(sd, small, dec[:, -1]).argmax(-1, keepdim=True)
                ids = torch.cat([ids, nxt], 1)
            t_dec = (time.perf_counter() - t0) / steps * cfg["max_new_tokens"]
        return scale_a * ta + scale_b * (t_enc + t_dec)

    what = ("oracle port (fp32 PyTorch restatement of the reference path), ONE clip: frame encoder at full depth on %d of %d "
            "frames (x%.0f), T5 at the full encoder length %d on %d+%d of 24+24 layers (x%.0f), %s; full widths"
            % (CPU_FRAMES, cfg["frames"], scale_a, info["L_enc"], CPU_T5_LAYERS, CPU_T5_LAYERS, scale_b,
               "loss + backward" if cfg["kind"] == "train" else "encoder + 3 no-cache decode steps of %d beams (x%d/3)"
               % (cfg["beams"], cfg["max_new_tokens"])))
    return step, what


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step, what = cpu_reference_sample(cfg, threads)
    for _ in range(max(args.warmup, 0)):
        step()
    ts = [step() for _ in range(args.steps)]
    per = sum(ts) / len(ts)                                  # seconds per clip
    val = 1.0 / per
    line = {"impl": "reference", "metric": cfg["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per * 1e3 * cfg["batch"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload(cfg, None), "sample": what},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": what},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------ eager-PyTorch-on-GPU arm
def run_eager(args, cfg):
    """The oracle in the reference's GPU regime (oracle/eager.py: fp16 ViT / Q-Former autocast, bf16 T5 autocast, torch dropout,
    GradScaler + AdamW), eager launches on cuda:0 -- what `lavis` itself would run on this box ("the number
    the new kernels must beat").  Batch falls back 4 -> 2 -> 1 when the eager activations do not fit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    assert cfg["kind"] == "train", "the eager arm covers the training configs"
    from mr_blip_b200.dims import FULL, init_state_dict
    from mr_blip_b200.tokenizer import SyntheticT5Tokenizer
    from oracle import eager
    torch.cuda.set_device(0)
    sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
    tr = eager.EagerTrainer(sd, FULL, SyntheticT5Tokenizer(), train_dropout=not args.eval_dropout)
    del sd
    res = None
    b = cfg["batch"]
    while b >= 1 and res is None:
        samples = make_samples(cfg, seed=100, batch=b)
        samples["video"] = samples["video"].cuda()
        try:
            for _ in range(max(args.warmup, 1)):
                tr.step(samples, cfg["agg"])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                loss = tr.step(samples, cfg["agg"])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            res = {"value": b / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "batch": b, "loss": float(loss),
                   "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
        except torch.cuda.OutOfMemoryError:
            tr.opt.zero_grad(set_to_none=True)
            samples = None
            gc.collect()
            torch.cuda.empty_cache()
            b //= 2
    if res is None:
        res = {"unavailable": "out of memory at batch 1"}
    res.update(impl="eager", metric=cfg["metric"], steps=args.steps, warmup=max(args.warmup, 1),
               what="oracle/ restatement of the reference modules, eager PyTorch %s on cuda:0 in the reference's autocast regime "
                    "(fp16 ViT weights + fp16 autocast for ViT / Q-Former: eva_vit.py:397-412, blip2_mr.py:446, "
                    "moment_retrieval.py:217; bf16 autocast for T5: blip2_mr.py:512), %s, GradScaler + AdamW step; "
                    "video resident in HBM" % (torch.__version__, "eval-mode dropout" if args.eval_dropout else
                                               "train-mode dropout (torch RNG)"))
    emit(res)


def eager_subprocess(args, cfg_name):
    """Run the eager arm in its own process (its activations need most of the HBM) and return its JSON line."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "eager", "--config", cfg_name, "--steps", "3", "--warmup", "1"]
    if args.eval_dropout:
        cmd.append("--eval-dropout")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if lines:
            return json.loads(lines[-1])
        return {"unavailable": "no output (rc %d): %s" % (out.returncode, out.stderr.strip().splitlines()[-1][:200] if out.stderr.strip() else "")}
    except Exception as e:                                   # the eager arm is a reported baseline: never fail the bench on it
        return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


# ------------------------------------------------------------------------------------------ B200 arm
def workload(cfg, model):
    base = cfg["what"] + ", ViT-g (fp16) + FlanT5-XL (bf16) LoRA r=8"
    if cfg["kind"] == "generate":
        return base + ", eval mode; decode steps replay captured CUDA graphs"
    drop = "train-mode dropout" if (model is None or model.train_dropout) else "eval-mode dropout (A/B run)"
    return base + (", fwd+bwd + fused AdamW step, %s (Q-Former / T5 0.1, LoRA 0.05, counter-hash masks drawn per step); the device "
                   "half of the step replays one captured CUDA graph" % drop)


def run_b200(args, cfg_name, cfg):
    from mr_blip_b200 import _lib, ops, dist as mdist
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from mr_blip_b200.dims import FULL, init_state_dict
    import torch.distributed as tdist

    rank, world, local = mdist.init_distributed_mode()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    d = FULL
    train = cfg["kind"] == "train"
    eager_line = None
    if rank == 0 and world == 1 and train and not args.no_eager:
        eager_line = eager_subprocess(args, cfg_name)        # first, while this process holds no HBM
        print("[bench] eager arm: %s" % json.dumps(eager_line)[:300], file=sys.stderr, flush=True)
    t_build = time.time()
    sd = init_state_dict(d, seed=1234, lora_b_std=0.02, device="cuda")
    model = BLIP2_MR(dims=d, state_dict=sd, train_dropout=not args.eval_dropout, frame_token_aggregation=cfg["agg"]).to(dev)
    model.train(train)
    del sd
    samples = make_samples(cfg, seed=100 + rank)
    B = cfg["batch"]
    video_host = samples["video"].pin_memory()
    video_dev = video_host.to(dev)
    if train:
        trainable = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(trainable, lr=1e-5, weight_decay=0.05, fused=True)
        reducer = mdist.GradAllReducer(trainable, flat_fn=model.flat_grads)

        def step(video, read_result):
            samples["video"] = video
            loss = model(samples)["loss"]
            loss.backward()
            reducer()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return loss.item() if read_result else loss.detach()
        d2h = 4
    else:
        def step(video, read_result):
            samples["video"] = video
            out = model.generate(samples, num_beams=cfg["beams"], max_length=cfg["max_new_tokens"])
            return float(out["sequences"].shape[1])          # generate ends with the token ids on the host either way
        d2h = B * (cfg["max_new_tokens"] + 1) * 8
    build_s = time.time() - t_build

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(video, read_result, steps):
        barrier()
        l0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            last = step(video, read_result)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            ms = t.item()
        return ms, _lib.launch_count - l0, last

    def log(msg):
        print("[bench rank %d %.1fs] %s" % (rank, time.time() - t_build, msg), file=sys.stderr, flush=True)

    log("model built in %.1fs" % build_s)
    for _ in range(max(args.warmup, 3)):
        step(video_dev, False)
    log("warm-up done")
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches, last = timed(video_dev, False, args.steps)
    clocks = sampler.summary()
    log("timed region done: %.1f ms/step" % (ms / args.steps))
    for _ in range(2):
        step(video_host, True)
    ms_e2e, _, _ = timed(video_host, True, args.steps)
    log("e2e region done")
    # the same end-to-end step fed with RAW uint8 frames (normalisation fused into the patch extraction): a quarter of the PCIe
    # bytes per step.  Reported next to the headline e2e, which keeps the reference's fp32 input format.
    video_u8 = torch.randint(0, 256, tuple(video_host.shape), dtype=torch.uint8).pin_memory()
    for _ in range(3):
        step(video_u8, True)
    ms_u8, _, _ = timed(video_u8, True, args.steps)
    log("e2e uint8 region done")
    # the gradient exchange alone (N > 1): CUDA events around the all-reduce of a few more steps, max over ranks
    ar_ms = None
    if train and world > 1:
        ts = []
        for _ in range(3):
            samples["video"] = video_dev
            loss = model(samples)["loss"]
            loss.backward()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            tdist.barrier()
            a0.record()
            reducer()
            a1.record()
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
            opt.step()
            opt.zero_grad(set_to_none=True)
        t = torch.tensor([sorted(ts)[1]], device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        ar_ms = t.item()
        # every rank's OWN step time: the same steps with the exchange left out, so no rank waits for another; the spread over
        # the ranks (not the all-reduce) is what separates the N-GPU step from the 1-GPU step -- the job runs at the slowest replica
        samples["video"] = video_dev
        n_local = min(args.steps, 5)
        torch.cuda.synchronize()
        tdist.barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(n_local):
            model(samples)["loss"].backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
        r1.record()
        torch.cuda.synchronize()
        mine = torch.tensor([r0.elapsed_time(r1) / n_local], device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        tdist.all_gather(every, mine)
        replica_ms = [round(x.item(), 2) for x in every]
    clips = B * world * args.steps
    value = clips / (ms / 1e3)
    e2e = clips / (ms_e2e / 1e3)
    host = model._host_phase(samples, bucket=model.graph_bucket) if train else None
    small_h2d = B * 2100 * 4
    line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload(cfg, model), "name": cfg_name,
                       "l2": "working set per step (GBs of activations, %.1f MB frame tensor) exceeds the 126 MB L2"
                             % (video_host.numel() * 4 / 1e6),
                       "parallelism": "dp%d" % world, "weights": "seeded random init (device RNG)",
                       "model_build_s": round(build_s, 1)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(video_host.numel() * 4 + small_h2d), "d2h_bytes_per_step": d2h},
            "e2e_uint8_frames": {"value": clips / (ms_u8 / 1e3), "unit": UNIT, "ms_per_step": ms_u8 / args.steps,
                                 "h2d_bytes_per_step": int(video_u8.numel() + small_h2d), "d2h_bytes_per_step": d2h,
                                 "note": "raw uint8 frames, (x/255 - mean)/std fused into mrb_patchify_u8"},
            "gpu_launches": int(launches),
            "tensor_frac_of_step": round(value / world * cfg["flops_per_clip"] / 1e12 / peaks()[0]["bf16_tflops_sustained"], 4)}
    if train:
        line["loss"] = float(last)
        line["config"].update(L_enc=host["Le"], L_dec=host["Ld"])
    if ar_ms is not None:
        line["replicas_without_exchange"] = {"ms_per_step": replica_ms, "steps": n_local,
                                             "what": "each rank's own step time (same step, all-reduce left out, nobody waits): "
                                                     "the N-GPU step runs at the slowest of these"}
        line["grad_allreduce"] = {"ms": ar_ms, "bytes": int(sum(p.numel() for p in trainable) * 4), "share_of_step": ar_ms / (ms / args.steps),
                                  "what": "one in-place NCCL all-reduce (AVG) of the flat fp32 gradient buffer after backward, timed alone "
                                          "(barrier first), median of 3, max over ranks"}

    if rank == 0:
        # ---- roofline of the dominant kernel (gemm2_tcgen05_kernel, the 2-CTA tcgen05 GEMM, plus its 1-CTA sibling for the
        #      small shapes): CUDA events around every C-ABI call of one more (eager) pass, same kernels as the graph replays
        qf_engine = model.engines()[1]
        qf_engine.xattn_events = []
        ops.GEMM_PROFILE = []
        samples["video"] = video_dev
        model.cuda_graphs = False                  # per-launch CUDA events need the eager launch sequence
        vit_engine = model.engines()[0]
        vit_split, vit_engine.split = vit_engine.split, False      # ... and ONE stream in the ViT (the two-stream block loop
        #                                                            overlaps kernels, an event pair would time both halves)
        t5_engine = model.engines()[2]
        dg, t5_engine.decode_graphs = t5_engine.decode_graphs, False
        if train:
            model(samples)["loss"].backward()      # local fwd+bwd only: no collective, the other ranks are not in this step
        else:
            model.generate(samples, num_beams=cfg["beams"], max_length=cfg["max_new_tokens"])
        torch.cuda.synchronize()
        model.cuda_graphs, t5_engine.decode_graphs = True, dg
        vit_engine.split = vit_split
        prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
        xev, qf_engine.xattn_events = qf_engine.xattn_events, None
        x_ms = sum(a.elapsed_time(b) for a, b in xev)
        x_flops = B * cfg["frames"] * 6 * 1.137e9     # 1.137 GF per frame and cross-attention layer
        all_ms = sum(a.elapsed_time(b) for _, _, _, a, b in prof)
        all_flops = sum(2.0 * m * n * k for m, n, k, _, _ in prof)
        prof = [x for x in prof if x[0] >= 512 and x[1] >= 256]      # the launches mrb_gemm routes to gemm2_tcgen05_kernel
        flops = sum(2.0 * m * n * k for m, n, k, _, _ in prof)
        gms = sum(a.elapsed_time(b) for _, _, _, a, b in prof)
        pk, how = peaks()
        ach = flops / (gms / 1e3) / 1e12
        line["roofline"] = {"kernel": "gemm2_tcgen05_kernel", "bound": "tensor", "achieved": ach,
                            "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                            "peak_source": how + " bf16_tflops_sustained (kernel timed inside a long step)",
                            "launches_per_step": len(prof), "flops_per_launch": flops / max(len(prof), 1),
                            "avg_launch_ms": gms / max(len(prof), 1), "share_of_step": gms / (ms / args.steps),
                            "all_gemm_launches": {"ms": all_ms, "tflops": all_flops / (all_ms / 1e3) / 1e12,
                                                  "note": "incl. the 1-CTA kernel's decoder-sized / N=32 launches"},
                            # dram__bytes_read+write of ONE launch of the dominant shape (ViT fc1, M61680 N6144 K1408, bias+GELU)
                            # from profiles/ncu_gemm2_fc1_r02f.csv; its algorithmic bytes (A + B + C once) are 948.9e6
                            "traffic": NCU_FC1_TRAFFIC, "traffic_algorithmic": 948.9e6,
                            "traffic_source": "profiles/ncu_gemm2_fc1_r02f.csv (ncu --set full, one fc1 launch; constant, not "
                                              "measured in this run)"}
        xg_ms = qf_engine.time_xattn_path(B * cfg["frames"], train=train)
        line["qformer_xattn"] = {"what": "Q-Former cross-attention path: batched K/V projection GEMM (6 layers, tcgen05) + 6 attention cores",
                                 "flops_per_step": x_flops, "ms_per_step": xg_ms, "achieved": x_flops / (xg_ms / 1e3) / 1e12,
                                 "unit": "TFLOP/s", "frac": x_flops / (xg_ms / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
                                 "how": "the path's 7 kernels back to back in one CUDA graph (1.14 GB of K/V: larger than L2), 10 replays",
                                 "ms_eager_events": x_ms,
                                 "frac_eager_events": x_flops / (x_ms / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
                                 "eager_events_note": "CUDA events around the same 7 calls of one eager step: also counts the launch "
                                                      "gap in front of each small kernel"}
        if eager_line is not None:
            line["eager_gpu"] = eager_line
        if world == 1 and not args.no_cpu_baseline:
            del model
            gc.collect()
            threads = os.cpu_count() or 1
            cstep, what = cpu_reference_sample(cfg, threads)
            cstep()
            t = cstep()
            line["cpu_baseline"] = {"value": 1.0 / t, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": what + "; 1 warm-up + 1 timed"}
        emit(line)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "eager"])
    ap.add_argument("--config", default="qvh", choices=sorted(CONFIGS), help="BASELINE.json config (default: the QVH headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch-on-GPU arm (eager_gpu key)")
    ap.add_argument("--eval-dropout", action="store_true",
                    help="A/B only: run the training step with dropout rate 0 (NOT the reference's train() step)")
    ap.add_argument("--train-dropout", action="store_true", help="(default since round 2; kept for old command lines)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    _reserve_stdout()
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.impl == "eager":
        run_eager(args, cfg)
    else:
        run_b200(args, args.config, cfg)


if __name__ == "__main__":
    main()
