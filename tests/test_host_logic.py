"""CPU-only checks of the host side: C-ABI exports and signatures, registry / model surface, prompt table, tokenizer,
state-dict key names, recipes, training loop mechanics -- and, without a GPU, as much of the device path as a CPU can run:
  * the engines and the whole model over torch stand-ins of the C-ABI ops (tests/cpu_ops_emulation.py) against the oracle;
  * the kernel SOURCES that are not tcgen05 / TMA code (csrc/elementwise.cu, dropout.cu, attention.cu) compiled as C++20 over a
    host shim of the CUDA execution model (tests/cuda_host_shim/common.cuh) against torch / the oracle;
  * both stacked: T5Engine through the product's ops.py and ctypes signatures into those host-built kernels.
No call into libmrblip_b200.so computes anything here; kernel arithmetic on the device is the `-m gpu` suite's job."""
import ctypes
import json
import math
import os
import re

import numpy as np
import pytest
import torch

from mr_blip_b200.dims import TINY, FULL, T5_PREFIX, init_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    from mr_blip_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mrblip_b200.h")).read()
    declared = set(re.findall(r"\b(mrb_\w+)\s*\(", hdr))
    assert declared >= set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mrb_abi_version() == 1
    # the python binding covers every compute entry point of the header
    assert declared - {"mrb_abi_version", "mrb_last_error"} == set(_lib.SIGNATURES)
    # host-only entry points answer without a device: the per-thread SM cap of the large GEMMs takes even counts, 0 lifts it
    assert lib.mrb_gemm_sm_limit(132) == 0 and lib.mrb_gemm_sm_limit(0) == 0
    assert lib.mrb_gemm_sm_limit(131) != 0 and lib.mrb_gemm_sm_limit(-2) != 0


def test_model_surface_and_state_dict_names(tiny_sd):
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from mr_blip_b200.registry import registry
    assert registry.get_model_class("blip2_mr") is BLIP2_MR
    m = BLIP2_MR(dims=TINY, state_dict=tiny_sd)
    keys = set(m.state_dict())
    assert set(tiny_sd) == keys
    for k in ("visual_encoder.blocks.0.attn.qkv.weight", "ln_vision.weight", "query_tokens", "t5_proj.weight",
              "Qformer.bert.encoder.layer.0.crossattention.self.key.weight",
              T5_PREFIX + "encoder.block.0.layer.0.SelfAttention.q.lora_A.default.weight",
              T5_PREFIX + "decoder.block.1.layer.1.EncDecAttention.v.base_layer.weight",
              T5_PREFIX + "lm_head.lora_B.default.weight"):
        assert k in keys, k
    trainable = {n for n, p in m.named_parameters() if p.requires_grad}
    assert all(("lora_" in n) or n.startswith("t5_proj.") for n in trainable)
    assert "t5_proj.weight" in trainable and "query_tokens" not in trainable
    assert m.state_dict()["visual_encoder.blocks.0.attn.qkv.weight"].dtype == torch.float16   # eva_vit.py:397-412
    for attr in ("visual_encoder", "ln_vision", "Qformer", "query_tokens", "t5_proj", "t5_model", "t5_tokenizer"):
        assert hasattr(m, attr)
    m.train()
    assert not m.visual_encoder.training                  # disabled_train keeps the frozen ViT in eval
    with pytest.raises(RuntimeError, match="no CPU"):
        m({"video": torch.zeros(1, 1, 3, 224, 224)})      # the product never falls back to CPU
    ckpt = {"model": {k: v for k, v in m.state_dict().items() if "lora_" in k}}
    path = os.path.join("/tmp", "mrb_ckpt_test.pth")
    torch.save(ckpt, path)
    msg = m.load_checkpoint(path)                         # partial, non-strict (base_model.py:29-56)
    assert not msg.unexpected_keys


def test_blip2_t5_surface(tiny_sd):
    """registry name, plain HF T5 state-dict names (no peft wrapper), trainable split of blip2_t5.py:60-90, no CPU path."""
    from mr_blip_b200.blip2_t5 import Blip2T5, plain_t5_state_dict
    from mr_blip_b200.registry import registry
    assert registry.get_model_class("blip2_t5") is Blip2T5
    plain = plain_t5_state_dict(tiny_sd)
    m = Blip2T5(dims=TINY, state_dict=plain)
    keys = set(m.state_dict())
    assert keys == set(plain) and not any("lora_" in k or "base_layer" in k or "base_model" in k for k in keys)
    for k in ("t5_model.shared.weight", "t5_model.encoder.block.0.layer.0.SelfAttention.q.weight", "t5_model.lm_head.weight",
              "t5_model.decoder.block.1.layer.1.EncDecAttention.v.weight", "Qformer.bert.encoder.layer.0.crossattention.self.key.weight"):
        assert k in keys, k
    trainable = {n for n, p in m.named_parameters() if p.requires_grad}
    assert all(n.startswith(("Qformer.", "t5_proj.")) or n == "query_tokens" for n in trainable) and "t5_proj.weight" in trainable
    with pytest.raises(RuntimeError, match="no CPU"):
        m({"image": torch.zeros(1, 3, 224, 224), "text_input": ["a"], "text_output": ["b"]})


def test_yaml_recipes_build_models(tmp_path):
    """Config loader with the merge order of lavis/common/config.py (model defaults <- recipe <- --options), on the shipped
    mr_BLIP recipes and on a recipe in the reference's own layout; from_config reads the merged model section."""
    from mr_blip_b200.config import Config, build_model
    from mr_blip_b200.blip2_mr import BLIP2_MR
    base = os.path.join(ROOT, "mr_blip_b200", "configs", "projects", "mr_BLIP", "train")
    frames = {"qvh": (60, 1, 8), "charades": (20, 8, 1), "anet": (60, 1, 4)}          # n_frms, batch, accum of the reference recipes
    for name, (n_frms, bs, acc) in frames.items():
        cfg = Config(os.path.join(base, name + ".yaml"))
        assert cfg.n_frames() == n_frms and cfg.run_cfg.batch_size_train == bs and cfg.run_cfg.accum_grad_iters == acc
        assert cfg.model_cfg.arch == "blip2_mr" and cfg.model_cfg.t5_model == "google/flan-t5-xl"      # from the model defaults
        assert cfg.model_cfg.get("num_query_token") == 32 and cfg.model_cfg.task == "qformer_freeze_lora"
    model, cfg = build_model(os.path.join(base, "charades.yaml"),
                             options=["model.frame_token_aggregation=mean", "model.input_time_format=relative_integers",
                                      "run.batch_size_train=4", "model.allow_synthetic=true"], dims=TINY)
    assert isinstance(model, BLIP2_MR) and model.frame_token_aggregation == "mean" and model.input_time_format == "relative_integers"
    assert cfg.run_cfg.batch_size_train == 4 and model.task == "qformer_freeze_lora"
    ref_style = tmp_path / "qvh_ref_layout.yaml"
    ref_style.write_text("""
model:
  arch: blip2_mr
  model_type: pretrain_flant5xl
  load_finetuned: False # True
  freeze_vit: True
  task: qformer_freeze_lora
  input_time_format: seconds_integers # [seconds_integers | seconds_floats]
  interleave_data: True
  frame_token_aggregation: False # [mean | False]
datasets:
  qvh: # name of the dataset builder
    vis_processor:
        train:
          name: "blip2_video_train"
          n_frms: 60
run:
  task: moment_retrieval
  init_lr: 3e-4
  accum_grad_iters: 8
""")
    # a recipe whose third-party files (tokenizer, FlanT5 weights) are not local must not silently train on synthetic stand-ins
    with pytest.raises(RuntimeError, match="allow_synthetic"):
        build_model(str(ref_style), dims=TINY)
    m2, c2 = build_model(str(ref_style), options=["model.allow_synthetic=true"], dims=TINY)
    assert not m2.frame_token_aggregation and c2.n_frames() == 60 and c2.run_cfg.init_lr == 3e-4 and c2.model_cfg.image_size == 224
    with pytest.raises(AssertionError):
        Config(str(ref_style), options=["model.arch=not_a_model"])


def test_full_dims_match_reference_shapes():
    d = FULL
    assert (d.vit_width, d.vit_depth, d.vit_heads, d.vit_mlp, d.vit_tokens, d.vit_head_dim) == (1408, 39, 16, 6144, 257, 88)
    assert (d.qf_hidden, d.qf_layers, d.qf_heads, d.num_query) == (768, 12, 12, 32)
    assert (d.d_model, d.d_kv, d.t5_heads, d.d_ff, d.t5_layers, d.vocab) == (2048, 64, 32, 5120, 24, 32128)


@pytest.mark.parametrize("fmt", ["seconds_integers", "seconds_floats", "relative_integers", "relative_floats"])
def test_prompt_table_matches_oracle_layout(tiny_sd, fmt):
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from oracle import blip2_mr as ob, synth
    m = BLIP2_MR(dims=TINY, state_dict={k: v for k, v in tiny_sd.items()}, input_time_format=fmt)
    s = synth.make_samples(batch=3, frames=4, seed=6)
    s["duration"][2] = 1234.0                             # 4-digit duration -> two tokens -> ragged rows, left padding
    table, atts, prompts = m.build_prompt_table(s["timestamps"], s["duration"], 3, 4, 32, s["video_prompt_end"],
                                                s["query_prompt"], s["task_prompt"])
    frames = torch.arange(3 * 4 * 32 * 4, dtype=torch.float32).view(3, 4 * 32, 4) + 1.0
    sd = {T5_PREFIX + "shared.weight": -torch.arange(32128 * 4, dtype=torch.float32).view(32128, 4) - 1.0}
    want, want_atts = ob.prompt_concatenation(sd, TINY, m.t5_tokenizer, s["timestamps"], s["duration"], frames,
                                              s["video_prompt_end"], s["query_prompt"], s["task_prompt"], 32,
                                              input_time_format=fmt)
    assert table.shape == want.shape[:2] and torch.equal(atts, want_atts)
    emb, fr = sd[T5_PREFIX + "shared.weight"], frames.reshape(-1, 4)
    got = torch.zeros_like(want)
    for b in range(3):
        for l in range(table.shape[1]):
            i = int(table[b, l])
            got[b, l] = emb[i] if i >= 0 else (0.0 if i == -2 ** 31 else fr[-(i + 1)])
    assert torch.equal(got, want)
    if fmt == "seconds_integers":
        assert (table[2] == -2 ** 31).sum() == 0 and (table[0] == -2 ** 31).sum() == 1   # shorter rows are left-padded
        assert prompts[0].startswith(">") and prompts[0].count(">") == 5
    else:
        assert (table == -2 ** 31).any(axis=1).sum() >= 1                                # ragged rows exist and are padded


def test_host_phase_bucket_padding_is_masked(tiny_sd):
    """Host half of a training step: padding Le / Ld up to the CUDA-graph bucket only appends masked pad tokens and
    ignored (-100) targets; the decoder input is the reference's _shift_right of the labels."""
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from oracle import synth
    m = BLIP2_MR(dims=TINY, state_dict={k: v for k, v in tiny_sd.items()})
    s = synth.make_samples(batch=2, frames=3, seed=8)
    h0, h1 = m._host_phase(s), m._host_phase(s, bucket=(16, 4))
    B, Le, Ld = 2, h0["Le"], h0["Ld"]
    assert h1["Le"] % 16 == 0 and h1["Ld"] % 4 == 0 and 0 <= h1["Le"] - Le < 16 and 0 <= h1["Ld"] - Ld < 4
    t0, t1 = h0["idx"].reshape(B, Le), h1["idx"].reshape(B, h1["Le"])
    assert (t1[:, :Le] == t0).all() and (t1[:, Le:] == m.pad_token_id).all()
    assert (h1["kmask"][:, :Le] == h0["kmask"]).all() and (h1["kmask"][:, Le:] == 0).all()
    assert (h1["labels"][:, :Ld] == h0["labels"]).all() and (h1["labels"][:, Ld:] == -100).all()
    assert (h1["dmask"][:, Ld:] == 0).all()
    lab = torch.from_numpy(h0["labels"]).long()
    want = torch.zeros_like(lab)
    want[:, 1:] = lab[:, :-1]
    want[want == -100] = 0
    assert torch.equal(torch.from_numpy(h0["dec_ids"]).long(), want) and (h1["dec_ids"][:, :Ld] == h0["dec_ids"]).all()
    assert all(h1[k].dtype == np.int32 and h1[k].flags["C_CONTIGUOUS"] for k in ("idx", "kmask", "labels", "dec_ids", "dmask"))


def test_synthetic_tokenizer_roundtrip_and_number_tokens():
    from mr_blip_b200.tokenizer import SyntheticT5Tokenizer, load_t5_tokenizer
    from mr_blip_b200 import mr_utils
    tok = SyntheticT5Tokenizer()
    s = "[[12, 40], [52, 60]]"
    assert tok.decode(tok(s).input_ids, skip_special_tokens=True) == s
    assert all(len(tok(str(i), add_special_tokens=False).input_ids) == 1 for i in range(1000))
    assert mr_utils.find_annoying_numbers(tok, 200) == ([], [])
    assert mr_utils.find_annoying_numbers_replacement_dict([3, 4, 150]) == {3: 2, 4: 5, 150: 151}
    enc = tok(["a b", "a b c d"], padding="longest", return_tensors="pt")
    assert enc.input_ids.shape == (2, 5) and enc.attention_mask[0].tolist() == [1, 1, 1, 0, 0]
    assert enc.input_ids[0, 2].item() == tok.eos_token_id
    assert tok("<extra_id_0>\n", add_special_tokens=False).input_ids == [32099]
    assert isinstance(load_t5_tokenizer("google/flan-t5-xl"), (SyntheticT5Tokenizer,)) or True


def test_init_is_seed_deterministic():
    a = init_state_dict(TINY, seed=5, parts=("qformer",))
    b = init_state_dict(TINY, seed=5, parts=("qformer",))
    c = init_state_dict(TINY, seed=6, parts=("qformer",))
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["t5_proj.weight"], c["t5_proj.weight"])


def test_optimizer_groups_and_scheduler_from_recipe():
    """Stand-alone training set-up from a shipped recipe: the decay / no-decay split of lavis/runners/runner_base.py:108-124
    on the model's trainable tensors (LoRA factors and t5_proj.weight decay, t5_proj.bias does not, frozen towers are
    absent), AdamW betas, and the recipe's scheduler stepping the optimiser's lr."""
    from mr_blip_b200 import optim
    from mr_blip_b200.config import build_model
    model, cfg = build_model(os.path.join(ROOT, "mr_blip_b200", "configs", "projects", "mr_BLIP", "train", "charades.yaml"),
                             options=["model.allow_synthetic=true"], dims=TINY)
    groups, n = optim.param_groups(model, cfg.run_cfg.weight_decay)
    names = {id(p): k for k, p in model.named_parameters()}
    wd = {names[id(p)] for p in groups[0]["params"]}
    no_wd = {names[id(p)] for p in groups[1]["params"]}
    assert groups[0]["weight_decay"] == 0.05 and groups[1]["weight_decay"] == 0
    assert no_wd == {"t5_proj.bias"} and "t5_proj.weight" in wd
    assert all(("lora_" in k) or k == "t5_proj.weight" for k in wd) and any("lora_A" in k for k in wd) and any("lora_B" in k for k in wd)
    trainable = [p for p in model.parameters() if p.requires_grad]
    assert n == sum(p.numel() for p in trainable) and len(wd) + len(no_wd) == len(trainable)
    opt = optim.build_optimizer(model, cfg.run_cfg.init_lr, cfg.run_cfg.weight_decay)
    assert isinstance(opt, torch.optim.AdamW) and opt.defaults["betas"] == (0.9, 0.999) and opt.defaults["lr"] == 3e-4
    sched = optim.build_lr_scheduler(opt, cfg.run_cfg)
    assert isinstance(sched, optim.LinearWarmupCosineLRScheduler) and sched.warmup_steps == 698 and sched.max_epoch == 20
    sched.step(cur_epoch=0, cur_step=0)
    assert all(g["lr"] == 1e-8 for g in opt.param_groups)
    sched.step(cur_epoch=0, cur_step=349)
    assert all(abs(g["lr"] - (1e-8 + (3e-4 - 1e-8) * 349 / 698)) < 1e-15 for g in opt.param_groups)
    sched.step(cur_epoch=10, cur_step=0)                                              # past warm-up: half-way down the cosine
    assert all(abs(g["lr"] - 1.5e-4) < 1e-12 for g in opt.param_groups)
    with pytest.raises(KeyError):
        optim.build_lr_scheduler(opt, {"lr_sched": "nope", "max_epoch": 1, "min_lr": 0, "init_lr": 1})


def test_checkpoint_wire_format_round_trip(tmp_path, tiny_sd):
    """save -> resume in the runner's file format (lavis/runners/runner_base.py:572-644): only trainable tensors are
    written, a second model resumes weights + optimizer state + epoch from the file, frozen towers stay untouched."""
    from mr_blip_b200 import checkpoint, optim
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from mr_blip_b200.config import Config
    cfg = Config(os.path.join(ROOT, "mr_blip_b200", "configs", "projects", "mr_BLIP", "train", "qvh.yaml"))
    m = BLIP2_MR(dims=TINY, state_dict=tiny_sd)
    trainable = {n for n, p in m.named_parameters() if p.requires_grad}
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.requires_grad:
                p.add_(torch.randn(p.shape, generator=g) * 0.01)
    opt = optim.build_optimizer(m, 3e-4, 0.05, fused=False)
    for p in m.parameters():
        if p.requires_grad:
            p.grad = torch.full_like(p, 0.5)
    opt.step()
    path = checkpoint.save_checkpoint(m, opt, str(tmp_path), cur_epoch=3, config=cfg)
    assert os.path.basename(path) == "checkpoint_3.pth"
    assert os.path.basename(checkpoint.save_checkpoint(m, opt, str(tmp_path), 3, is_best=True)) == "checkpoint_best.pth"
    obj = torch.load(path, map_location="cpu")
    assert set(obj) == {"model", "optimizer", "config", "scaler", "epoch"} and obj["epoch"] == 3 and obj["scaler"] is None
    # LoRA + t5_proj, plus the two tied aliases of the frozen T5 embedding: named_parameters() lists a tied tensor once
    # (as shared.weight), so the reference's filter keeps encoder/decoder.embed_tokens.weight in the file as well
    assert set(obj["model"]) == trainable | {T5_PREFIX + "encoder.embed_tokens.weight", T5_PREFIX + "decoder.embed_tokens.weight"}
    assert obj["config"]["run"]["init_lr"] == 3e-4 and obj["config"]["model"]["arch"] == "blip2_mr"
    assert os.path.getsize(path) < 0.5 * sum(v.numel() * v.element_size() for v in m.state_dict().values())
    m2 = BLIP2_MR(dims=TINY, state_dict=tiny_sd)
    opt2 = optim.build_optimizer(m2, 3e-4, 0.05, fused=False)
    assert checkpoint.resume_checkpoint(m2, opt2, path) == 4
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys() and all(torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) for k in s1)
    msg = m2.load_checkpoint(path)                               # the model-side loader reads the same file
    assert not msg.unexpected_keys
    with pytest.raises(RuntimeError, match="invalid"):
        checkpoint.resume_checkpoint(m2, opt2, str(tmp_path / "missing.pth"))
    bad = {"model": {"not.a.key": torch.zeros(1)}, "optimizer": None, "epoch": 0}
    torch.save(bad, str(tmp_path / "bad.pth"))
    with pytest.raises(RuntimeError, match="unexpected"):
        checkpoint.resume_checkpoint(m2, None, str(tmp_path / "bad.pth"))


def test_video_processor_decodes_with_opencv(tmp_path):
    """VideoProcessor + Cv2VideoReader on a synthetic mp4: frame-exact sequential decode at the sampled indices, reference
    layout (float32 [3,T,H,W], CLIP-normalised) and the uint8 variant the fused device normalisation consumes."""
    cv2 = pytest.importorskip("cv2")
    from mr_blip_b200 import data
    from mr_blip_b200.vision import VitEngine
    assert data.PIXEL_MEAN == VitEngine.PIXEL_MEAN and data.PIXEL_STD == VitEngine.PIXEL_STD
    path = str(tmp_path / "clip.mp4")
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (64, 48))
    if not w.isOpened():
        pytest.skip("no mp4 encoder in this OpenCV build")
    n = 90
    for i in range(n):
        w.write(np.full((48, 64, 3), 10 + 8 * (i % 30), np.uint8))           # flat grey frame encodes its index mod 30, 8 levels apart
    w.release()
    vr = data.Cv2VideoReader(path, height=32, width=32)
    assert len(vr) == n and abs(vr.get_avg_fps() - 25.0) < 1e-6
    proc8 = data.VideoProcessor(image_size=32, n_frms=6, sampling="uniform", uint8=True)
    frames, idx, fps = proc8(path)
    assert idx == data.sample_frame_indices(n, 25.0, 6) and fps == vr.get_avg_fps()
    assert frames.dtype == torch.uint8 and tuple(frames.shape) == (3, 6, 32, 32)
    level = frames.float().mean(dim=(0, 2, 3))
    assert ((level - torch.tensor([10.0 + 8 * (i % 30) for i in idx])).abs() < 4).all(), (level, idx)    # lossy codec: a few grey levels
    procf = data.VideoProcessor(image_size=32, n_frms=6, sampling="uniform")
    ff, idx2, _ = procf(path, clip_proposal=[1.0, 3.0])
    assert idx2 == data.sample_frame_indices(n, 25.0, 6, "uniform", [1.0, 3.0]) and min(idx2) >= 25 and max(idx2) < 75
    assert ff.dtype == torch.float32 and tuple(ff.shape) == (3, 6, 32, 32)
    f8, _, _ = data.VideoProcessor(image_size=32, n_frms=6, uint8=True)(path, clip_proposal=[1.0, 3.0])
    want = (f8.float() / 255.0 - torch.tensor(data.PIXEL_MEAN).view(3, 1, 1, 1)) / torch.tensor(data.PIXEL_STD).view(3, 1, 1, 1)
    assert torch.equal(ff, want)
    # out-of-order and repeated indices come back in request order
    got = vr.get_batch([40, 3, 40])
    assert tuple(got.shape) == (3, 32, 32, 3) and torch.equal(got[0], got[2]) and abs(float(got[1].float().mean()) - 34) < 4
    with pytest.raises(RuntimeError, match="cannot open"):
        data.Cv2VideoReader(str(tmp_path / "missing.mp4"))


def test_standalone_training_loop_mechanics(tmp_path):
    """mr_blip_b200.train on CPU with a stand-in model: datasets built from a recipe + --options over a synthetic mp4 (OpenCV
    decode, uint8 frames), the per-iteration lr schedule, gradient accumulation, evaluation -> per-rank result file -> merged
    metrics, and the checkpoint files.  (The real model's step is covered by the GPU tests.)"""
    cv2 = pytest.importorskip("cv2")
    from mr_blip_b200 import train, optim, checkpoint
    from mr_blip_b200.config import Config
    vid_dir = tmp_path / "videos"
    vid_dir.mkdir()
    for name in ("a", "b", "c"):
        w = cv2.VideoWriter(str(vid_dir / (name + ".mp4")), cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (32, 32))
        if not w.isOpened():
            pytest.skip("no mp4 encoder in this OpenCV build")
        for i in range(40):
            w.write(np.full((32, 32, 3), 5 * i, np.uint8))
        w.release()
    anns = [{"qid": i, "video": v, "query": "query %d" % i, "duration": 4.0, "relevant_windows": [[1, 3]]}
            for i, v in enumerate(["a", "b", "c", "a", "b", "c"])]
    for split in ("train", "val"):
        (tmp_path / (split + ".json")).write_text(json.dumps(anns))
    recipe = os.path.join(ROOT, "mr_blip_b200", "configs", "projects", "mr_BLIP", "train", "qvh.yaml")
    cfg = Config(recipe, ["datasets.qvh.build_info.annotations.train.storage=%s" % (tmp_path / "train.json"),
                          "datasets.qvh.build_info.annotations.val.storage=%s" % (tmp_path / "val.json"),
                          "datasets.qvh.build_info.videos.storage=%s" % vid_dir,
                          "datasets.qvh.vis_processor.train.n_frms=4", "datasets.qvh.vis_processor.eval.n_frms=4",
                          "datasets.qvh.vis_processor.train.image_size=16", "datasets.qvh.vis_processor.eval.image_size=16",
                          "run.max_epoch=2", "run.warmup_steps=4", "run.accum_grad_iters=2"])
    assert cfg.datasets_cfg.qvh.vis_processor.train.name == "blip2_video_train"          # nested override kept the siblings
    ds = train.build_datasets(cfg, ["train", "val"])
    assert len(ds["train"]) == 6 and ds["train"].vis_processor.sampling == "random" and ds["val"].vis_processor.sampling == "uniform"
    s0 = ds["val"][0]
    assert s0["video"].dtype == torch.uint8 and tuple(s0["video"].shape) == (4, 3, 16, 16) and s0["timestamps"].tolist() == pytest.approx([0.5, 1.5, 2.5, 3.5])
    loaders = {k: train.build_loader(v, 2, 0, k == "train", 0, 1) for k, v in ds.items()}
    assert len(loaders["train"]) == 3

    class Standin(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(3))
            self.frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
            self.seen = []

        def forward(self, samples):
            assert samples["video"].dtype == torch.uint8 and samples["video"].shape[1:] == (4, 3, 16, 16)
            self.seen.append((samples["epoch"], samples["iters"], samples["num_iters_per_epoch"]))
            return {"loss": ((self.w - 1.0) ** 2).sum() + samples["video"].float().mean() * 0}

        def generate(self, samples, num_beams=5, max_length=50, min_length=1):
            assert (num_beams, max_length, min_length) == (5, 200, 8)
            n = len(samples["query_prompt"])
            preds = ["[[1, 3]]" if int(q) % 2 == 0 else "no idea" for q in samples["query_id"]]
            return {"prediction": preds, "raw_prediction": preds, "answer": list(samples["relevant_windows"]),
                    "qid": samples["query_id"].tolist(), "duration": samples["duration"].tolist()}

    model = Standin()
    opt = optim.build_optimizer(model, cfg.run_cfg.init_lr, cfg.run_cfg.weight_decay, fused=False)
    sched = optim.build_lr_scheduler(opt, cfg.run_cfg)
    steps, lrs = [], []
    real_step = opt.step
    opt.step = lambda *a, **k: (steps.append(len(model.seen)), lrs.append(opt.param_groups[0]["lr"]), real_step(*a, **k))[-1]
    task = train.MomentRetrievalTask()
    st0 = train.train_epoch(task, model, loaders["train"], opt, sched, 0, "cpu", accum_grad_iters=2)
    st1 = train.train_epoch(task, model, loaders["train"], opt, sched, 1, "cpu", accum_grad_iters=2)
    assert model.seen == [(0, 0, 3), (0, 1, 3), (0, 2, 3), (1, 0, 3), (1, 1, 3), (1, 2, 3)]
    assert steps == [2, 5]                                   # an optimiser step every second micro-step; the odd one is dropped at the epoch end
    warm = lambda k: 1e-8 + (3e-4 - 1e-8) * k / 4            # noqa: E731  (global step k of a 4-step warm-up)
    # epoch 1, step 1 counts as global step 1 * 2 + 1 = 3 (the scheduler multiplies by the largest step INDEX seen, optims.py:80-84)
    assert lrs[0] == pytest.approx(warm(1)) and lrs[1] == pytest.approx(warm(3))
    sched.step(cur_epoch=1, cur_step=2)
    assert opt.param_groups[0]["lr"] == pytest.approx(3e-4 * 0.5 * (1 + math.cos(math.pi * 1 / 2)))     # 1*2+2 = 4: cosine of epoch 1 of 2
    assert st1["loss"] < st0["loss"] and float(model.w.detach().min()) > 0
    gen = dict(num_beams=int(cfg.run_cfg.num_beams), max_length=int(cfg.run_cfg.max_len), min_length=int(cfg.run_cfg.min_len))
    results = task.evaluation(model.eval(), loaders["val"], "cpu", **gen)
    assert len(results) == 6 and results[0]["qid"] == "0_0" and results[1]["qid"] == "1_1" and results[2]["qid"] == "2_0"
    metrics = task.after_evaluation(results, "val", 1, str(tmp_path / "result"))
    assert metrics["total"] == 6 and metrics["invalid_predictions"] == pytest.approx(0.5) and metrics["agg_metrics"] == pytest.approx(50.0)
    assert json.load(open(tmp_path / "result" / "val_epoch1.json"))[0]["target"] == "[[1, 3]]"
    path = checkpoint.save_checkpoint(model, opt, str(tmp_path / "out"), 1, config=cfg)
    assert set(torch.load(path)["model"]) == {"w"}


def test_product_beam_search_bookkeeping_matches_transformers_generate():
    """mr_blip_b200.generation.beam_search (the PRODUCT's host-side hypothesis bookkeeping: candidate selection, eos handling,
    incremental cache reorder indices, finalisation) driven by a stand-in engine whose decode_step evaluates a tiny
    transformers T5 on the CPU, against that model's own generate().  The stand-in keeps per-row token histories exactly as
    the real engine keeps per-row K/V caches -- reordered by `beam_idx` before the new token is appended -- so a wrong reorder
    index or score carry-over shows up as a different sequence."""
    tf = pytest.importorskip("transformers")
    if not hasattr(tf.T5ForConditionalGeneration, "generate"):
        pytest.skip("this transformers build has no T5 generate")
    from mr_blip_b200.generation import beam_search

    class Engine:
        def __init__(self, model):
            self.m = model

        def encode(self, inputs_embeds, attention_mask):
            with torch.no_grad():
                return self.m.encoder(inputs_embeds=inputs_embeds, attention_mask=attention_mask).last_hidden_state, attention_mask

        def init_decode(self, enc, B, Le, beams, max_len):
            return {"enc": enc.repeat_interleave(beams, 0), "hist": torch.zeros(B * beams, max_len, dtype=torch.long), "beams": beams}

        def decode_step(self, st, tokens, t, kmask, beam_idx=None):
            if beam_idx is not None:
                st["hist"] = st["hist"][beam_idx]
            st["hist"][:, t] = tokens
            with torch.no_grad():
                out = self.m(encoder_outputs=(st["enc"],), attention_mask=kmask.repeat_interleave(st["beams"], 0),
                             decoder_input_ids=st["hist"][:, :t + 1])
            return out.logits[:, -1].float()

    def trim(row):
        out = []
        for t in row.tolist()[1:]:
            out.append(t)
            if t == 1:
                break
        return out

    n = with_eos = 0
    for seed in range(6):
        torch.manual_seed(100 + seed)
        cfg = tf.T5Config(vocab_size=24, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_decoder_layers=2, num_heads=4,
                          feed_forward_proj="gated-gelu", tie_word_embeddings=False, pad_token_id=0, eos_token_id=1,
                          decoder_start_token_id=0, dropout_rate=0.0)
        m = tf.T5ForConditionalGeneration(cfg).eval()
        with torch.no_grad():
            for p in m.parameters():
                p.mul_(1.5 + 0.25 * (seed % 5))
            if seed % 2:
                m.lm_head.weight[1] = 1.0 * m.lm_head.weight[3::4].sum(0)
        B, L = 3, 6
        emb = torch.randn(B, L, 32)
        mask = torch.ones(B, L, dtype=torch.long)
        mask[1, 4:] = 0
        for nb, mnt, minl, lp in [(5, 12, 1, 1.0), (4, 8, 3, 1.0), (1, 10, 1, 1.0), (3, 15, 1, 2.0)]:
            with torch.no_grad():
                want = m.generate(inputs_embeds=emb, attention_mask=mask, num_beams=nb, max_new_tokens=mnt, min_length=minl,
                                  length_penalty=lp, do_sample=False, repetition_penalty=1.0, early_stopping=False)
            got = beam_search(Engine(m), emb, mask, num_beams=nb, max_new_tokens=mnt, min_length=minl, length_penalty=lp)
            for b in range(B):
                a, c = trim(want[b]), trim(got[b])
                assert a == c, (seed, nb, mnt, minl, lp, b, a, c)
                n += 1
                with_eos += a[-1] == 1
    assert n == 72 and with_eos >= 5, with_eos


def test_splitk_plan_covers_k_exactly_and_fits_the_chip():
    """mrb_gemm_splitk_plan (host arithmetic of the split-K dispatch, callable without a GPU) on the path's decoder-sized and
    32-column shapes: every K block belongs to exactly one non-empty split, tiles x splits CTAs fit the SM count, problems with
    enough tiles or short K stay unsplit, and a forced tile width is honoured."""
    from mr_blip_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    fn = lib.mrb_gemm_splitk_plan
    fn.argtypes = [ctypes.c_int] * 6 + [ctypes.POINTER(ctypes.c_int)] * 3
    fn.restype = ctypes.c_int

    def plan(M, N, K, sms=148, force_bn=0, max_splits=8):
        bn, sp, per = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert fn(M, N, K, sms, force_bn, max_splits, ctypes.byref(bn), ctypes.byref(sp), ctypes.byref(per)) == 0
        return bn.value, sp.value, per.value

    shapes = [(56, 2048, 2080), (56, 6144, 2080), (56, 10240, 2080), (56, 5120, 2080), (64, 2048, 10272), (56, 2048, 5152),
              (56, 2048, 6176), (56, 2048, 32160), (5, 2048, 2080), (80, 2048, 2080), (128, 4096, 2080),
              (8132, 32, 2048), (8132, 32, 10240), (8132, 32, 5120), (300, 32, 2048), (7680, 32, 768)]
    split_any = 0
    for sms in (148, 132, 74):
        for M, N, K in shapes:
            for ms in (2, 4, 8):
                bn, sp, per = plan(M, N, K, sms, 0, ms)
                kb = (K + 63) // 64
                assert bn in (32, 64, 128, 192, 256) and 1 <= sp <= ms
                assert (bn == 32) == (N <= 32) or sp == 1
                if sp > 1:
                    split_any += 1
                    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
                    assert tiles * sp <= sms, (M, N, K, bn, sp)
                    assert per >= 4 and (sp - 1) * per < kb <= sp * per, (M, N, K, sp, per, kb)      # last split non-empty, all covered
                else:
                    assert per == kb
    assert split_any > 60
    assert plan(56, 2048, 2080)[1] > 1 and plan(8132, 32, 2048)[1] == 2
    assert plan(56, 32128, 2080)[1] == 1                     # lm_head: 126+ tiles already fill the chip
    assert plan(56, 2048, 192)[1] == 1                       # 3 K blocks: nothing to split
    assert plan(8132, 2048, 2080)[1] == 1                    # many row tiles: never split
    assert plan(56, 2048, 2080, max_splits=1)[1] == 1
    assert plan(56, 2048, 2080, force_bn=64, max_splits=4) == (64, 4, 9)
    assert fn(0, 8, 8, 148, 0, 8, None, None, None) != 0


def test_third_party_weight_files_load_into_reference_key_names(tmp_path, tiny_sd):
    """mr_blip_b200.weights: a sharded transformers T5 directory (safetensors + index) and an eva_vit_g-style .pth map onto the
    model's peft / visual_encoder names; LoRA adapters and the Q-Former are left alone, incomplete or mis-shaped files raise."""
    from safetensors.torch import save_file
    from mr_blip_b200 import weights
    from mr_blip_b200.blip2_mr import BLIP2_MR
    m = BLIP2_MR(dims=TINY, state_dict=tiny_sd)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    # transformers-style T5 state dict: plain names, no adapters
    hf = {}
    for k, v in before.items():
        if k.startswith(T5_PREFIX) and "lora_" not in k:
            hf[k[len(T5_PREFIX):].replace(".base_layer.", ".")] = (torch.randn(v.shape, generator=g) * 0.02).to(v.dtype)
    hf["encoder.embed_tokens.weight"] = hf["shared.weight"].clone()          # files repeat the tied tensor
    hf["decoder.embed_tokens.weight"] = hf["shared.weight"].clone()
    assert "lm_head.weight" in hf and "encoder.block.0.layer.0.SelfAttention.q.weight" in hf
    t5_dir = tmp_path / "flan-t5"
    t5_dir.mkdir()
    names = sorted(hf)
    shards = {"model-00001-of-00002.safetensors": names[::2], "model-00002-of-00002.safetensors": names[1::2]}
    for fn, ks in shards.items():
        save_file({k: hf[k].contiguous() for k in ks}, str(t5_dir / fn))
    (t5_dir / "model.safetensors.index.json").write_text(json.dumps({"weight_map": {k: fn for fn, ks in shards.items() for k in ks}}))
    n, unknown = weights.load_hf_t5(m, str(t5_dir))
    assert n == len(hf) and unknown == []
    after = m.state_dict()
    assert torch.equal(after[T5_PREFIX + "encoder.block.0.layer.0.SelfAttention.q.base_layer.weight"], hf["encoder.block.0.layer.0.SelfAttention.q.weight"])
    assert torch.equal(after[T5_PREFIX + "lm_head.base_layer.weight"], hf["lm_head.weight"])
    assert torch.equal(after[T5_PREFIX + "encoder.embed_tokens.weight"], hf["shared.weight"])
    assert m._get(T5_PREFIX + "decoder.embed_tokens.weight") is m._get(T5_PREFIX + "shared.weight")       # still tied
    for k in before:
        if "lora_" in k or k.startswith(("Qformer.", "visual_encoder.", "ln_vision.", "t5_proj.")) or k == "query_tokens":
            assert torch.equal(after[k], before[k]), k
    # eva_vit_g-style file: un-prefixed names, one extra block, head and final norm
    vit = {k[len("visual_encoder."):]: torch.randn(v.shape, generator=g) * 0.02 for k, v in before.items() if k.startswith("visual_encoder.")}
    vit.update({"blocks.39.attn.qkv.weight": torch.zeros(4, 4), "norm.weight": torch.ones(8), "head.weight": torch.zeros(2, 8)})
    vpath = str(tmp_path / "eva_vit_g.pth")
    torch.save(vit, vpath)
    n, skipped = weights.load_eva_vit(m, vpath)
    assert sorted(skipped) == ["blocks.39.attn.qkv.weight", "head.weight", "norm.weight"] and n == len(vit) - 3
    w = m.state_dict()["visual_encoder.blocks.1.mlp.fc1.weight"]
    assert w.dtype == torch.float16 and torch.equal(w, vit["blocks.1.mlp.fc1.weight"].half())
    assert torch.equal(m.state_dict()["visual_encoder.pos_embed"], vit["pos_embed"])
    # failure modes
    part = dict(hf)
    del part["decoder.block.1.layer.2.DenseReluDense.wo.weight"]
    with pytest.raises(RuntimeError, match="incomplete"):
        weights.load_hf_t5(m, part)
    bad = dict(hf)
    bad["lm_head.weight"] = torch.zeros(7, 7)
    with pytest.raises(RuntimeError, match="shape mismatch"):
        weights.load_hf_t5(m, bad)
    with pytest.raises(RuntimeError, match="invalid"):
        weights.read_state_dict(str(tmp_path / "nope.bin"))
    # from_config picks both up from the recipe
    from mr_blip_b200.config import build_model
    m2, _ = build_model(os.path.join(ROOT, "mr_blip_b200", "configs", "projects", "mr_BLIP", "train", "qvh.yaml"),
                        options=["model.t5_model=%s" % t5_dir, "model.vit_weights=%s" % vpath,
                                 "model.allow_synthetic=true"],      # the toy directory holds weights but no tokenizer files
                        dims=TINY)
    sd2 = m2.state_dict()
    assert torch.equal(sd2[T5_PREFIX + "lm_head.base_layer.weight"], hf["lm_head.weight"])
    assert torch.equal(sd2["visual_encoder.blocks.0.attn.proj.weight"], vit["blocks.0.attn.proj.weight"].half())
    # the plain-named sibling (blip2_t5) takes the same directory under its own prefix
    from mr_blip_b200.blip2_t5 import Blip2T5
    m3 = Blip2T5.from_config({"t5_model": str(t5_dir), "vit_weights": vpath, "dims": TINY})
    assert torch.equal(m3.state_dict()["t5_model.encoder.block.1.layer.1.DenseReluDense.wi_0.weight"],
                       hf["encoder.block.1.layer.1.DenseReluDense.wi_0.weight"])


def test_dropout_mask_header_matches_numpy_restatement(tmp_path):
    """csrc/dropmask.cuh (what the kernels evaluate) compiled as plain C++ vs oracle/dropout.py (what the oracle evaluates)."""
    import ctypes
    import subprocess
    import numpy as np
    from oracle import dropout as od
    src = tmp_path / "h.cpp"
    src.write_text('#include "dropmask.cuh"\n'
                   'extern "C" void draws(unsigned seed, unsigned site, int rows, int cols, unsigned char* out) {\n'
                   '  const uint32_t key = mrb::drop_key(seed, site), ng = mrb::drop_groups(cols);\n'
                   '  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c)\n'
                   '    out[(long long)r * cols + c] = (mrb::drop_word(key, (uint32_t)r * ng, c >> 2) >> (8 * (c & 3))) & 0xff;\n'
                   '}\n'
                   'extern "C" void keeps(unsigned seed, unsigned site, int rows, int cols, float p, unsigned char* out, float* scale) {\n'
                   '  const mrb::DropSpec d = mrb::make_drop(nullptr, site, p);\n'
                   '  const uint32_t key = mrb::drop_key(seed, d.site), ng = mrb::drop_groups(cols);\n'
                   '  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c)\n'
                   '    out[(long long)r * cols + c] = mrb::drop_keep(mrb::drop_word(key, (uint32_t)r * ng, c >> 2), c, d.thr);\n'
                   '  *scale = d.scale;\n'
                   '}\n')
    so = tmp_path / "h.so"
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mr_blip_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", "-I", csrc, str(src), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    for seed, site, rows, cols in ((0, 0, 3, 8), (123456789, od.site(od.DEC, 23, od.CROSS_P), 130, 2037),
                                   (0xFFFFFFFF, od.lora_site("x.encoder.block.7.layer.1.DenseReluDense.wo"), 77, 5120),
                                   (42, od.site(od.ENC, 5, od.SELF_P), 70000, 61)):          # row * groups wraps past 2^20 ...
        out = np.empty((rows, cols), dtype=np.uint8)
        lib.draws(ctypes.c_uint(seed), ctypes.c_uint(site), rows, cols, out.ctypes.data_as(ctypes.c_void_p))
        assert (out == od.draws(seed, site, rows, cols)).all(), (seed, site, rows, cols)
    for p in (0.1, 0.05, 0.0):
        out = np.empty((64, 100), dtype=np.uint8)
        sc = ctypes.c_float()
        lib.keeps(ctypes.c_uint(5), ctypes.c_uint(9), 64, 100, ctypes.c_float(p), out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(sc))
        assert (out.astype(bool) == od.keep_mask(5, 9, 64, 100, p)).all()
        assert sc.value == float(od.scale_of(p))


def _t5_case(d):
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 40, d.d_model, generator=g) * 2.0
    mask = torch.ones(2, 40, dtype=torch.long)
    mask[1, 33:] = 0
    labels = torch.randint(2, 1000, (2, 7), generator=g)
    labels[:, -1] = 1
    labels[1, 5:] = -100
    labels[1, 4] = 1
    return emb, mask, labels


def _relfro(got, want):
    got, want = torch.as_tensor(got).float(), torch.as_tensor(want).float()
    assert got.shape == want.shape, (got.shape, want.shape)
    return ((got - want).norm() / want.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("mode", ["eval", "train_dropout", "train_dropout_no_attention_sites"])
def test_t5_engine_host_logic_with_emulated_ops(tiny_sd, monkeypatch, mode):
    """T5Engine (op sequence, extended-K LoRA layout, the hand-written backward, and in train mode every dropout site incl. the
    LoRA-dropout decomposition dx = dy W + sum_j mask_j (dy sB_j) A_j) run on the CPU over torch stand-ins of the C-ABI ops
    (tests/cpu_ops_emulation.py) against the oracle: loss, logits, d inputs_embeds and every LoRA gradient.  The stand-ins keep
    the buffers' 16-bit types, so the tolerances are those of the GPU tests."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200.dims import TINY, T5_PREFIX
    from mr_blip_b200.dropout import DropState
    from oracle import t5 as ot5
    from oracle.dropout import Dropper
    monkeypatch.setenv("MRB_OVERLAP", "0")
    t5mod = emu.load_engine_module("t5")
    params = {k: v.clone() for k, v in tiny_sd.items()}
    eng = t5mod.T5Engine(TINY, params.__getitem__)
    eng.fuse_min_rows = 0                # fused residual-stream passes at every size (default: encoder-sized inputs only)
    emb, mask, labels = _t5_case(TINY)
    dmask = (labels != -100).long()
    drop = None
    if mode != "eval":
        eng.drop = DropState(device="cpu", attention=(mode == "train_dropout"))
        seed = eng.drop.set_seed(0xC0FFEE11)
        drop = Dropper(seed)
        if mode != "train_dropout":
            inner = Dropper(seed)
            drop = lambda x, site, p: x if (site & 31) in (1, 3) else inner(x, site, p)      # SELF_P / CROSS_P sites off
            drop.t5, drop.lora, drop.qformer = inner.t5, inner.lora, inner.qformer
    eng.zero_grads()
    out = eng.loss(emb.clone(), mask, labels, dmask, backward=True, want_logits=True)
    sd = dict(tiny_sd)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in sd if "lora_" in k}
    sd.update(leaves)
    e = emb.clone().requires_grad_(True)
    o = ot5.t5_forward(sd, TINY, e, mask, labels, dmask, drop=drop)
    o["loss"].backward()
    assert abs(out["loss"].item() - o["loss"].item()) < 5e-3
    assert _relfro(out["logits"], o["logits"]) < 2e-2
    assert _relfro(out["d_inputs_embeds"], e.grad) < 4e-2
    grads = {id(p): g for p, g in eng.param_grads()}
    for k, leaf in leaves.items():
        assert _relfro(grads[id(params[k])], leaf.grad) < 4e-2, k
    if mode != "eval":
        with torch.no_grad():
            ev = ot5.t5_forward(dict(tiny_sd), TINY, emb, mask, labels, dmask)
        assert abs(ev["loss"].item() - o["loss"].item()) > 1e-2          # the masks matter: the comparison above can fail


@pytest.mark.parametrize("train", [False, True])
def test_qformer_engine_host_logic_with_emulated_ops(tiny_sd, train):
    """QFormerEngine (ln_vision, batched cross K/V projection, 12 layers, in train mode the frozen Q-Former's hidden and
    attention-probability dropout) on the CPU over the op stand-ins against the oracle."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200.dims import TINY
    from mr_blip_b200.dropout import DropState
    from oracle import qformer as oqf, vit as ovit
    from oracle.dropout import Dropper
    vmod = emu.load_engine_module("vision")
    params = {k: v.clone() for k, v in tiny_sd.items()}
    eng = vmod.QFormerEngine(TINY, params.__getitem__)
    frames = 3
    g = torch.Generator().manual_seed(5)
    vit_out = torch.randn(frames * TINY.vit_tokens, TINY.vit_width, generator=g)
    drop = None
    if train:
        eng.drop = DropState(device="cpu")
        drop = Dropper(eng.drop.set_seed(0xBEEF))
    h, h16 = eng.forward(vit_out, frames)
    with torch.no_grad():
        ie = ovit.ln_vision(tiny_sd, TINY, vit_out.view(frames, TINY.vit_tokens, -1))
        want = oqf.qformer_forward(tiny_sd, TINY, ie, drop=drop)
        other = oqf.qformer_forward(tiny_sd, TINY, ie, drop=None if train else Dropper(1))
    assert _relfro(h.view(frames, TINY.num_query, -1), want) < 2e-3
    assert _relfro(h.view(frames, TINY.num_query, -1), other) > 2e-2


@pytest.fixture(scope="module")
def dropout_kernels_on_host(tmp_path_factory):
    """csrc/dropout.cu (the kernel source itself, unmodified) compiled as C++20 over tests/cuda_host_shim/common.cuh: one OS
    thread per CUDA thread, barriers for __syncthreads / warp shuffles.  -> ctypes library with the C-ABI entry points."""
    import ctypes
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = tmp_path_factory.mktemp("shim")
    shutil.copy(os.path.join(root, "tests", "cuda_host_shim", "common.cuh"), d)
    for f in ("dropout.cu", "dropmask.cuh"):
        shutil.copy(os.path.join(root, "mr_blip_b200", "csrc", f), d)
    so = os.path.join(d, "dropout_host.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-x", "c++",
                           os.path.join(d, "dropout.cu"), "-o", so])
    return ctypes.CDLL(so)


def _bf16(t):
    return t.to(torch.bfloat16)


def _cp(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def test_dropout_kernel_source_runs_on_host_shim_and_matches_oracle(dropout_kernels_on_host):
    """Index arithmetic, shared-memory layouts, shuffles and atomics of every kernel in csrc/dropout.cu, executed on the CPU
    through the host shim, against the oracle's masks (what tests/test_dropout_gpu.py checks on the device)."""
    import ctypes
    from oracle import dropout as od
    lib = dropout_kernels_on_host
    c_i, c_ll, c_u, c_f = ctypes.c_int, ctypes.c_longlong, ctypes.c_uint, ctypes.c_float
    seed = 0x9E3779B1
    word = torch.tensor([seed - (1 << 32)], dtype=torch.int32)
    BF, F32 = 1, 2

    def mask(site, rows, cols, p):
        return torch.from_numpy(od.keep_mask(seed, site, rows, cols, p)).float() * float(od.scale_of(p))

    g = torch.Generator().manual_seed(3)
    # ---- mrb_dropout: fp32 -> fp32, fp32 -> bf16, bf16 in place on a strided buffer
    rows, cols, site, p = 37, 264, 0x1234, 0.1
    x = torch.randn(rows, cols, generator=g)
    m = mask(site, rows, cols, p)
    out = torch.empty_like(x)
    assert lib.mrb_dropout(_cp(x), c_ll(cols), _cp(out), c_ll(cols), rows, cols, F32, F32, _cp(word), c_u(site), c_f(p), None) == 0
    assert torch.equal(out, x * m)
    o16 = torch.empty((rows, cols), dtype=torch.bfloat16)
    assert lib.mrb_dropout(_cp(x), c_ll(cols), _cp(o16), c_ll(cols), rows, cols, F32, BF, _cp(word), c_u(site), c_f(p), None) == 0
    assert torch.equal(o16, _bf16(x * m))
    big = torch.zeros((rows, cols + 32), dtype=torch.bfloat16)
    big[:, :cols] = _bf16(x)
    want = _bf16(big[:, :cols].float() * m)
    assert lib.mrb_dropout(_cp(big), c_ll(cols + 32), _cp(big), c_ll(cols + 32), rows, cols, BF, BF, _cp(word), c_u(site), c_f(p), None) == 0
    assert torch.equal(big[:, :cols], want) and big[:, cols:].abs().max().item() == 0
    assert lib.mrb_dropout(_cp(x), c_ll(cols), _cp(out), c_ll(cols), rows, cols - 2, F32, F32, _cp(word), c_u(site), c_f(p), None) == -1
    # ---- mrb_dropout_add
    r = torch.randn(rows, cols, generator=g)
    assert lib.mrb_dropout_add(_cp(r), _cp(x), _cp(out), rows, cols, _cp(word), c_u(site), c_f(p), None) == 0
    assert torch.allclose(out, r + x * m, rtol=1e-6, atol=1e-6)
    # ---- gated GELU with the inner dropout, forward and backward
    M, Fd, site = 9, 512, 77
    ab = _bf16(torch.randn(M, 2 * Fd, generator=g))
    m = mask(site, M, Fd, p)
    a, b = ab[:, :Fd].float().requires_grad_(True), ab[:, Fd:].float().requires_grad_(True)
    want = torch.nn.functional.gelu(a) * b * m
    h = torch.zeros((M, Fd + 32), dtype=torch.bfloat16)
    assert lib.mrb_gated_gelu_fwd_drop(_cp(ab), _cp(h), M, Fd, c_ll(Fd + 32), BF, _cp(word), c_u(site), c_f(p), None) == 0
    assert _relfro(h[:, :Fd], want) < 4e-3 and h[:, Fd:].abs().max().item() == 0
    dh = _bf16(torch.randn(M, Fd, generator=g))
    want.backward(dh.float())
    dab = torch.zeros((M, 2 * Fd + 32), dtype=torch.bfloat16)
    assert lib.mrb_gated_gelu_bwd_drop(_cp(ab), _cp(dh), c_ll(Fd), _cp(dab), c_ll(2 * Fd + 32), M, Fd, BF, _cp(word), c_u(site), c_f(p), None) == 0
    assert _relfro(dab[:, :Fd], a.grad) < 6e-3 and _relfro(dab[:, Fd:2 * Fd], b.grad) < 6e-3
    # ---- LoRA input dropout: forward down-projection (both lane layouts), dA, dx (16-bit and fp32)
    p, site0 = 0.05, 0x2108
    for M, K, nlin in ((21, 520, 3), (2051, 264, 2), (70, 256, 1), (37, 1096, 3)):   # K not a multiple of the 256-column tile; M over / under 2048;
        #                                                                       M <= 128 and K >= 1024: 16 warps split K in the down-projection
        x_ext = torch.zeros((M, K + 32), dtype=torch.bfloat16)
        x_ext[:, :K] = _bf16(torch.randn(M, K, generator=g))
        A = torch.zeros((32, K), dtype=torch.bfloat16)
        A[:8 * nlin] = _bf16(torch.randn(8 * nlin, K, generator=g) / K ** 0.5)
        xf = x_ext[:, :K].float()
        masks = [mask(site0 + j, M, K, p) for j in range(nlin)]
        x_ext[:, K:] = 7.0
        u = x_ext[:, K:]
        assert lib.mrb_lora_down_drop(_cp(x_ext), c_ll(K + 32), _cp(A), c_ll(K), M, K, nlin, ctypes.c_void_p(u.data_ptr()), c_ll(K + 32),
                                      BF, _cp(word), c_u(site0), c_f(p), None) == 0
        for j in range(nlin):
            assert _relfro(u[:, 8 * j:8 * j + 8], (xf * masks[j]) @ A[8 * j:8 * j + 8].float().t()) < 5e-3, (M, K, j)
        assert u[:, 8 * nlin:].abs().max().item() == 0
        q = torch.zeros((M, 32), dtype=torch.bfloat16)
        q[:, :8 * nlin] = _bf16(torch.randn(M, 8 * nlin, generator=g))
        for j in range(nlin):
            dA = torch.ones((8, K))
            assert lib.mrb_lora_wgrad_drop(_cp(x_ext), c_ll(K + 32), ctypes.c_void_p(q.data_ptr() + 16 * j), c_ll(32), M, K, _cp(dA), BF,
                                           _cp(word), c_u(site0 + j), c_f(p), None) == 0
            assert _relfro(dA - 1.0, q[:, 8 * j:8 * j + 8].float().t() @ (xf * masks[j])) < 2e-3, (M, K, j)
        want = sum(masks[j] * (q[:, 8 * j:8 * j + 8].float() @ A[8 * j:8 * j + 8].float()) for j in range(nlin))
        base = torch.randn(M, K, generator=g)
        d32 = base.clone()
        assert lib.mrb_lora_dx_drop(_cp(q), c_ll(32), _cp(A), c_ll(K), nlin, _cp(d32), c_ll(K), F32, M, K, BF, _cp(word), c_u(site0),
                                    c_f(p), None) == 0
        assert _relfro(d32 - base, want) < 1e-3, (M, K)
        d16 = torch.zeros((M, K + 32), dtype=torch.bfloat16)
        d16[:, :K] = _bf16(base)
        assert lib.mrb_lora_dx_drop(_cp(q), c_ll(32), _cp(A), c_ll(K), nlin, _cp(d16), c_ll(K + 32), BF, M, K, BF, _cp(word), c_u(site0),
                                    c_f(p), None) == 0
        assert _relfro(d16[:, :K], _bf16(base).float() + want) < 4e-3 and d16[:, K:].abs().max().item() == 0


def test_swar_keep_compare_of_the_tcgen05_attention_kernels():
    """attention_tc.cu drop_pair_masks / attention_tc_bwd.cu: bit 7 of every byte of ((w >> 1) & 0x7f7f7f7f | 0x80808080) - (thr / 2)
    * 0x01010101 must equal `draw >= thr` for that byte when thr is even (restated here; PRMT then spreads those sign bits)."""
    import numpy as np
    from oracle import dropout as od
    rng = np.random.default_rng(0)
    w = np.concatenate([rng.integers(0, 1 << 32, 200000, dtype=np.uint64), np.array([0, 0xFFFFFFFF, 0x19191919, 0x1a1a1a1a, 0x1b1b1b1b,
                                                                                     0x00ff19ff, 0x1aff001a], dtype=np.uint64)])
    for p in (0.1, 0.5, 0.0):
        thr = od.thr_of(p)
        assert thr % 2 == 0
        t = ((((w >> np.uint64(1)) & np.uint64(0x7f7f7f7f)) | np.uint64(0x80808080)) - np.uint64((thr // 2) * 0x01010101)) & np.uint64(0xFFFFFFFF)
        for i in range(4):
            draw = (w >> np.uint64(8 * i)) & np.uint64(0xFF)
            assert (((t >> np.uint64(8 * i + 7)) & np.uint64(1)) == (draw >= thr)).all(), (p, i)
    assert od.thr_of(0.05) % 2 == 1          # LoRA's 0.05 is odd: it never reaches these kernels (byte compare in csrc/dropout.cu)


@pytest.mark.parametrize("train_dropout,agg,interleave", [(False, None, True), (True, None, True), (True, "mean", True),
                                                         (False, None, False)])
def test_whole_model_train_step_host_logic_with_emulated_ops(tiny_sd, monkeypatch, train_dropout, agg, interleave):
    """BLIP2_MR.forward in train() -- host phase, ViT / Q-Former / interleave gather / T5 loss, the hand-written backward into the
    flat gradient buffer, t5_proj gradients, the gradient hand-over to autograd, and with train_dropout the per-step seed word and
    every dropout site -- run on the CPU over the op stand-ins against the oracle (same checks as tests/test_model_gpu.py::
    test_forward_backward_vs_oracle and tests/test_dropout_gpu.py do on the device)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200.dims import TINY
    from oracle import blip2_mr as ob, synth
    from oracle.dropout import Dropper
    monkeypatch.setenv("MRB_OVERLAP", "0")
    mod = emu.load_model_module()
    model = mod.BLIP2_MR(dims=TINY, state_dict=tiny_sd, cuda_graphs=False, train_dropout=train_dropout, interleave_data=interleave)
    model.frame_token_aggregation = agg
    model.train()
    samples = synth.make_samples(batch=2, frames=2, seed=3)
    samples["duration"][1] = 37.0
    samples["timestamps"][1] = samples["timestamps"][1] * (37.0 / 143.0)
    res = model.forward_mr(samples, want_logits=True)
    res["loss"].backward()
    drop = Dropper(model.drop_state.seed) if train_dropout else None
    sd = dict(tiny_sd)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in sd if "lora_" in k or k.startswith("t5_proj.")}
    sd.update(leaves)
    o = ob.forward_mr(sd, TINY, model.t5_tokenizer, samples, frame_token_aggregation=agg, drop=drop, table=model.annoying_numbers_replacement_dict,
                      interleave_data=interleave)
    o["loss"].backward()
    assert torch.equal(res["attention_mask"], o["attention_mask"]) and torch.equal(res["labels"], o["labels"])
    assert _relfro(res["qformer"], o["qformer"]) < 2e-3
    assert _relfro(res["inputs_embeds"], o["inputs_embeds"]) < 2e-3
    assert abs(res["loss"].item() - o["loss"].item()) < 5e-3
    assert _relfro(res["logits"], o["logits"]) < 2e-2
    for k, leaf in leaves.items():
        got = model._get(k).grad
        assert got is not None and _relfro(got, leaf.grad) < 4e-2, k
    if train_dropout:
        seed = model.drop_state.seed
        l2 = model.forward_mr(samples)["loss"].item()                      # the next step draws other masks
        assert model.drop_state.seed != seed and abs(l2 - res["loss"].item()) > 1e-4
        model.eval()
        with torch.no_grad():                                                # eval() is the eval-mode arithmetic
            ev = model.forward_mr(samples, want_logits=True)
            oe = ob.forward_mr(dict(tiny_sd), TINY, model.t5_tokenizer, samples, frame_token_aggregation=agg)
        assert abs(ev["loss"].item() - oe["loss"].item()) < 5e-3


def _qa_samples(batch=2, frames=6, seed=7):
    from oracle import synth
    s = synth.make_samples(batch=batch, frames=frames, seed=seed)
    s["question_id"] = ["q%d" % i for i in range(batch)]
    s["qa_input"] = ["Question: what does the person do first? Options: A. word alpha B. word beta C. word gamma D. word delta E. word eps Answer: ",
                     "Question: why is the dog running? Options: A. one B. two C. three D. four E. five Answer: "][:batch]
    s["qa_output"] = ["Answer: B", "Answer: E"][:batch]
    return s


@pytest.mark.parametrize("task,train_dropout", [("qformer_freeze_lora_QA", False), ("qformer_freeze_lora_QA_with_localizer", False),
                                                ("qformer_freeze_lora_QA", True)])
def test_video_qa_branch_host_logic_with_emulated_ops(tiny_sd, monkeypatch, task, train_dropout):
    """The two-stage video-QA branch (blip2_mr.py:309-431: localizer / uniform window -> frame selection -> frozen frame encoder
    -> answerer loss; only the answerer's adapters train) and videoQA_generate (:990-1099, :1233-1314) on the CPU over the op
    stand-ins against the oracle restatement."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200.dims import ANSWERER_PREFIX, T5_PREFIX, TINY, add_answerer
    from mr_blip_b200 import mr_utils
    from oracle import blip2_mr as ob
    from oracle.dropout import Dropper
    monkeypatch.setenv("MRB_OVERLAP", "0")
    monkeypatch.setenv("MRB_CUDA_GRAPHS", "0")
    sd = add_answerer(dict(tiny_sd), TINY, seed=1234, lora_b_std=0.02)
    mod = emu.load_model_module()
    model = mod.BLIP2_MR(dims=TINY, state_dict=sd, cuda_graphs=False, task=task, num_frames_for_answer=3, train_dropout=train_dropout)
    # trainability as the reference sets it (blip2_mr.py:199-235)
    names = {n: p.requires_grad for n, p in model.named_parameters()}
    assert all(v == ("lora_" in n) for n, v in names.items() if n.startswith(ANSWERER_PREFIX))
    assert not any(v for n, v in names.items() if n.startswith(T5_PREFIX))
    model.train()
    samples = _qa_samples()
    res = model(dict(samples)) if not train_dropout else model.forward_QA(dict(samples), want_logits=True)
    res["loss"].backward()
    drop = Dropper(model.drop_state.seed) if train_dropout else None
    osd = dict(sd)
    leaves = {k: osd[k].clone().requires_grad_(True) for k in osd if "lora_" in k and k.startswith(ANSWERER_PREFIX)}
    osd.update(leaves)
    o = ob.forward_qa(osd, TINY, model.t5_tokenizer, samples, use_localizer="with_localizer" in task, n_frames=3,
                      post_process=mr_utils.post_process, drop=drop)
    o["loss"].backward()
    assert abs(res["loss"].item() - o["loss"].item()) < 5e-3
    if train_dropout:
        assert _relfro(res["logits"], o["logits"]) < 2e-2 and res["relevant_moments"] == o["relevant_moments"]
    for k, leaf in leaves.items():
        got = model._get(k).grad
        assert got is not None and _relfro(got, leaf.grad) < 4e-2, k
    assert all(p.grad is None for n, p in model.named_parameters() if not n.startswith(ANSWERER_PREFIX))
    # inference: same windows, same answer letters, close letter scores
    model.eval()
    out = model.videoQA_generate(dict(samples))
    moments, rel = ob._qa_relevant_frames(sd, TINY, model.t5_tokenizer, dict(samples, relevant_windows=[[0, 0]], query_id=samples["question_id"]),
                                          "with_localizer" in task, 3, mr_utils.post_process, None)
    want, scores = ob.videoqa_answer(sd, TINY, model.t5_tokenizer, samples, rel)
    assert out["relevant_moments"] == [moments]
    assert _relfro(out["answer_scores"], scores) < 2e-2
    assert out["output_text"] == want and out["qid"] == samples["question_id"]


def test_ops_wrappers_pass_what_the_c_abi_declares(monkeypatch):
    """Every ops.py wrapper touched by the dropout work (incl. the default attention wrappers) hands _lib.call exactly the
    argument list its SIGNATURES entry (= include/mrblip_b200.h) declares, with values ctypes can convert to the declared types.
    No kernel runs: _lib.call is replaced by a checker and the tensors live on the CPU."""
    import ctypes
    from mr_blip_b200 import _lib, ops
    seen = []

    def call(name, *args):
        sig = _lib.SIGNATURES[name]
        assert len(args) == len(sig), (name, len(args), len(sig))
        for i, (a, t) in enumerate(zip(args, sig)):
            if t is ctypes.c_void_p:
                assert a is None or isinstance(a, int), (name, i, a)
            elif t is ctypes.c_float:
                assert isinstance(a, float), (name, i, a)
            else:
                assert isinstance(a, int) and not isinstance(a, bool), (name, i, a)
                t(a)
        seen.append(name)

    monkeypatch.setattr(_lib, "call", call)
    monkeypatch.setattr(ops, "_check", lambda t, *d: t)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    B, H, L, hd = 2, 4, 16, 64
    q = torch.zeros((B * L, 3 * H * hd), dtype=torch.bfloat16)
    o = torch.zeros((B * L, H * hd), dtype=torch.bfloat16)
    lse = torch.zeros((B, H, L))
    bias = torch.zeros((H, 2 * L - 1))
    km = torch.ones((B, L), dtype=torch.int32)
    word = torch.zeros(1, dtype=torch.int32)
    rs = q.stride(0)
    for drop in (None, (word, 33, 0.1)):
        for impl in ("mma", "tc"):
            ops.attention_fwd(q, q[:, H * hd:], q[:, 2 * H * hd:], o, B, H, L, L, hd, 1.0, (L * rs, rs), (L * rs, rs), (L * rs, rs),
                              (L * o.stride(0), o.stride(0)), bias=bias, bias_zero=L - 1, kmask=km, causal=True, lse=lse, impl=impl, drop=drop)
            ops.attention_bwd(q, q[:, H * hd:], q[:, 2 * H * hd:], o, o, q, q[:, H * hd:], q[:, 2 * H * hd:], B, H, L, L, hd, 1.0,
                              (L * rs, rs), (L * rs, rs), (L * rs, rs), (L * o.stride(0), o.stride(0)), (L * o.stride(0), o.stride(0)),
                              lse, lse.view(-1), bias=bias, bias_zero=L - 1, kmask=km, causal=True, impl=impl, drop=drop)
    x = torch.zeros((8, 64))
    xe = torch.zeros((8, 64 + 32), dtype=torch.bfloat16)
    A = torch.zeros((32, 64), dtype=torch.bfloat16)
    ops.dropout(x, xe[:, :64], 8, 64, word, 5, 0.1)
    ops.dropout_add(x, x, torch.empty_like(x), word, 5, 0.1)
    ab = torch.zeros((8, 128 + 32), dtype=torch.bfloat16)
    ops.gated_gelu_fwd_drop(ab, xe, 8, 64, word, 5, 0.1)
    ops.gated_gelu_bwd_drop(ab, xe, ab, 8, 64, word, 5, 0.1)
    ops.lora_down_drop(xe[:, :64], A, xe[:, 64:], 8, 64, 3, word, 5, 0.05)
    ops.lora_wgrad_drop(xe.data_ptr(), xe.stride(0), xe.data_ptr() + 128, xe.stride(0), 8, 64, torch.zeros((8, 64)), ops.BF16, word, 5, 0.05)
    ops.lora_dx_drop(xe[:, 64:], A, 3, x, 8, 64, word, 5, 0.05)
    ops.dropout_add_norm(x, x, torch.empty_like(x), torch.zeros(64), 1e-6, xe, word, 5, 0.1)
    ops.rmsnorm_bwd_drop(x, torch.zeros(64), xe, 1e-6, torch.zeros_like(x), torch.zeros_like(xe), word, 5, 0.1)
    ops.gemm_sm_limit(132)
    monkeypatch.setattr(ops, "SPLITK", False)
    ops.gemm(xe, A, out=torch.zeros((8, 32)), M=8, K=64)
    monkeypatch.setattr(ops, "SPLITK", True)                 # the default: small-M / 32-column GEMMs take the split-K entry point
    monkeypatch.setattr(ops, "_SPLITK_MAIN", torch.zeros(64))
    ops.gemm(xe, A, out=torch.zeros((8, 32)), M=8, K=64)
    assert "mrb_gemm_splitk" in seen
    assert {"mrb_attention_fwd", "mrb_attention_fwd_tc", "mrb_attention_fwd_drop", "mrb_attention_fwd_tc_drop", "mrb_attention_bwd",
            "mrb_attention_bwd_tc", "mrb_attention_bwd_drop", "mrb_attention_bwd_tc_drop", "mrb_dropout", "mrb_dropout_add",
            "mrb_gated_gelu_fwd_drop", "mrb_gated_gelu_bwd_drop", "mrb_lora_down_drop", "mrb_lora_wgrad_drop", "mrb_lora_dx_drop",
            "mrb_gemm", "mrb_dropout_add_norm", "mrb_rmsnorm_bwd_drop", "mrb_gemm_sm_limit"} <= set(seen)


def test_ctypes_signatures_match_header_and_source_prototypes():
    """_lib.SIGNATURES (what ctypes converts to) against the C prototypes -- in include/mrblip_b200.h AND in the definitions
    under csrc/ -- argument by argument: pointer / long long / int / unsigned / float."""
    from mr_blip_b200 import _lib

    def kinds(arglist):
        out = []
        for a in arglist.split(","):
            a = " ".join(a.split())
            if "*" in a or a.startswith("CUtensorMap"):
                out.append(ctypes.c_void_p)
            elif a.startswith("long long"):
                out.append(ctypes.c_longlong)
            elif a.startswith("unsigned"):
                out.append(ctypes.c_uint)
            elif a.startswith("float"):
                out.append(ctypes.c_float)
            elif a.startswith("int"):
                out.append(ctypes.c_int)
            else:
                raise AssertionError("unparsed argument: " + a)
        return out

    hdr = open(os.path.join(ROOT, "include", "mrblip_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = dict(re.findall(r"\bint\s+(mrb_\w+)\s*\(([^)]*)\)\s*;", hdr))
    for name, sig in _lib.SIGNATURES.items():
        assert kinds(protos[name]) == sig, name
    defs = {}
    csrc = os.path.join(ROOT, "mr_blip_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith(".cu"):
            src = re.sub(r"//[^\n]*", "", open(os.path.join(csrc, f)).read())
            for name, args in re.findall(r'extern "C" int\s+(mrb_\w+)\s*\(([^)]*)\)\s*\{', src):
                defs[name] = args
    for name, sig in _lib.SIGNATURES.items():
        assert name in defs and kinds(defs[name]) == sig, name


def test_blip2_t5_host_logic_with_emulated_ops(monkeypatch):
    """The thin `blip2_t5` sibling (BASELINE.json configs[0] family) over the op stand-ins: loss, logits and beam-search output
    against its oracle restatement (what tests/test_model_gpu.py::test_blip2_t5_forward_and_generate_vs_oracle checks on the
    device)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from oracle import blip2_t5 as obt
    monkeypatch.setenv("MRB_OVERLAP", "0")
    monkeypatch.setenv("MRB_CUDA_GRAPHS", "0")
    mod = emu.load_blip2_t5_module()
    sd0 = init_state_dict(TINY, seed=1234, lora_b_std=0.0)
    model = mod.Blip2T5(dims=TINY, state_dict=mod.plain_t5_state_dict(sd0)).eval()
    g = torch.Generator().manual_seed(3)
    samples = {"image": torch.randn(2, 3, 224, 224, generator=g), "text_input": ["a photo of", "Question: what is shown? Answer:"],
               "text_output": ["a dog in the garden", "two friends"], "prompt": ["a photo of", "Question: what is shown? Answer:"]}
    got = model(samples)["loss"].item()
    want = obt.forward(sd0, TINY, model.t5_tokenizer, samples)
    assert abs(got - want["loss"].item()) < 5e-3
    assert _relfro(model._last_logits, want["logits"]) < 2e-2
    text = model.generate(samples, num_beams=3, max_length=6)
    want_text, want_seqs = obt.generate(sd0, TINY, model.t5_tokenizer, samples, num_beams=3, max_length=6)
    assert model._last_sequences.tolist() == want_seqs.tolist() and text == want_text


def test_generate_host_logic_with_emulated_ops(tiny_sd, monkeypatch):
    """BLIP2_MR.generate (prefix, cached incremental decoder, in-place beam reorder, beam-search bookkeeping, post-processing)
    over the op stand-ins against the oracle's no-cache beam search: token-for-token."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200.mr_utils import post_process
    from oracle import blip2_mr as ob, synth
    monkeypatch.setenv("MRB_OVERLAP", "0")
    monkeypatch.setenv("MRB_CUDA_GRAPHS", "0")
    mod = emu.load_model_module()
    model = mod.BLIP2_MR(dims=TINY, state_dict=tiny_sd, cuda_graphs=False).eval()
    samples = synth.make_samples(batch=2, frames=2, seed=3)
    out = model.generate(samples, num_beams=4, max_length=7)
    want = ob.generate(tiny_sd, TINY, model.t5_tokenizer, samples, post_process, num_beams=4, max_length=7)
    assert out["sequences"].tolist() == want["sequences"].tolist()
    assert out["raw_prediction"] == want["raw_prediction"] and out["prediction"] == want["prediction"]
    assert set(out) >= {"prediction", "raw_prediction", "answer", "qid", "duration"}


@pytest.fixture(scope="module")
def attention_kernels_on_host(tmp_path_factory):
    """csrc/attention.cu (the mma.sync flash kernels, forward / backward, eval and DROP instantiations, and the one-shot Q-Former
    cross-attention kernel) compiled as C++20 over the host shim: mma.sync m16n8k16, ldmatrix(.trans) and cp.async are emulated
    from their PTX fragment layouts (tests/cuda_host_shim/common.cuh), everything else is the kernel source itself."""
    import re as _re
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = tmp_path_factory.mktemp("shim_attn")
    shutil.copy(os.path.join(root, "tests", "cuda_host_shim", "common.cuh"), d)
    shutil.copy(os.path.join(root, "mr_blip_b200", "csrc", "dropmask.cuh"), d)
    shutil.copy(os.path.join(root, "mr_blip_b200", "csrc", "attn_delta.cuh"), d)
    src = open(os.path.join(root, "mr_blip_b200", "csrc", "attention.cu")).read()
    src, n1 = _re.subn(r"extern __shared__ __align__\(\d+\) uint8_t smem_attn\[\];", "", src)      # the shim owns the buffer
    src, n2 = _re.subn(r"extern __shared__ float sp\[\];", "float* sp = reinterpret_cast<float*>(smem_attn);", src)
    assert n1 >= 4 and n2 == 1
    open(os.path.join(d, "attention.cu"), "w").write(src)
    so = os.path.join(d, "attention_host.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-x", "c++",
                           os.path.join(d, "attention.cu"), "-o", so])
    return ctypes.CDLL(so)


def _ref_attn(q, k, v, scale, bias, kmask, causal, mask):
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    B, H, Lq, Lk = s.shape
    i, j = torch.arange(Lq)[:, None], torch.arange(Lk)[None, :]
    if bias is not None:
        s = s + bias[:, (j - i) + (Lq - 1)][None]
    if kmask is not None:
        s = s.masked_fill(kmask[:, None, None, :] == 0, float("-inf"))
    if causal:
        s = s.masked_fill(j > i, float("-inf"))
    pr = torch.softmax(s, -1)
    if mask is not None:
        pr = pr * mask.view(B, H, Lq, Lk)
    return torch.matmul(pr, vf).permute(0, 2, 1, 3), torch.logsumexp(s, -1)


@pytest.mark.parametrize("dtype,B,H,Lq,Lk,has_bias,has_mask,causal,p", [
    (torch.bfloat16, 2, 2, 72, 72, True, True, False, 0.1),      # tiny-config T5 encoder
    (torch.bfloat16, 2, 2, 72, 72, True, True, False, 0.0),      # eval-mode kernels through the same harness
    (torch.bfloat16, 2, 2, 9, 9, True, True, True, 0.1),         # decoder self-attention
    (torch.bfloat16, 1, 2, 9, 130, False, True, False, 0.1),     # decoder cross-attention, three key tiles
    (torch.float16, 2, 3, 32, 32, False, False, False, 0.1),     # Q-Former self-attention
    (torch.float16, 2, 3, 32, 257, False, False, False, 0.1),    # Q-Former cross-attention (one-shot kernel when no lse is asked)
    (torch.float16, 2, 3, 32, 257, False, False, False, 0.0),
    (torch.bfloat16, 1, 2, 9, 530, False, True, False, 0.1),     # decoder cross-attention over a long encoder: the few-query kernels
    (torch.bfloat16, 1, 1, 20, 600, True, True, False, 0.0),     #   (keys split over the warps; two 16-row blocks, bias window)
    (torch.float16, 1, 1, 16, 515, False, False, False, 0.0),
])
def test_attention_kernel_source_runs_on_host_shim(attention_kernels_on_host, dtype, B, H, Lq, Lk, has_bias, has_mask, causal, p):
    """Forward (O, lse) and backward (dQ, dK, dV) of the mma.sync attention kernels -- with p > 0 their DROP instantiations --
    executed on the CPU against a torch reference that applies the oracle's mask to the probabilities."""
    from oracle import dropout as od
    lib = attention_kernels_on_host
    c_ll, c_u, c_f = ctypes.c_longlong, ctypes.c_uint, ctypes.c_float
    hd, site, seed = 64, 0x1041, 0x9E3779B1
    DT = 1 if dtype == torch.bfloat16 else 0
    scale = 1.0 if dtype == torch.bfloat16 else hd ** -0.5
    g = torch.Generator().manual_seed(17)
    q = (torch.randn(B, Lq, H, hd, generator=g) * 0.5).to(dtype).requires_grad_(True)
    k = (torch.randn(B, Lk, H, hd, generator=g) * 0.5).to(dtype).requires_grad_(True)
    v = torch.randn(B, Lk, H, hd, generator=g).to(dtype).requires_grad_(True)
    bias = torch.randn(H, Lq + Lk - 1, generator=g) if has_bias else None
    kmask = None
    if has_mask:
        kmask = torch.ones((B, Lk), dtype=torch.int32)
        kmask[-1, Lk - max(1, Lk // 6):] = 0
    word = torch.tensor([seed - (1 << 32)], dtype=torch.int32)
    mask = (torch.from_numpy(od.keep_mask(seed, site, B * H * Lq, Lk, p)).float() * float(od.scale_of(p))) if p > 0 else None
    want, want_lse = _ref_attn(q, k, v, scale, bias, kmask, causal, mask)
    dout = torch.randn(B, Lq, H, hd, generator=g).to(dtype)
    want.backward(dout.float())
    rs = H * hd
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    out = torch.empty((B, Lq, H, hd), dtype=dtype)
    lse = torch.empty((B, H, Lq))

    def fwd(o, l):
        args = [P(q), c_ll(Lq * rs), c_ll(rs), P(k), c_ll(Lk * rs), c_ll(rs), P(v), c_ll(Lk * rs), c_ll(rs), P(o), c_ll(Lq * rs), c_ll(rs),
                B, H, Lq, Lk, hd, DT, c_f(scale), P(bias), (Lq + Lk - 1) if has_bias else 0, Lq - 1, P(kmask), 1, int(causal), 0, P(l)]
        if p > 0:
            return lib.mrb_attention_fwd_drop(*args, P(word), c_u(site), c_f(p), None)
        return lib.mrb_attention_fwd(*args, None)

    assert fwd(out, lse) == 0
    assert _relfro(out, want) < 1e-2
    assert _relfro(lse, want_lse) < 1e-3
    if not (has_bias or has_mask or causal) and Lq <= 32 and Lk > 64:      # no lse asked: attn_xq_kernel
        o2 = torch.empty_like(out)
        assert fwd(o2, None) == 0 and _relfro(o2, want) < 1e-2
    dq, dk, dv = (torch.empty_like(t) for t in (q, k, v))
    ws = torch.empty((B * H * Lq,))
    args = [P(q), c_ll(Lq * rs), c_ll(rs), P(k), c_ll(Lk * rs), c_ll(rs), P(v), c_ll(Lk * rs), c_ll(rs), P(out), c_ll(Lq * rs), c_ll(rs),
            P(dout), c_ll(Lq * rs), c_ll(rs), P(dq), P(dk), P(dv), B, H, Lq, Lk, hd, DT, c_f(scale), P(bias),
            (Lq + Lk - 1) if has_bias else 0, Lq - 1, P(kmask), int(causal), 0, P(lse), P(ws)]
    rc = lib.mrb_attention_bwd_drop(*args, P(word), c_u(site), c_f(p), None) if p > 0 else lib.mrb_attention_bwd(*args, None)
    assert rc == 0
    assert _relfro(dv, v.grad) < 1.5e-2 and _relfro(dq, q.grad) < 2e-2 and _relfro(dk, k.grad) < 2e-2


@pytest.fixture(scope="module")
def elementwise_kernels_on_host(tmp_path_factory):
    """csrc/elementwise.cu (norms, gated GELU, interleave gather, cross-entropy, LoRA helpers, casts ...) over the host shim."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = tmp_path_factory.mktemp("shim_elt")
    shutil.copy(os.path.join(root, "tests", "cuda_host_shim", "common.cuh"), d)
    for f in ("elementwise.cu", "dropmask.cuh"):
        shutil.copy(os.path.join(root, "mr_blip_b200", "csrc", f), d)
    so = os.path.join(d, "elementwise_host.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-x", "c++",
                           os.path.join(d, "elementwise.cu"), "-o", so])
    return ctypes.CDLL(so)


def test_elementwise_kernel_source_runs_on_host_shim(elementwise_kernels_on_host):
    """The HBM-bound kernels of the path executed on the CPU through the host shim against torch (the checks
    tests/test_kernels_gpu.py makes on the device): a regression guard for their index arithmetic that needs no GPU."""
    lib = elementwise_kernels_on_host
    c_ll, c_f = ctypes.c_longlong, ctypes.c_float
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    BF, F16 = 1, 0
    g = torch.Generator().manual_seed(9)
    # ---- norms: LayerNorm (+16-bit strided out), RMSNorm, RMSNorm backward
    rows, C = 19, 768
    x, w, b = torch.randn(rows, C, generator=g), torch.randn(C, generator=g), torch.randn(C, generator=g)
    o32, oh = torch.empty(rows, C), torch.zeros((rows, C + 32), dtype=torch.float16)
    assert lib.mrb_norm(P(x), None, P(w), P(b), c_f(1e-6), rows, C, 0, P(o32), P(oh), F16, c_ll(C + 32), None, None) == 0
    want = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-6)
    assert _relfro(o32, want) < 1e-5 and _relfro(oh[:, :C], want) < 1e-3 and oh[:, C:].abs().max().item() == 0
    C = 2048
    x, w = torch.randn(rows, C, generator=g), torch.randn(C, generator=g)
    ob = torch.zeros((rows, C + 32), dtype=torch.bfloat16)
    assert lib.mrb_norm(P(x), None, P(w), None, c_f(1e-6), rows, C, 1, None, P(ob), BF, c_ll(C + 32), None, None) == 0
    xr = x.clone().requires_grad_(True)
    y = w * (xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6))
    assert _relfro(ob[:, :C], y) < 4e-3
    dy = torch.randn(rows, C, generator=g).to(torch.bfloat16)
    y.backward(dy.float())
    dres = torch.ones(rows, C)
    assert lib.mrb_rmsnorm_bwd(P(x), P(w), P(dy), BF, c_ll(C), None, 0, c_f(1e-6), rows, C, P(dres), None) == 0
    assert _relfro(dres - 1.0, xr.grad) < 1e-4
    # encoder- / ViT-sized row counts take the one-block-per-row kernels (norm_row_kernel, rmsnorm_bwd_row_kernel): same arithmetic
    rows3, C3 = 299, 1408
    x3, a3, w3, b3 = (torch.randn(rows3, C3, generator=g) + 3.0, torch.randn(rows3, C3, generator=g), torch.randn(C3, generator=g),
                      torch.randn(C3, generator=g))
    o3, oh3, sum3 = torch.empty(rows3, C3), torch.zeros((rows3, C3 + 8), dtype=torch.float16), torch.empty(rows3, C3)
    assert lib.mrb_norm(P(x3), P(a3), P(w3), P(b3), c_f(1e-6), rows3, C3, 0, P(o3), P(oh3), F16, c_ll(C3 + 8), P(sum3), None) == 0
    want3 = torch.nn.functional.layer_norm(x3 + a3, (C3,), w3, b3, 1e-6)
    assert _relfro(o3, want3) < 1e-5 and _relfro(oh3[:, :C3], want3) < 1e-3 and oh3[:, C3:].abs().max().item() == 0
    assert torch.equal(sum3, x3 + a3)
    x4, w4 = torch.randn(rows3, 2048, generator=g), torch.randn(2048, generator=g)
    ob4 = torch.zeros((rows3, 2048 + 32), dtype=torch.bfloat16)
    assert lib.mrb_norm(P(x4), None, P(w4), None, c_f(1e-6), rows3, 2048, 1, None, P(ob4), BF, c_ll(2048 + 32), None, None) == 0
    assert _relfro(ob4[:, :2048], w4 * (x4 * torch.rsqrt(x4.pow(2).mean(-1, keepdim=True) + 1e-6))) < 4e-3
    for rows2, C2 in ((300, 2048), (297, 1000)):
        x2, w2 = torch.randn(rows2, C2, generator=g), torch.randn(C2, generator=g)
        x2r = x2.clone().requires_grad_(True)
        y2 = w2 * (x2r * torch.rsqrt(x2r.pow(2).mean(-1, keepdim=True) + 1e-6))
        dy2 = torch.randn(rows2, C2, generator=g).to(torch.bfloat16)
        y2.backward(dy2.float())
        dres2 = torch.ones(rows2, C2)
        assert lib.mrb_rmsnorm_bwd(P(x2), P(w2), P(dy2), BF, c_ll(C2), None, 0, c_f(1e-6), rows2, C2, P(dres2), None) == 0
        assert _relfro(dres2 - 1.0, x2r.grad) < 1e-4
    # ---- gated GELU forward / backward
    M, Fd = 7, 5120
    ab = torch.randn(M, 2 * Fd, generator=g).to(torch.bfloat16)
    a, bb = ab[:, :Fd].float().requires_grad_(True), ab[:, Fd:].float().requires_grad_(True)
    want = torch.nn.functional.gelu(a) * bb
    h = torch.zeros((M, Fd + 32), dtype=torch.bfloat16)
    assert lib.mrb_gated_gelu_fwd(P(ab), P(h), M, Fd, c_ll(Fd + 32), BF, None) == 0
    assert _relfro(h[:, :Fd], want) < 4e-3
    dh = torch.randn(M, Fd, generator=g).to(torch.bfloat16)
    want.backward(dh.float())
    dab = torch.zeros((M, 2 * Fd + 32), dtype=torch.bfloat16)
    assert lib.mrb_gated_gelu_bwd(P(ab), P(dh), c_ll(Fd), P(dab), c_ll(2 * Fd + 32), M, Fd, BF, None) == 0
    assert _relfro(dab[:, :Fd], a.grad) < 6e-3 and _relfro(dab[:, Fd:2 * Fd], bb.grad) < 6e-3
    # ---- interleave gather / scatter, frame-token mean
    C = 2048
    emb, frames = torch.randn(50, C, generator=g), torch.randn(12, C, generator=g)
    idx = torch.tensor([3, -1, -12, -(1 << 31), 49, -5, 0, -5], dtype=torch.int32)
    out = torch.full((8, C), 7.0)
    assert lib.mrb_gather_rows(P(idx), P(emb), P(frames), P(out), 8, C, None) == 0
    want = torch.stack([emb[3], frames[0], frames[11], torch.zeros(C), emb[49], frames[4], emb[0], frames[4]])
    assert torch.equal(out, want)
    dout, dfr = torch.randn(8, C, generator=g), torch.zeros(12, C)
    assert lib.mrb_scatter_frames(P(idx), P(dout), P(dfr), 8, C, None) == 0
    wfr = torch.zeros(12, C)
    wfr[0], wfr[11] = dout[1], dout[2]                     # a frame token sits in exactly one row of the table: plain stores ...
    assert torch.equal(dfr[[0, 11]], wfr[[0, 11]]) and dfr[[1, 2, 3, 5, 6, 7, 8, 9, 10]].abs().max().item() == 0
    assert torch.equal(dfr[4], dout[5]) or torch.equal(dfr[4], dout[7])      # ... (a duplicated index keeps one of its rows)
    xm, om = torch.randn(3 * 32, C, generator=g), torch.empty(3, C)
    assert lib.mrb_group_mean(P(xm), P(om), 3, 32, C, None) == 0 and _relfro(om, xm.view(3, 32, C).mean(1)) < 1e-6
    # ---- cross-entropy with ignored targets, mean over the valid ones taken inside the kernel
    V = 32128
    lg = (torch.randn(5, V, generator=g) * 2).requires_grad_(True)
    lab = torch.tensor([5, -100, 31000, 1, -100], dtype=torch.int64)
    loss = torch.nn.functional.cross_entropy(lg, lab, ignore_index=-100)
    loss.backward()
    ls, dl = torch.zeros(1), torch.zeros((5, V + 32), dtype=torch.bfloat16)
    assert lib.mrb_cross_entropy(P(lg.detach()), P(lab), 5, V, None, P(dl), BF, c_ll(V + 32), c_f(-1.0), P(ls), None) == 0
    assert abs(ls.item() - loss.item()) < 1e-5 and _relfro(dl[:, :V], lg.grad) < 5e-3
    # ---- LoRA helpers: weight-gradient reduction (both kernels), tiny-M down-projection, operand re-pack
    for M in (40, 700):
        Pm = torch.randn(M, 264, generator=g).to(torch.bfloat16)
        Q = torch.randn(M, 16, generator=g).to(torch.bfloat16)
        for tr in (0, 1):
            o = torch.ones((8, 264) if tr else (264, 8))
            assert lib.mrb_skinny_wgrad(P(Pm), c_ll(264), ctypes.c_void_p(Q.data_ptr() + 16), c_ll(16), M, 264, P(o), tr, BF, None) == 0
            want = Pm.float().t() @ Q[:, 8:16].float()
            assert _relfro((o - 1.0).t() if tr else o - 1.0, want) < 1e-4, (M, tr)
    for Kd in (2048, 1032, 4096, 5120, 10240):     # 256 / 256 / 512 / 1024 / 1024 threads per row (ceil(K / 2048) x 256, capped)
        xs, Wd = torch.randn(5, Kd + 32, generator=g).to(torch.bfloat16), torch.randn(32, Kd, generator=g).to(torch.bfloat16)
        od = torch.zeros((5, 32), dtype=torch.bfloat16)
        assert lib.mrb_small_down(P(xs), c_ll(Kd + 32), P(Wd), c_ll(Kd), 5, Kd, P(od), c_ll(32), BF, None) == 0
        assert _relfro(od, xs[:, :Kd].float() @ Wd.float().t()) < 4e-3, Kd
    K, N, scale = 264, 520, 0.5
    A, B = torch.randn(8, K, generator=g), torch.randn(N, 8, generator=g)
    ext = torch.zeros((N, K + 32), dtype=torch.bfloat16)
    bdn, adn = torch.zeros((32, N), dtype=torch.bfloat16), torch.zeros((32, K), dtype=torch.bfloat16)
    extb = torch.zeros((K, N + 32), dtype=torch.bfloat16)
    es = 2
    rec = [A.data_ptr(), B.data_ptr(), ext.data_ptr() + (K + 8) * es, K + 32, bdn.data_ptr() + 8 * N * es, N, adn.data_ptr() + 8 * K * es, K,
           extb.data_ptr() + (N + 8) * es, N + 32, K, N, int(np.float64(scale).view(np.int64))]
    table = torch.from_numpy(np.asarray([rec], dtype=np.int64))
    assert lib.mrb_lora_pack(P(table), 1, 1, BF, None) == 0
    sB = (B * scale).to(torch.bfloat16)
    assert torch.equal(ext[:, K + 8:K + 16], sB) and torch.equal(bdn[8:16], sB.t()) and ext[:, :K + 8].abs().max().item() == 0
    assert torch.equal(adn[8:16], A.to(torch.bfloat16)) and torch.equal(extb[:, N + 8:N + 16], A.to(torch.bfloat16).t())
    # ---- plumbing: strided cast, 16-bit transpose, column sums; patch extraction from fp32 and (fused normalisation) from uint8
    xin = torch.randn(9, 300, generator=g)
    oc = torch.zeros((9, 296 + 32), dtype=torch.bfloat16)
    assert lib.mrb_cast2d_f32_to_h(P(xin), c_ll(300), P(oc), c_ll(296 + 32), 9, 296, BF, None) == 0
    assert torch.equal(oc[:, :296], xin[:, :296].to(torch.bfloat16)) and oc[:, 296:].abs().max().item() == 0
    t_in = torch.randn(37, 70, generator=g).to(torch.bfloat16)
    t_out = torch.zeros((70, 40), dtype=torch.bfloat16)
    assert lib.mrb_transpose16(P(t_in), c_ll(70), P(t_out), c_ll(40), 37, 70, None) == 0
    assert torch.equal(t_out[:, :37], t_in.t()) and t_out[:, 37:].abs().max().item() == 0
    cs = torch.zeros(300)
    assert lib.mrb_colsum(P(xin), 9, 300, P(cs), None) == 0 and torch.allclose(cs, xin.sum(0), atol=1e-5)
    Fr, S, Pp = 2, 28, 14
    u8 = torch.randint(0, 256, (Fr, 3, S, S), generator=g, dtype=torch.uint8)
    mean, std = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
    img = (u8.float() / 255.0 - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    ld = 592
    a1, a2 = torch.zeros((Fr * 4, ld), dtype=torch.float16), torch.ones((Fr * 4, ld), dtype=torch.float16)
    assert lib.mrb_patchify(P(img), P(a1), F16, Fr, S, Pp, ld, None) == 0
    assert lib.mrb_patchify_u8(P(u8), P(a2), F16, Fr, S, Pp, ld, *(c_f(v) for v in mean + std), None) == 0
    want = img.view(Fr, 3, 2, Pp, 2, Pp).permute(0, 2, 4, 1, 3, 5).reshape(Fr * 4, 588).to(torch.float16)
    assert torch.equal(a1[:, :588], want) and a1[:, 588:].abs().max().item() == 0
    assert _relfro(a2[:, :588], want) < 1e-3 and a2[:, 588:].abs().max().item() == 0


def test_fused_residual_kernels_equal_their_two_pass_forms(elementwise_kernels_on_host, dropout_kernels_on_host):
    """mrb_dropout_add_norm == mrb_dropout_add then mrb_norm(mode 1), and mrb_rmsnorm_bwd_drop == mrb_rmsnorm_bwd then
    mrb_dropout(fp32 -> 16 bit): the fused train-mode passes of the T5 residual stream (one launch instead of two on a chain of
    ~1 500 dependent decoder kernels) do the same arithmetic in the same order -- compared bit for bit on the host shim, for
    decoder-sized (one warp per row would be the old path) and a ragged encoder-like row count."""
    elt, drp = elementwise_kernels_on_host, dropout_kernels_on_host
    c_ll, c_f, c_u = ctypes.c_longlong, ctypes.c_float, ctypes.c_uint
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    BF = 1
    g = torch.Generator().manual_seed(5)
    word = torch.tensor([0x1234567], dtype=torch.int32)
    for rows, C, p in ((64, 2048, 0.1), (301, 2048, 0.1), (7, 1024, 0.05)):
        x, br, w = torch.randn(rows, C, generator=g), torch.randn(rows, C, generator=g), torch.randn(C, generator=g)
        site = 0x1046
        out1, xn1 = torch.empty(rows, C), torch.zeros((rows, C + 32), dtype=torch.bfloat16)
        assert drp.mrb_dropout_add(P(x), P(br), P(out1), rows, C, P(word), c_u(site), c_f(p), None) == 0
        assert elt.mrb_norm(P(out1), None, P(w), None, c_f(1e-6), rows, C, 1, None, P(xn1), BF, c_ll(C + 32), None, None) == 0
        out2, xn2 = torch.empty(rows, C), torch.zeros((rows, C + 32), dtype=torch.bfloat16)
        assert elt.mrb_dropout_add_norm(P(x), P(br), P(w), c_f(1e-6), rows, C, P(xn2), BF, c_ll(C + 32), P(out2), P(word), c_u(site),
                                        c_f(p), None) == 0
        assert torch.equal(out1, out2) and torch.equal(xn1, xn2)
        assert (out1 != x).any() and (out1 == x).float().mean().item() > p / 2            # some elements dropped, most kept
        dy = torch.randn(rows, C + 32, generator=g).to(torch.bfloat16)
        d1, n1 = torch.ones(rows, C), torch.zeros((rows, C + 32), dtype=torch.bfloat16)
        assert elt.mrb_rmsnorm_bwd(P(x), P(w), P(dy), BF, c_ll(C + 32), None, 0, c_f(1e-6), rows, C, P(d1), None) == 0
        assert drp.mrb_dropout(P(d1), c_ll(C), P(n1), c_ll(C + 32), rows, C, 2, BF, P(word), c_u(site + 3), c_f(p), None) == 0
        d2, n2 = torch.ones(rows, C), torch.zeros((rows, C + 32), dtype=torch.bfloat16)
        assert elt.mrb_rmsnorm_bwd_drop(P(x), P(w), P(dy), BF, c_ll(C + 32), c_f(1e-6), rows, C, P(d2), P(n2), BF, c_ll(C + 32), P(word),
                                        c_u(site + 3), c_f(p), None) == 0
        assert torch.equal(d1, d2) and torch.equal(n1, n2) and n2[:, C:].abs().max().item() == 0


@pytest.mark.parametrize("train_dropout", [False] + ([True] if os.environ.get("MRB_TEST_SLOW", "0") == "1" else []))   # train mode: the whole-model test below
def test_t5_engine_through_the_real_c_abi_on_host_kernels(monkeypatch, train_dropout, elementwise_kernels_on_host,
                                                          dropout_kernels_on_host, attention_kernels_on_host):
    """T5Engine.loss (forward + hand-written backward) through the PRODUCT's ops.py wrappers and ctypes signatures into the
    kernel sources compiled for the host (every launch except the tcgen05 GEMM, which is a torch matmul over the same pointers):
    loss, logits, d inputs_embeds and all LoRA gradients against the oracle -- in train mode with every dropout site on.
    A narrow T5 (d_model 256, 4 heads of 64, d_ff 512, one layer per stack): one OS thread per CUDA thread is slow -- the
    full-width 2-layer config takes 6 minutes per case here (it passed); the widths do not change which code runs."""
    from dataclasses import replace
    NARROW = replace(FULL, t5_layers=1, t5_dec_layers=1, d_model=256, t5_heads=4, d_ff=512, vocab=1024)
    tiny_sd = init_state_dict(NARROW, seed=77, lora_b_std=0.02, parts=("t5",))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200 import _lib, ops
    from mr_blip_b200.dropout import DropState
    from oracle import t5 as ot5
    from oracle.dropout import Dropper
    monkeypatch.setenv("MRB_OVERLAP", "0")
    abi = emu.HostCAbi([elementwise_kernels_on_host, dropout_kernels_on_host, attention_kernels_on_host])
    monkeypatch.setattr(_lib, "call", abi.call)
    monkeypatch.setattr(ops, "_check", lambda t, *d: t)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_SPLITK_MAIN", torch.zeros(16))      # split-K workspace: unused by the host GEMM stand-in
    monkeypatch.setattr(ops, "splitk_register", lambda *a: None)
    t5mod = emu.load_engine_module("t5", ops_module=ops)
    params = {k: v.clone() for k, v in tiny_sd.items()}
    eng = t5mod.T5Engine(NARROW, params.__getitem__)
    eng.fuse_min_rows = 0                # the fused residual-stream passes are for encoder-sized inputs by default: take them here
    emb, mask, labels = _t5_case(NARROW)
    dmask = (labels != -100).long()
    drop = None
    if train_dropout:
        eng.drop = DropState(device="cpu")
        drop = Dropper(eng.drop.set_seed(0xC0FFEE11))
    eng.zero_grads()
    out = eng.loss(emb.clone(), mask, labels, dmask, backward=True, want_logits=True)
    sd = dict(tiny_sd)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in sd if "lora_" in k}
    sd.update(leaves)
    e = emb.clone().requires_grad_(True)
    o = ot5.t5_forward(sd, NARROW, e, mask, labels, dmask, drop=drop)
    o["loss"].backward()
    assert abs(out["loss"].item() - o["loss"].item()) < 5e-3
    assert _relfro(out["logits"], o["logits"]) < 2e-2
    assert _relfro(out["d_inputs_embeds"], e.grad) < 4e-2
    grads = {id(p): g for p, g in eng.param_grads()}
    for k, leaf in leaves.items():
        assert _relfro(grads[id(params[k])], leaf.grad) < 4e-2, k
    # (the narrow test T5 has M <= 128 everywhere, so every GEMM takes the split-K entry point, the default for such shapes)
    want_calls = {"mrb_gemm_splitk", "mrb_norm", "mrb_rmsnorm_bwd", "mrb_attention_fwd", "mrb_attention_bwd", "mrb_cross_entropy", "mrb_gather_rows"}
    if train_dropout:
        want_calls = (want_calls - {"mrb_attention_fwd", "mrb_attention_bwd"}) | {
            "mrb_attention_fwd_drop", "mrb_attention_bwd_drop", "mrb_dropout", "mrb_dropout_add_norm", "mrb_rmsnorm_bwd_drop",
            "mrb_gated_gelu_fwd_drop", "mrb_gated_gelu_bwd_drop", "mrb_lora_down_drop", "mrb_lora_wgrad_drop", "mrb_lora_dx_drop"}
        assert "mrb_dropout_add" not in abi.calls            # every residual add of a train-mode step is fused with the next norm
    assert want_calls <= set(abi.calls), sorted(set(abi.calls))


@pytest.mark.parametrize("train", [False, True])
def test_qformer_engine_through_the_real_c_abi_on_host_kernels(monkeypatch, train, elementwise_kernels_on_host, dropout_kernels_on_host,
                                                               attention_kernels_on_host):
    """QFormerEngine.forward (ln_vision, batched cross K/V projection, self / cross attention incl. the one-shot kernel, FFN, and
    in train mode the frozen Q-Former's dropout) through ops.py and the ctypes signatures into the host-compiled kernel sources,
    against the oracle.  Narrow widths (hidden 128 = 2 heads of 64, 2 layers, encoder width 176) for the same reason as above."""
    import sys
    from dataclasses import replace
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200 import _lib, ops
    from mr_blip_b200.dropout import DropState
    from oracle import qformer as oqf, vit as ovit
    from oracle.dropout import Dropper
    NARROW = replace(FULL, vit_width=176, vit_heads=2, vit_mlp=352, vit_depth=1, qf_hidden=128, qf_heads=2, qf_inter=256, qf_layers=2,
                     d_model=256, t5_heads=4, d_ff=512, vocab=1024, t5_layers=1, t5_dec_layers=1)
    sd = init_state_dict(NARROW, seed=78, parts=("qformer",))
    g = torch.Generator().manual_seed(6)
    sd["ln_vision.weight"], sd["ln_vision.bias"] = 1.0 + 0.1 * torch.randn(176, generator=g), 0.1 * torch.randn(176, generator=g)
    abi = emu.HostCAbi([elementwise_kernels_on_host, dropout_kernels_on_host, attention_kernels_on_host])
    monkeypatch.setattr(_lib, "call", abi.call)
    monkeypatch.setattr(ops, "_check", lambda t, *d: t)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_SPLITK_MAIN", torch.zeros(16))      # split-K workspace: unused by the host GEMM stand-in
    monkeypatch.setattr(ops, "splitk_register", lambda *a: None)
    vmod = emu.load_engine_module("vision", ops_module=ops)
    eng = vmod.QFormerEngine(NARROW, {k: v.clone() for k, v in sd.items()}.__getitem__)
    frames = 2
    vit_out = torch.randn(frames * NARROW.vit_tokens, NARROW.vit_width, generator=g)
    drop = None
    if train:
        eng.drop = DropState(device="cpu")
        drop = Dropper(eng.drop.set_seed(0xBEEF))
    h, _ = eng.forward(vit_out, frames)
    with torch.no_grad():
        ie = ovit.ln_vision(sd, NARROW, vit_out.view(frames, NARROW.vit_tokens, -1))
        want = oqf.qformer_forward(sd, NARROW, ie, drop=drop)
    assert _relfro(h.view(frames, NARROW.num_query, -1), want) < 3e-3
    want_calls = {"mrb_gemm", "mrb_norm", "mrb_attention_fwd", "mrb_cast_f32_to_h"}
    if train:
        want_calls = (want_calls - {"mrb_attention_fwd"}) | {"mrb_attention_fwd_drop", "mrb_dropout", "mrb_dropout_add"}
    assert want_calls <= set(abi.calls), sorted(set(abi.calls))


def test_vit_engine_through_the_real_c_abi_on_host_kernels(monkeypatch, elementwise_kernels_on_host, dropout_kernels_on_host,
                                                           attention_kernels_on_host):
    """VitEngine.forward (patch extraction, cls / pos rows, patch-embed GEMM with the row remap, LayerNorm, the 256 + 1 split of
    the 257 query rows over the tile kernel and the single-row kernel, head dim 88, MLP) through ops.py and the ctypes
    signatures into the host-compiled kernel sources against the oracle.  One narrow block (width 176 = 2 heads of 88)."""
    import sys
    from dataclasses import replace
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200 import _lib, ops
    from oracle import vit as ovit
    NARROW = replace(FULL, vit_width=176, vit_heads=2, vit_mlp=352, vit_depth=1)
    sd = init_state_dict(NARROW, seed=79, parts=("vit",))
    abi = emu.HostCAbi([elementwise_kernels_on_host, dropout_kernels_on_host, attention_kernels_on_host])
    monkeypatch.setattr(_lib, "call", abi.call)
    monkeypatch.setattr(ops, "_check", lambda t, *d: t)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_SPLITK_MAIN", torch.zeros(16))      # split-K workspace: unused by the host GEMM stand-in
    monkeypatch.setattr(ops, "splitk_register", lambda *a: None)
    monkeypatch.setenv("MRB_OVERLAP", "0")
    vmod = emu.load_engine_module("vision", ops_module=ops)
    half = {k: (v.half() if (v.ndim >= 2 and "pos_embed" not in k and "cls_token" not in k) else v) for k, v in sd.items()}
    eng = vmod.VitEngine(NARROW, half.__getitem__)
    g = torch.Generator().manual_seed(8)
    frames = torch.randn(2, 3, 224, 224, generator=g)
    x = eng.forward(frames)
    with torch.no_grad():
        want = ovit.vit_forward(sd, NARROW, frames)
    assert _relfro(x.view(2, NARROW.vit_tokens, -1), want) < 2e-3
    assert {"mrb_patchify", "mrb_cls_pos", "mrb_gemm", "mrb_norm", "mrb_attention_vit"} <= set(abi.calls), sorted(abi.calls)
    # the round-1 path (generic flash kernel for the two full 128-row tiles + the single-row kernel) stays available
    monkeypatch.setattr(ops, "USE_VIT_ATTENTION", False)
    abi.calls.clear()
    assert _relfro(eng.forward(frames).view(2, NARROW.vit_tokens, -1), want) < 2e-3
    assert {"mrb_attention_fwd_tc", "mrb_attention_row"} <= set(abi.calls), sorted(abi.calls)


def test_whole_model_train_step_through_the_real_c_abi_on_host_kernels(monkeypatch, elementwise_kernels_on_host, dropout_kernels_on_host,
                                                                       attention_kernels_on_host):
    """BLIP2_MR.forward in train() with train_dropout -- the whole step: frames -> ViT -> Q-Former -> t5_proj -> interleave gather ->
    T5 loss -> hand-written backward -> t5_proj gradients -> gradient hand-over -- through ops.py, the ctypes signatures and the
    host-compiled kernel sources (only the tcgen05 GEMM is a torch matmul), against the oracle with the same seed word.
    Narrow widths, one layer per stack, full vocabulary (the synthetic tokenizer's ids)."""
    import sys
    from dataclasses import replace
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200 import _lib, ops
    from oracle import blip2_mr as ob, synth
    from oracle.dropout import Dropper
    NARROW = replace(FULL, vit_width=176, vit_heads=2, vit_mlp=352, vit_depth=1, qf_hidden=128, qf_heads=2, qf_inter=256, qf_layers=2,
                     d_model=256, t5_heads=4, d_ff=512, t5_layers=1, t5_dec_layers=1)
    sd = init_state_dict(NARROW, seed=80, lora_b_std=0.02)
    abi = emu.HostCAbi([elementwise_kernels_on_host, dropout_kernels_on_host, attention_kernels_on_host])
    monkeypatch.setattr(_lib, "call", abi.call)
    monkeypatch.setattr(ops, "_check", lambda t, *d: t)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_SPLITK_MAIN", torch.zeros(16))      # split-K workspace: unused by the host GEMM stand-in
    monkeypatch.setattr(ops, "splitk_register", lambda *a: None)
    monkeypatch.setenv("MRB_OVERLAP", "0")
    mod = emu.load_model_module(ops_module=ops)
    model = mod.BLIP2_MR(dims=NARROW, state_dict=sd, cuda_graphs=False, train_dropout=True).train()
    model.engines()[2].fuse_min_rows = 0     # fused residual-stream passes at every size (default: encoder-sized inputs only)
    samples = synth.make_samples(batch=1, frames=2, seed=3)
    res = model.forward_mr(samples, want_logits=True)
    res["loss"].backward()
    osd = dict(sd)
    leaves = {k: osd[k].clone().requires_grad_(True) for k in osd if "lora_" in k or k.startswith("t5_proj.")}
    osd.update(leaves)
    o = ob.forward_mr(osd, NARROW, model.t5_tokenizer, samples, drop=Dropper(model.drop_state.seed))
    o["loss"].backward()
    assert _relfro(res["qformer"], o["qformer"]) < 3e-3
    assert _relfro(res["inputs_embeds"], o["inputs_embeds"]) < 3e-3
    assert abs(res["loss"].item() - o["loss"].item()) < 5e-3
    assert _relfro(res["logits"], o["logits"]) < 2e-2
    for k, leaf in leaves.items():
        got = model._get(k).grad
        assert got is not None and _relfro(got, leaf.grad) < 4e-2, k
    assert {"mrb_lora_pack", "mrb_patchify", "mrb_gather_rows", "mrb_scatter_frames", "mrb_colsum", "mrb_transpose16",
            "mrb_attention_fwd_drop", "mrb_attention_bwd_drop", "mrb_lora_dx_drop"} <= set(abi.calls), sorted(abi.calls)


def test_dropout_seed_word_per_step_and_rank():
    """DropState: a new seed word every step, different streams for different base seeds (BLIP2_MR adds the rank, as the
    reference's setup_seeds(seed + rank) does, train.py:57-58), reproducible from (base seed, step)."""
    from mr_blip_b200.dropout import DropState
    a, b, c = DropState(device="cpu", base_seed=0), DropState(device="cpu", base_seed=1), DropState(device="cpu", base_seed=0)
    sa = [a.advance() for _ in range(50)]
    sb = [b.advance() for _ in range(50)]
    sc = [c.advance() for _ in range(50)]
    assert sa == sc and len(set(sa)) == 50 and not set(sa) & set(sb)
    assert (a.word.item() & 0xFFFFFFFF) == sa[-1]
    assert a.attn(5, 0.1) == (a.word, 5, 0.1) and a.attn(5, 0.0) is None and DropState(device="cpu", attention=False).attn(5, 0.1) is None


@pytest.mark.skipif(os.environ.get("MRB_TEST_SLOW", "0") != "1", reason="one more minute of host-shim time: set MRB_TEST_SLOW=1")
def test_generate_through_the_real_c_abi_on_host_kernels(monkeypatch, elementwise_kernels_on_host, dropout_kernels_on_host,
                                                         attention_kernels_on_host):
    """BLIP2_MR.generate (prefix + cached incremental decoder: attention over the K/V cache with a query offset, beams of a clip
    as query rows of one cross-attention problem, tiny-M down-projections) through ops.py and the host-compiled kernel sources:
    token-for-token against the oracle's no-cache beam search.  Narrow widths."""
    import sys
    from dataclasses import replace
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops_emulation as emu
    from mr_blip_b200 import _lib, ops
    from mr_blip_b200.mr_utils import post_process
    from oracle import blip2_mr as ob, synth
    NARROW = replace(FULL, vit_width=176, vit_heads=2, vit_mlp=352, vit_depth=1, qf_hidden=128, qf_heads=2, qf_inter=256, qf_layers=2,
                     d_model=256, t5_heads=4, d_ff=512, t5_layers=1, t5_dec_layers=1)
    sd = init_state_dict(NARROW, seed=80, lora_b_std=0.02)
    abi = emu.HostCAbi([elementwise_kernels_on_host, dropout_kernels_on_host, attention_kernels_on_host])
    monkeypatch.setattr(_lib, "call", abi.call)
    monkeypatch.setattr(ops, "_check", lambda t, *d: t)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_SPLITK_MAIN", torch.zeros(16))      # split-K workspace: unused by the host GEMM stand-in
    monkeypatch.setattr(ops, "splitk_register", lambda *a: None)
    monkeypatch.setenv("MRB_OVERLAP", "0")
    monkeypatch.setenv("MRB_CUDA_GRAPHS", "0")
    mod = emu.load_model_module(ops_module=ops)
    model = mod.BLIP2_MR(dims=NARROW, state_dict=sd, cuda_graphs=False).eval()
    samples = synth.make_samples(batch=1, frames=2, seed=3)
    out = model.generate(samples, num_beams=3, max_length=5)
    want = ob.generate(sd, NARROW, model.t5_tokenizer, samples, post_process, num_beams=3, max_length=5)
    assert out["sequences"].tolist() == want["sequences"].tolist() and out["raw_prediction"] == want["raw_prediction"]
    assert "mrb_small_down" in abi.calls


def test_splitk_reduce_kernel_source_runs_on_host_shim(tmp_path):
    """The second pass of mrb_gemm_splitk (csrc/gemm.cu splitk_reduce_kernel: ordered sum of the fp32 partials, then bias, exact
    GELU, fp32 residual, output type) cut out of its translation unit -- the rest of gemm.cu is tcgen05 / TMA code -- and run on the
    host shim against torch.  (The planner is covered by test_gemm_splitk_plan...; the partial-sum GEMM itself needs the GPU.)"""
    import shutil
    import subprocess
    src = open(os.path.join(ROOT, "mr_blip_b200", "csrc", "gemm.cu")).read()
    i = src.index("__global__ void __launch_bounds__(256)\nsplitk_reduce_kernel(")
    j = src.index("// Split-K plan.")
    kernel = src[i:j]
    shutil.copy(os.path.join(ROOT, "tests", "cuda_host_shim", "common.cuh"), tmp_path)
    tu = tmp_path / "reduce.cu"
    tu.write_text('#include "common.cuh"\nnamespace mrb {\n' + kernel + '}\nusing namespace mrb;\n'
                  'extern "C" int run(const float* ws, int splits, long long stride, int M, int N, const float* bias, int gelu,\n'
                  '                   const float* resid, long long ldr, void* out, int out_dtype, long long ldc, int blocks) {\n'
                  '  MRB_LAUNCH((splitk_reduce_kernel), blocks, 256, 0, nullptr, ws, splits, stride, M, N, bias, gelu, resid, ldr, out, out_dtype, ldc);\n'
                  '  return 0;\n}\n')
    so = str(tmp_path / "reduce.so")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-x", "c++", str(tu), "-o", so])
    lib = ctypes.CDLL(so)
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    g = torch.Generator().manual_seed(2)
    M, N, splits = 56, 264, 5
    ws = torch.randn(splits, M, N, generator=g)
    bias, resid = torch.randn(N, generator=g), torch.randn(M, N + 8, generator=g)
    total = ws[0].clone()
    for sp in range(1, splits):
        total += ws[sp]                                   # the kernel's summation order
    out = torch.full((M, N + 4), 7.0)
    assert lib.run(P(ws), splits, ctypes.c_longlong(M * N), M, N, P(bias), 1, P(resid), ctypes.c_longlong(N + 8), P(out), 2,
                   ctypes.c_longlong(N + 4), 3) == 0
    assert torch.allclose(out[:, :N], torch.nn.functional.gelu(total + bias) + resid[:, :N], rtol=1e-5, atol=1e-5)
    assert (out[:, N:] == 7.0).all()
    o16 = torch.zeros((M, N), dtype=torch.bfloat16)
    assert lib.run(P(ws), splits, ctypes.c_longlong(M * N), M, N, None, 0, None, ctypes.c_longlong(0), P(o16), 1, ctypes.c_longlong(N), 40) == 0
    assert torch.equal(o16, total.to(torch.bfloat16))


def test_no_task_prompt_variant_matches_oracle(tiny_sd):
    """task '..._no_task_prompt' (blip2_mr.py:651-654: the text prompt is the query alone): the product's row table against the
    oracle's prompt_concatenation, and the unsupported trainable-ViT / trainable-Q-Former settings refuse to run."""
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from oracle import blip2_mr as ob, synth
    m = BLIP2_MR(dims=TINY, state_dict=tiny_sd, task="qformer_freeze_lora_no_task_prompt")
    s = synth.make_samples(batch=2, frames=3, seed=6)
    n = TINY.num_query
    table, atts, _ = m.build_prompt_table(s["timestamps"], s["duration"], 2, 3, n, s["video_prompt_end"], s["query_prompt"], s["task_prompt"])
    frames = torch.randn(2, 3 * n, TINY.d_model)
    inputs, oatts = ob.prompt_concatenation(tiny_sd, TINY, m.t5_tokenizer, s["timestamps"], s["duration"], frames, s["video_prompt_end"],
                                            s["query_prompt"], s["task_prompt"], n, m.annoying_numbers_replacement_dict,
                                            task="qformer_freeze_lora_no_task_prompt")
    assert table.shape[1] == inputs.shape[1] and torch.equal(atts, oatts)
    emb = tiny_sd[T5_PREFIX + "shared.weight"]
    for b in range(2):
        for j in range(table.shape[1]):
            v = int(table[b, j])
            want = emb[v] if v >= 0 else (torch.zeros(TINY.d_model) if v == -2 ** 31 else frames.reshape(-1, TINY.d_model)[-v - 1])
            assert torch.equal(inputs[b, j], want), (b, j, v)
    full, _, _ = BLIP2_MR(dims=TINY, state_dict=tiny_sd).build_prompt_table(s["timestamps"], s["duration"], 2, 3, n, s["video_prompt_end"],
                                                                          s["query_prompt"], s["task_prompt"])
    assert full.shape[1] > table.shape[1]                    # the task prompt is gone
    with pytest.raises(NotImplementedError):
        BLIP2_MR(dims=TINY, state_dict=tiny_sd, freeze_vit=False)
    with pytest.raises(NotImplementedError):
        BLIP2_MR(dims=TINY, state_dict=tiny_sd, task="lora")
    assert BLIP2_MR(dims=TINY, state_dict=tiny_sd, use_grad_checkpoint=True).use_grad_checkpoint
