"""The oracle (oracle/*.py, fp32 CPU restatement) against the golden vectors produced by the
reference's own modules (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from mr_blip_b200.dims import TINY, T5_PREFIX
from mr_blip_b200.tokenizer import SyntheticT5Tokenizer
from mr_blip_b200 import mr_utils
from oracle import vit as ovit, qformer as oqf, t5 as ot5, blip2_mr as ob, synth
from oracle.beam_search import beam_search

TOL = dict(rtol=2e-4, atol=2e-4)   # fp32 vs fp32, different op order only


def _np(x):
    return x.detach().numpy()


def test_vision_stack_matches_reference(tiny_sd, golden_dir):
    gold = np.load(os.path.join(golden_dir, "vision_tiny.npz"))
    d = TINY
    frames = torch.randn(2, 3, d.img_size, d.img_size, generator=torch.Generator().manual_seed(7))
    assert abs(frames.double().sum().item() - float(gold["frames_checksum"])) < 1e-6, "input RNG drifted"
    with torch.no_grad():
        v = ovit.vit_forward(tiny_sd, d, frames)
        ie = ovit.ln_vision(tiny_sd, d, v)
        q = oqf.qformer_forward(tiny_sd, d, ie)
        p = torch.nn.functional.linear(q, tiny_sd["t5_proj.weight"], tiny_sd["t5_proj.bias"])
    tk = gold["vit_tokens"].tolist()
    np.testing.assert_allclose(_np(v[:, tk]), gold["vit_out"], **TOL)
    np.testing.assert_allclose(_np(ie[:, tk]), gold["image_embeds"], **TOL)
    np.testing.assert_allclose(_np(q), gold["qformer_out"], **TOL)
    np.testing.assert_allclose(_np(p[:, :, ::8]), gold["t5_proj_out"], **TOL)


def _t5_inputs(d):
    g = torch.Generator().manual_seed(11)
    emb = torch.randn(2, 72, d.d_model, generator=g) * 2.0
    mask = torch.ones(2, 72, dtype=torch.long)
    mask[1, 60:] = 0
    labels = torch.randint(2, 1000, (2, 9), generator=g)
    labels[:, -1] = 1
    labels[1, 6:] = -100
    labels[1, 5] = 1
    return emb, mask, labels


def test_t5_loss_logits_and_lora_grads_match_reference(tiny_sd, golden_dir):
    gold = np.load(os.path.join(golden_dir, "t5_tiny.npz"))
    d = TINY
    emb, mask, labels = _t5_inputs(d)
    assert abs(emb.double().sum().item() - float(gold["emb_checksum"])) < 1e-6
    assert (labels.numpy() == gold["labels"]).all()
    probes = [k[3:] for k in gold.files if k.startswith("gA.")]
    sd = dict(tiny_sd)
    leaves = {}
    for name in probes:
        for ab in ("lora_A", "lora_B"):
            k = f"{T5_PREFIX}{name}.{ab}.default.weight"
            leaves[k] = sd[k].clone().requires_grad_(True)
            sd[k] = leaves[k]
    emb.requires_grad_(True)
    out = ot5.t5_forward(sd, d, emb, mask, labels, (labels != -100).long())
    out["loss"].backward()
    assert abs(out["loss"].item() - float(gold["loss"])) < 1e-4
    np.testing.assert_allclose(_np(out["logits"][:, :, :256]), gold["logits_head"], **TOL)
    np.testing.assert_allclose(_np(torch.logsumexp(out["logits"], -1)), gold["logits_lse"], **TOL)
    np.testing.assert_allclose(_np(out["encoder_last_hidden_state"][:, ::8, ::4]), gold["enc_out"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(_np(emb.grad[:, ::4, ::4]), gold["d_emb"], rtol=1e-3, atol=1e-6)
    for name in probes:
        gA = leaves[f"{T5_PREFIX}{name}.lora_A.default.weight"].grad
        gB = leaves[f"{T5_PREFIX}{name}.lora_B.default.weight"].grad
        if name == "lm_head":
            gB = gB[::16]
        np.testing.assert_allclose(_np(gA), gold["gA." + name], rtol=2e-3, atol=2e-6, err_msg=name)
        np.testing.assert_allclose(_np(gB), gold["gB." + name], rtol=2e-3, atol=2e-6, err_msg=name)


def test_forward_mr_matches_reference_composition(tiny_sd, golden_dir):
    gold = np.load(os.path.join(golden_dir, "forward_mr_tiny.npz"))
    samples = synth.make_samples(batch=2, frames=3, seed=3)
    assert abs(samples["video"].double().sum().item() - float(gold["video_checksum"])) < 1e-6
    with torch.no_grad():
        out = ob.forward_mr(tiny_sd, TINY, SyntheticT5Tokenizer(), samples)
    assert out["inputs_embeds"].shape[1] == int(gold["L_enc"])
    assert (out["labels"].numpy() == gold["labels"]).all()
    assert abs(out["loss"].item() - float(gold["loss"])) < 2e-4
    np.testing.assert_allclose(_np(out["inputs_embeds"][:, ::16, ::8]), gold["inputs_embeds"], **TOL)
    np.testing.assert_allclose(_np(out["logits"][:, :, :256]), gold["logits_head"], rtol=1e-3, atol=1e-3)


def test_string_helpers_match_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "mr_utils_golden.json")))
    for s, want in gold["post_process"]:
        assert mr_utils.post_process(s) == want, s
    for s, want in gold["moment_str_to_list"]:
        assert mr_utils.moment_str_to_list(mr_utils.post_process(s)) == want, s
    si = gold["seconds_integers"]
    ts = [torch.tensor(t) for t in si["timestamps"]]
    table = {int(k): v for k, v in si["table"].items()}
    out = mr_utils.get_timestamps_as_seconds_integers(ts, torch.tensor(si["durations"]), table)
    assert [t.tolist() for t in out[0]] == si["out_ts"] and out[1] == si["out_dur"] and out[2] == si["out_prompt"]
    o_ts, o_d = ob.seconds_integers(ts, torch.tensor(si["durations"]), table)
    assert o_ts == si["out_ts"] and o_d == si["out_dur"]


def test_time_format_helpers_match_reference(golden_dir):
    """relative_integers / seconds_floats / relative_floats and convert_to_absolute_time against the outputs of the
    reference's own functions (tests/golden/make_golden_timefmt.py), incl. the strings the timestamps are tokenised from."""
    gold = json.load(open(os.path.join(golden_dir, "time_formats_golden.json")))
    ts = [torch.tensor(t) for t in gold["timestamps"]]
    du = torch.tensor(gold["durations"])
    for name in ("relative_integers", "seconds_floats", "relative_floats"):
        t, d, prompts = getattr(mr_utils, "get_timestamps_as_" + name)(ts, du, {})
        want = gold[name]
        assert prompts == want["prompt"], name
        assert [[str(v) for v in x.tolist()] for x in t] == want["ts_str"], name        # what _clean_ids tokenises
        assert [float(x) for x in d] == want["dur"]
    for fmt, (preds, want) in gold["absolute"].items():
        assert mr_utils.convert_to_absolute_time(preds, gold["durations"], fmt) == want, fmt


def test_moment_retrieval_metrics_match_reference(golden_dir):
    """R1@IoU / mIoU / mAP@IoU / invalid count against the reference's own evaluation code run on synthetic predictions
    (tests/golden/make_golden_mr_eval.py), plus the task-level report built from window strings."""
    from mr_blip_b200 import mr_eval
    gold = json.load(open(os.path.join(golden_dir, "mr_eval_golden.json")))
    sa = gold["single_ap"]
    assert np.allclose(mr_eval.average_precision(sa["gt"], sa["pred"]), sa["ap"], atol=1e-12)
    for case in gold["cases"]:
        got = mr_eval.moment_retrieval_metrics(case["records"])
        assert got["MR-mAP"] == case["MR-mAP"] and got["MR-R1"] == case["MR-R1"]
        assert abs(got["MR-R1-avg"] - case["MR-R1-avg"]) < 1e-9 and abs(got["MR-mIoU"] - case["MR-mIoU"]) < 1e-12
        assert got["MR-invalid_pred_num"] == case["MR-invalid_pred_num"]
    recs = gold["cases"][1]["records"]
    results = [{"qid": r["qid"], "prediction": str(r["pred_relevant_windows"]), "target": str(r["relevant_windows"])} for r in recs]
    rep = mr_eval.report_metrics(results)
    assert rep["total"] == len(recs) and rep["r1"] == gold["cases"][1]["MR-R1"] and rep["mAP"] == gold["cases"][1]["MR-mAP"]
    assert abs(rep["agg_metrics"] - gold["cases"][1]["MR-R1-avg"]) < 1e-9


def test_beam_search_degenerates_to_greedy_and_respects_eos():
    V = 12
    table = torch.full((V, V), -5.0)
    for i in range(V):
        table[i, (i + 3) % V] = 2.0       # deterministic chain 0->3->6->9->0...
    table[9, 1] = 4.0                     # 9 -> eos

    def step(ids):
        return table[ids[:, -1]]

    out = beam_search(step, batch=2, num_beams=1, max_new_tokens=10)
    assert out.tolist() == [[0, 3, 6, 9, 1]] * 2
    out5 = beam_search(step, batch=1, num_beams=5, max_new_tokens=10)
    assert out5[0, 0].item() == 0 and 1 in out5[0].tolist()
    capped = beam_search(lambda ids: table[ids[:, -1]].index_fill(1, torch.tensor([1]), -50.0), 1, 3, max_new_tokens=4)
    assert capped.shape[1] <= 5


def test_lr_schedules_match_reference(golden_dir):
    """mr_blip_b200.optim's schedulers against lr traces of the reference's own (lavis/common/optims.py), bit for bit
    (same float expressions): warm-up counted in global steps, cosine / step decay per epoch."""
    from mr_blip_b200 import optim
    from mr_blip_b200.registry import registry

    class Opt:
        def __init__(self):
            self.param_groups = [{"lr": None}, {"lr": None}]

    cases = json.load(open(os.path.join(golden_dir, "lr_sched_golden.json")))
    assert len(cases) >= 5
    for c in cases:
        opt = Opt()
        sched = registry.get_lr_scheduler_class(c["sched"])(optimizer=opt, **c["kwargs"])
        got = []
        for e in range(c["epochs"]):
            for i in range(c["iters"]):
                sched.step(cur_epoch=e, cur_step=i)
                assert opt.param_groups[0]["lr"] == opt.param_groups[1]["lr"]
                got.append(opt.param_groups[0]["lr"])
        assert got == c["lr"], c["sched"]
    assert optim.LinearWarmupCosineLRScheduler is registry.get_lr_scheduler_class("linear_warmup_cosine_lr")


def test_frame_sampling_and_sample_dict_match_reference(golden_dir, tmp_path):
    """mr_blip_b200.data against the reference's load_video index logic (all three sampling modes, clip proposals, short
    videos; same `random` seed -> same draws) and MomentRetrievalDataset.__getitem__ / collater output."""
    import random
    from mr_blip_b200 import data
    gold = json.load(open(os.path.join(golden_dir, "data_golden.json")))
    assert len(gold["indices"]) >= 10
    for c in gold["indices"]:
        random.seed(c["seed"])
        got = data.sample_frame_indices(c["vlen"], c["fps"], c["n_frms"], c["sampling"], c["clip"])
        assert [int(i) for i in got] == c["indices"], c
        rng = random.Random(c["seed"])                                       # an explicit generator draws the same stream
        assert [int(i) for i in data.sample_frame_indices(c["vlen"], c["fps"], c["n_frms"], c["sampling"], c["clip"], rng)] == c["indices"]

    class Reader:                                                            # the stub the golden generator decoded with
        def __init__(self, uri, height=-1, width=-1):
            self.h, self.w = height, width

        def __len__(self):
            return 4500

        def get_avg_fps(self):
            return 29.97

        def get_batch(self, idx):
            t = torch.tensor(idx, dtype=torch.float32).remainder(256).view(-1, 1, 1, 1)
            return t.expand(len(idx), self.h, self.w, 3).to(torch.uint8)

    path = tmp_path / "train.json"
    path.write_text(json.dumps(gold["annotations"]))

    def vis(video_path, clip_proposal=None):                                 # un-normalised float frames, as the generator's processor
        vr = Reader(video_path, 4, 4)
        idx = data.sample_frame_indices(len(vr), vr.get_avg_fps(), 6, "uniform", clip_proposal)
        return vr.get_batch(idx).permute(3, 0, 1, 2).float(), idx, vr.get_avg_fps()

    ds = data.MomentRetrievalDataset(vis, None, "/data/videos", [str(path)])
    assert len(ds) == len(gold["samples"])
    for i, want in enumerate(gold["samples"]):
        s = ds[i]
        assert set(s) == {"video", "duration", "query_id", "timestamps", "video_prompt_end", "query_prompt", "task_prompt", "relevant_windows"}
        for k in ("query_id", "video_prompt_end", "query_prompt", "task_prompt", "relevant_windows"):
            assert s[k] == want[k], k
        assert s["timestamps"].tolist() == want["timestamps"] and str(s["timestamps"].dtype) == want["timestamps_dtype"]
        assert float(s["duration"]) == want["duration"] and str(s["duration"].dtype) == want["duration_dtype"]
        assert list(s["video"].shape) == want["video_shape"] and str(s["video"].dtype) == want["video_dtype"]
        assert s["video"][:, 0, 0, 0].tolist() == want["video_frame_values"]
    b = ds.collater([ds[0], ds[1]])
    c = gold["collated"]
    assert list(b["video"].shape) == c["video_shape"] and list(b["timestamps"].shape) == c["timestamps_shape"]
    assert b["duration"].tolist() == c["duration"] and str(b["duration"].dtype) == c["duration_dtype"]
    assert b["query_id"].tolist() == c["query_id"] and list(b["relevant_windows"]) == c["relevant_windows"]
    assert ds.annotation[2]["instance_id"] == "2"


def test_beam_search_matches_transformers_generate():
    """The oracle's beam search (and through it the product's, tests/test_model_gpu.py) against GenerationMixin.generate of the
    INSTALLED transformers on random tiny T5s: beams 1-5, min_length, length penalties, padded encoder rows, with and without
    eos reached.  The reference pins transformers 4.46.1, which is not in this image; this pins the restatement to the
    library's published algorithm as shipped in the installed version (sequences compared up to the first eos -- the pad
    value after it differs between versions and is dropped by decoding anyway)."""
    tf = pytest.importorskip("transformers")
    if not hasattr(tf.T5ForConditionalGeneration, "generate"):
        pytest.skip("this transformers build has no T5 generate")

    def trim(row):
        out = []
        for t in row.tolist()[1:]:
            out.append(t)
            if t == 1:
                break
        return out

    n = with_eos = 0
    for seed in range(8):
        torch.manual_seed(seed)
        cfg = tf.T5Config(vocab_size=24, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_decoder_layers=2, num_heads=4,
                          feed_forward_proj="gated-gelu", tie_word_embeddings=False, pad_token_id=0, eos_token_id=1,
                          decoder_start_token_id=0, dropout_rate=0.0)
        m = tf.T5ForConditionalGeneration(cfg).eval()
        with torch.no_grad():
            for p in m.parameters():
                p.mul_(1.5 + 0.25 * (seed % 5))                                       # sharper / flatter next-token distributions
            if seed % 2:
                m.lm_head.weight[1] = 1.0 * m.lm_head.weight[3::4].sum(0)                  # make eos competitive
        B, L = 4, 6
        emb = torch.randn(B, L, 32)
        mask = torch.ones(B, L, dtype=torch.long)
        mask[2, 4:] = 0
        for nb, mnt, minl, lp in [(5, 12, 1, 1.0), (4, 8, 3, 1.0), (1, 10, 1, 1.0), (3, 15, 1, 2.0), (5, 6, 1, 0.5)]:
            with torch.no_grad():
                want = m.generate(inputs_embeds=emb, attention_mask=mask, num_beams=nb, max_new_tokens=mnt, min_length=minl,
                                  length_penalty=lp, do_sample=False, repetition_penalty=1.0, early_stopping=False)
                enc = m.encoder(inputs_embeds=emb, attention_mask=mask).last_hidden_state

            def step(ids):
                k = ids.shape[0] // B
                with torch.no_grad():
                    return m(encoder_outputs=(enc.repeat_interleave(k, 0),), attention_mask=mask.repeat_interleave(k, 0),
                             decoder_input_ids=ids).logits[:, -1]

            got = beam_search(step, batch=B, num_beams=nb, max_new_tokens=mnt, min_length=minl, length_penalty=lp)
            for b in range(B):
                a, c = trim(want[b]), trim(got[b])
                assert a == c, (seed, nb, mnt, minl, lp, b, a, c)
                n += 1
                with_eos += a[-1] == 1
    assert n == 160 and 12 <= with_eos <= 148, with_eos        # both finished and length-capped hypotheses were exercised


def test_train_crop_matches_reference_transform(golden_dir):
    """data.random_resized_crop against the reference's RandomResizedCropVideo + ToUint8 (bicubic, one window per clip) on the
    generator's synthetic clip under the same torch seed.  Same crop window required; pixel values may differ by one grey
    level where another CPU's interpolation rounds a .999 the other way."""
    import sys
    sys.path.insert(0, golden_dir)
    from make_golden_crop import synth_clip
    from mr_blip_b200 import data
    gold = np.load(os.path.join(golden_dir, "crop_golden.npz"))
    clip = synth_clip()
    for k in range(3):
        seed, size, s0, s1 = gold["meta%d" % k].tolist()
        torch.manual_seed(seed)
        got = data.random_resized_crop(clip, size, scale=(s0 / 1000.0, s1 / 1000.0))
        want = torch.from_numpy(gold["case%d" % k])
        assert got.dtype == torch.uint8 and got.shape == want.shape
        diff = (got.int() - want.int()).abs()
        assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 1e-3, (k, diff.max().item())
        torch.manual_seed(seed)
        again = data.random_resized_crop(clip.to(torch.uint8), size, scale=(s0 / 1000.0, s1 / 1000.0))     # uint8 frames in, as decoded
        assert torch.equal(again, got)


def test_train_mode_dropout_placement(tiny_sd, golden_dir):
    """The reference's Q-Former and T5 run in .train() after torch.manual_seed (tests/golden/make_golden_train_mode.py); the oracle
    with torch's own dropout at ITS sites consumes the same RNG stream, so it reproduces those outputs only if every mask sits
    where the reference's nn.Dropout calls sit (same tensors, same shapes, same order)."""
    import importlib.util
    from oracle.dropout import Dropper
    spec = importlib.util.spec_from_file_location("make_golden_train_mode", os.path.join(golden_dir, "make_golden_train_mode.py"))
    src = open(spec.origin).read()
    ns = {}
    exec(src[src.index("def inputs(d):"):src.index("def main():")], {"torch": torch}, ns)     # the seeded inputs, not the reference
    image_embeds, emb, mask, labels = ns["inputs"](TINY)
    gold = np.load(os.path.join(golden_dir, "train_mode_tiny.npz"))
    nthreads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        drop = Dropper(seed=0, lora=0.0, torch_rng=True)
        with torch.no_grad():
            torch.manual_seed(int(gold["rng_seed"]))
            q = oqf.qformer_forward(tiny_sd, TINY, image_embeds, drop=drop)
            torch.manual_seed(int(gold["rng_seed"]))
            out = ot5.t5_forward(tiny_sd, TINY, emb, mask, labels, (labels != -100).long(), drop=drop)
    finally:
        torch.set_num_threads(nthreads)
    np.testing.assert_allclose(_np(q), gold["qformer_out"], rtol=1e-3, atol=1e-4)
    assert abs(out["loss"].item() - float(gold["loss"])) < 1e-3
    np.testing.assert_allclose(_np(torch.logsumexp(out["logits"], -1)), gold["logits_lse"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(_np(out["logits"][:, :, :64]), gold["logits_head"], rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(_np(out["encoder_last_hidden_state"][:, ::4, ::8]), gold["enc_out"], rtol=1e-3, atol=1e-3)
    # and the eval-mode outputs differ from these by far more than the tolerance (the test can fail)
    with torch.no_grad():
        ev = ot5.t5_forward(tiny_sd, TINY, emb, mask, labels, (labels != -100).long())
    assert abs(ev["loss"].item() - float(gold["loss"])) > 1e-2


def test_counter_hash_masks_statistics_and_determinism():
    """oracle/dropout.py: keep rate = 1 - round(256 p) / 256, E[drop(x)] = x, independent sites / seeds / rows, p = 0 identity."""
    from oracle import dropout as od
    for p in (0.1, 0.05):
        m = od.keep_mask(7, od.site(od.ENC, 3, od.SELF_P), 2048, 2037, p)
        want = 1.0 - od.thr_of(p) / 256.0
        assert abs(m.mean() - want) < 5e-4, (p, m.mean(), want)
        assert abs(m.mean(axis=0) - want).max() < 0.05 and abs(m.mean(axis=1) - want).max() < 0.05
        assert abs(float(od.scale_of(p)) * want - 1.0) < 1e-6
    a = od.draws(7, od.site(od.ENC, 0, od.FF_RES), 64, 2048)
    assert (a == od.draws(7, od.site(od.ENC, 0, od.FF_RES), 64, 2048)).all()
    for other in (od.draws(8, od.site(od.ENC, 0, od.FF_RES), 64, 2048), od.draws(7, od.site(od.ENC, 1, od.FF_RES), 64, 2048),
                  od.draws(7, od.site(od.DEC, 0, od.FF_RES), 64, 2048)):
        c = np.corrcoef(a.ravel().astype(np.float64), other.ravel().astype(np.float64))[0, 1]
        assert abs(c) < 0.02, c
    c = np.corrcoef(a[:-1].ravel().astype(np.float64), a[1:].ravel().astype(np.float64))[0, 1]
    assert abs(c) < 0.02
    assert len(np.unique(a)) == 256
    x = torch.randn(5, 3, 44)
    d = od.Dropper(seed=3)
    assert d(x, 5, 0.0) is x
    y = d(x, 5, 0.1)
    keep = torch.from_numpy(od.keep_mask(3, 5, 15, 44, 0.1)).view(5, 3, 44)
    assert torch.equal(y != 0, keep & (x != 0)) and torch.allclose(y[keep], x[keep] * float(od.scale_of(0.1)))
    sites = {od.lora_site(f"t5_model.base_model.model.{s}.block.{i}.layer.{l}.{n}")
             for s in ("encoder", "decoder") for i in range(24)
             for l, n in ((0, "SelfAttention.q"), (0, "SelfAttention.k"), (0, "SelfAttention.v"), (0, "SelfAttention.o"),
                          (1, "EncDecAttention.q"), (1, "EncDecAttention.k"), (1, "EncDecAttention.v"), (1, "EncDecAttention.o"),
                          (2, "DenseReluDense.wi_0"), (2, "DenseReluDense.wi_1"), (2, "DenseReluDense.wo"))}
    sites.add(od.lora_site("t5_model.base_model.model.lm_head"))
    assert len(sites) == 2 * 24 * 11 + 1


def test_qa_frame_selection_matches_reference(golden_dir):
    """mr_blip_b200/qa.py against the reference's own get_relevant_frames / extract_frames (blip2_mr.py:1101-1165) on 96 cases:
    unparsable predictions, several windows, ends past the video, empty / inverted windows, padding and uniform thinning."""
    from mr_blip_b200 import qa
    both = json.load(open(os.path.join(golden_dir, "qa_frames_golden.json")))
    gold = both["selected"]
    assert len(gold) == 96 and len(both["resampled"]) == 33
    for c in both["resampled"]:                              # resample_frames=True: windows handed to the answerer's video processor
        calls = []

        def processor(path, clip_proposal=None):
            calls.append([path, [float(x) for x in clip_proposal]])
            return torch.full((3, 2, 1, 1), float(len(calls))), None, None
        samples = {"video": torch.zeros(1, 5, 3, 1, 1), "duration": torch.tensor([c["duration"]]), "video_path": [c["video_path"]]}
        m_in = c["moment_in"]
        moments, frames = qa.relevant_frames_resampled(samples, [m_in if isinstance(m_in, str) else list(m_in)], processor)
        assert [float(x) for x in moments[0]] == c["moment"] and calls == c["calls"] and list(frames.shape) == c["shape"], c
    for c in gold:
        T = c["T"]
        samples = {"video": torch.arange(T, dtype=torch.float32).view(1, T, 1, 1, 1), "timestamps": torch.tensor(c["timestamps"])[None],
                   "duration": torch.tensor([c["duration"]])}
        m = qa.relevant_moments_from_predictions([c["prediction"]], samples["duration"])
        assert [float(x) for x in m[0]] == c["moment"], c
        fr = qa.extract_frames(samples, m, c["n"])
        assert fr.shape == (1, c["n"], 1, 1, 1) and fr.view(-1).long().tolist() == c["frames"], c


def test_oracle_host_text_matches_reference(golden_dir):
    """oracle/host_text.py (the oracle's own restatement of the prompt strings, moment parsing and QA frame selection, so that
    no oracle comparison routes through mr_blip_b200/) against the outputs of the reference's own functions."""
    from oracle import host_text as ht
    si = json.load(open(os.path.join(golden_dir, "mr_utils_golden.json")))
    for s, want in si["moment_str_to_list"]:
        assert ht.parse_moments(mr_utils.post_process(s)) == want, s
    si = si["seconds_integers"]
    table = {int(k): v for k, v in si["table"].items()}
    got = ht.video_prompt("seconds_integers", [torch.tensor(t) for t in si["timestamps"]], torch.tensor(si["durations"]), table)
    assert got == si["out_prompt"]
    gold = json.load(open(os.path.join(golden_dir, "time_formats_golden.json")))
    ts, du = [torch.tensor(t) for t in gold["timestamps"]], torch.tensor(gold["durations"])
    for name in ("relative_integers", "seconds_floats", "relative_floats"):
        assert ht.video_prompt(name, ts, du, {}) == gold[name]["prompt"], name
    for c in json.load(open(os.path.join(golden_dir, "qa_frames_golden.json")))["selected"]:
        T = c["T"]
        samples = {"video": torch.arange(T, dtype=torch.float32).view(1, T, 1, 1, 1), "timestamps": torch.tensor(c["timestamps"])[None],
                   "duration": torch.tensor([c["duration"]])}
        win = ht.qa_window(c["prediction"], samples["duration"][0])
        assert [float(x) for x in win] == c["moment"], c
        assert ht.qa_frames(samples, [win], c["n"]).view(-1).long().tolist() == c["frames"], c


def test_dropout_mask_torch_matches_numpy():
    """The torch-integer evaluation of the counter-hash mask (used when the oracle runs on the GPU box) is the numpy one."""
    from oracle import dropout as od
    for rows, cols, p, site, seed in [(7, 13, 0.1, 5, 123), (33, 2048, 0.05, 0x2108, 0x9E3779B1), (5, 257, 0.1, 77, 1), (64, 2037, 0.1, 4097, 2 ** 32 - 1)]:
        assert (od.keep_mask(seed, site, rows, cols, p) == od.keep_mask_torch(seed, site, rows, cols, p, "cpu").numpy()).all()


def test_oracle_does_not_import_the_product():
    """oracle/ is the checker: none of its modules may import mr_blip_b200 (two helpers used to)."""
    import ast as _ast
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in glob.glob(os.path.join(root, "oracle", "*.py")):
        for node in _ast.walk(_ast.parse(open(f).read())):
            names = [a.name for a in node.names] if isinstance(node, _ast.Import) else \
                [node.module or ""] if isinstance(node, _ast.ImportFrom) else []
            assert not any(n.split(".")[0] == "mr_blip_b200" for n in names), f
