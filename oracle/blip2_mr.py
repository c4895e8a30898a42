"""fp32 restatement of BLIP2_MR.forward_mr / prompt_concatenation / generate
(lavis/models/blip2_mr_models/blip2_mr.py:433-570, 572-824, 826-988) composed around the pinned
sub-module oracles.  blip2_mr.py itself cannot be imported here (peft / tokenizer files absent), so
this file follows it line by line: interleave_data=True, task='qformer_freeze_lora' (every lavis/projects/mr_BLIP
yaml), input_time_format 'seconds_integers' (default) plus the other integer / float formats of utils.py:437-512.
Test infrastructure only (see oracle/__init__.py)."""
import contextlib

import torch

from . import vit as _vit, qformer as _qf, t5 as _t5, host_text as _ht
from .beam_search import beam_search


def _on(dev, enc):
    """Tokenizer output (CPU tensors) -> the device the weights live on (the oracle runs on the CPU in the tests and, for the
    full-depth parity test and the eager-GPU arm of bench.py, on the GPU box's device)."""
    return enc.input_ids.to(dev), enc.attention_mask.to(dev)


def seconds_integers(timestamps, durations, table):
    """utils.py:388-434."""
    ts, ds = [], []
    for t, dur in zip(timestamps, durations):
        ts.append([int(table.get(round(x.item()), round(x.item()))) for x in t])
        ds.append(table.get(round(dur.item()), round(dur.item())))
    return ts, ds


def time_values(fmt, timestamps, durations, table):
    """Values whose str() the reference tokenises per frame, and per clip for the duration (blip2_mr.py:600-630 dispatch to
    utils.py:388-529; the float formats go through a float32 tensor, so 149.6 reads back as 149.60000610351562)."""
    if fmt == "seconds_integers":
        return seconds_integers(timestamps, durations, table)
    ts, ds = [], []
    for t, dur in zip(timestamps, durations):
        du = dur.item()
        if fmt == "relative_integers":       # utils.py:437-461
            ts.append([int(round(x.item() / du, 2) * 100) for x in t])
        elif fmt == "seconds_floats":        # utils.py:464-484
            ts.append(torch.tensor([round(x.item(), 2) for x in t]).tolist())
        elif fmt == "relative_floats":       # utils.py:487-512 (the extra duration element is never indexed)
            ts.append(torch.tensor([round(x.item() / du, 2) for x in t] + [round(du)]).tolist())
        else:
            raise ValueError(fmt)
        ds.append(torch.as_tensor(durations)[len(ds)].item())      # durations pass through unchanged
    return ts, ds


def _amp(dev, dtype, on):
    """The reference's GPU regime (amp=True): torch.amp.autocast of `dtype`; off (fp32, the parity oracle) otherwise."""
    return torch.autocast(torch.device(dev).type, dtype=dtype) if on else contextlib.nullcontext()


def frame_tokens(sd, d, video, frame_token_aggregation=None, drop=None, amp=False):
    """blip2_mr.py:443-510: ViT -> ln_vision -> Q-Former -> t5_proj (-> mean) -> [b, t*n, c].
    amp=True: the reference's GPU regime -- ViT + ln_vision under fp16 autocast (blip2_mr.py:446, fp16 ViT weights:
    eva_vit.py:397-412), Q-Former and t5_proj under the training loop's fp16 autocast (moment_retrieval.py:217, `amp: True` in every
    mr_BLIP recipe)."""
    b, t = video.shape[:2]
    image = video.reshape(-1, *video.shape[2:])
    with _amp(image.device, torch.float16, amp):
        image_embeds = _vit.ln_vision(sd, d, _vit.vit_forward(sd, d, image))
        q = _qf.qformer_forward(sd, d, image_embeds, drop=drop)
        f = torch.nn.functional.linear(q, sd["t5_proj.weight"], sd["t5_proj.bias"])
        if frame_token_aggregation:
            f = f.mean(dim=1, keepdim=True)
    return f.reshape(b, -1, f.shape[-1]), {"image_embeds": image_embeds, "qformer": q}


def clean_timestamp_ids(tok, values):
    """get_clean_timestamp_tokens_and_embs, blip2_mr.py:1561-1608: tokenize str(v) without specials,
    drop a leading id 3."""
    ids = tok([str(v) for v in values], add_special_tokens=False)["input_ids"]
    return [i[1:] if i[0] == 3 else i for i in ids]


def prompt_concatenation(sd, d, tok, timestamps, durations, frames_for_t5, video_prompt_end, query_prompt,
                         task_prompt, n_per_frame, table=None, max_txt_len=200, prefix=_t5.PREFIX,
                         input_time_format="seconds_integers", interleave_data=True, task="qformer_freeze_lora"):
    """blip2_mr.py:572-824, interleave branch:
       [f_0 (n) | ts_0 | f_1 | ts_1 | ... | '>' | duration] (left-padded) ++ video_prompt_end ++ query+task;
    interleave_data=False (:784-822): video_prompt (the timestamps as text) ++ all frame tokens ++ video_prompt_end ++ query+task."""
    emb = sd[prefix + "shared.weight"]
    dev = emb.device
    timestamps, durations = torch.as_tensor(timestamps).cpu(), torch.as_tensor(durations).cpu()
    if "no_task_prompt" in task:                              # blip2_mr.py:651-654: the query alone
        task_prompt = [""] * len(query_prompt)
    if not interleave_data:
        video_prompt = _ht.video_prompt(input_time_format, timestamps, durations, table or {})
        kw = dict(padding="longest", truncation=True, max_length=max_txt_len, return_tensors="pt")
        vp_ids, vp_m = _on(dev, tok(video_prompt, add_special_tokens=False, **kw))
        end_ids, end_m = _on(dev, tok(video_prompt_end, add_special_tokens=False, **kw))
        text_ids, text_m = _on(dev, tok([q + t for q, t in zip(query_prompt, task_prompt)], **kw))
        inputs = torch.cat([emb[vp_ids], frames_for_t5, emb[end_ids], emb[text_ids]], dim=1)
        atts = torch.cat([vp_m, torch.ones(frames_for_t5.shape[:2], dtype=torch.long, device=dev), end_m, text_m], dim=1)
        return inputs, atts
    ts, ds = time_values(input_time_format, timestamps, durations, table or {})
    end_ids, end_m = _on(dev, tok(video_prompt_end, padding="longest", add_special_tokens=False, truncation=True,
                                  max_length=max_txt_len, return_tensors="pt"))
    text_ids, text_m = _on(dev, tok([q + t for q, t in zip(query_prompt, task_prompt)], padding="longest", truncation=True,
                                    max_length=max_txt_len, return_tensors="pt"))
    sep = tok.convert_tokens_to_ids(">")
    B, TN, C = frames_for_t5.shape
    T = TN // n_per_frame
    rows = []
    for j in range(B):
        ts_ids = clean_timestamp_ids(tok, ts[j])
        dur_ids = clean_timestamp_ids(tok, [ds[j]])[0]
        parts = []
        for i in range(T):
            parts.append(frames_for_t5[j, i * n_per_frame:(i + 1) * n_per_frame])
            parts.append(emb[torch.tensor(ts_ids[i], device=dev)])
        parts.append(emb[torch.tensor([sep], device=dev)])
        parts.append(emb[torch.tensor(dur_ids, device=dev)])
        rows.append(torch.cat(parts))
    L = max(len(r) for r in rows)
    # reference pads with pad_token_id * ones (= zeros) on the LEFT and still marks them attended (:744-779)
    rows = [torch.cat([torch.zeros(L - len(r), C, device=dev, dtype=r.dtype), r]) if len(r) < L else r for r in rows]
    inter = torch.stack(rows)
    inputs = torch.cat([inter, emb[end_ids], emb[text_ids]], dim=1)
    atts = torch.cat([torch.ones(B, L, dtype=torch.long, device=dev), end_m, text_m], dim=1)
    return inputs, atts


def forward_mr(sd, d, tok, samples, frame_token_aggregation=None, table=None, max_txt_len=200,
               input_time_format="seconds_integers", drop=None, interleave_data=True, amp=False, task="qformer_freeze_lora"):
    """blip2_mr.py:433-570 -> dict(loss, logits, inputs_embeds, attention_mask, labels, ...).  drop: None = eval mode, a
    Dropper (oracle/dropout.py) = the train-mode dropout of the Q-Former, T5 and LoRA inputs (the ViT stays in eval mode:
    blip2_mr.py:136-137).  amp=True: the reference's autocast regime on a GPU (fp16 frame encoder, bf16 T5: blip2_mr.py:446,512)
    -- the eager-PyTorch-on-B200 bar of BASELINE.md section 4; amp=False is the fp32 parity oracle."""
    f, aux = frame_tokens(sd, d, samples["video"], frame_token_aggregation, drop=drop, amp=amp)
    n = 1 if frame_token_aggregation else d.num_query
    with _amp(f.device, torch.bfloat16, amp):
        inputs, atts = prompt_concatenation(sd, d, tok, samples["timestamps"], samples["duration"], f,
                                            samples["video_prompt_end"], samples["query_prompt"],
                                            samples["task_prompt"], n, table, max_txt_len, input_time_format=input_time_format,
                                            interleave_data=interleave_data, task=task)
        ans_ids, ans_m = _on(inputs.device, tok(samples["relevant_windows"], padding="longest", truncation=True,
                                                max_length=max_txt_len, return_tensors="pt"))
        labels = ans_ids.masked_fill(ans_ids == tok.pad_token_id, -100)
        out = _t5.t5_forward(sd, d, inputs, atts, labels, ans_m, drop=drop)
    out.update(inputs_embeds=inputs, attention_mask=atts, labels=labels, frames_for_t5=f, **aux)
    return out


@torch.no_grad()
def generate(sd, d, tok, samples, post_process, num_beams=5, max_length=50, min_length=1, length_penalty=1.0,
             frame_token_aggregation=None, table=None):
    """blip2_mr.py:826-946 (no-cache decoder re-run each step, as the reference effectively does)."""
    f, _ = frame_tokens(sd, d, samples["video"], frame_token_aggregation)
    n = 1 if frame_token_aggregation else d.num_query
    inputs, atts = prompt_concatenation(sd, d, tok, samples["timestamps"], samples["duration"], f,
                                        samples["video_prompt_end"], samples["query_prompt"],
                                        samples["task_prompt"], n, table)
    enc = _t5.t5_encoder(sd, d, inputs, atts)
    B = inputs.shape[0]
    enc_b = enc.repeat_interleave(num_beams, dim=0)
    atts_b = atts.repeat_interleave(num_beams, dim=0)

    def step(ids):
        dec = _t5.t5_decoder(sd, d, ids, enc_b, atts_b)
        return _t5.t5_logits(sd, d, dec[:, -1])

    seqs = beam_search(step, B, num_beams, max_length, min_length, length_penalty,
                       eos_id=tok.eos_token_id, pad_id=tok.pad_token_id)
    raw = tok.batch_decode(seqs, skip_special_tokens=True)
    dur = samples["duration"]
    return {"prediction": [post_process(p) for p in raw], "raw_prediction": raw, "sequences": seqs,
            "answer": samples["relevant_windows"], "qid": samples["query_id"],
            "duration": dur.tolist() if torch.is_tensor(dur) else dur}


# ---------------------------------------------------------------------------------------------- two-stage video QA
ANSWERER_PREFIX = "answerer_model.base_model.model."
ANSWER_IDS = [71, 272, 205, 309, 262]          # A B C D E (blip2_mr.py:1297)


def _qa_relevant_frames(sd, d, tok, samples, use_localizer, n_frames, post_process, frame_token_aggregation):
    """blip2_mr.py:328-362 / 1011-1049: window from the localizer's prediction or the whole video, then extract_frames
    (oracle/host_text.py: the oracle's own restatement of get_relevant_frames / extract_frames, pinned by qa_frames_golden.json)."""
    if use_localizer:
        pred = generate(sd, d, tok, samples, post_process, frame_token_aggregation=frame_token_aggregation)["prediction"]
        moments = [_ht.qa_window(p, samples["duration"][i]) for i, p in enumerate(pred)]
    else:
        moments = [[0, x.item()] for x in samples["duration"]]
    return moments, _ht.qa_frames(samples, moments, n_frames)


def _qa_inputs(sd, tok, f, texts, max_txt_len):
    ids, m = _on(f.device, tok(texts, padding="longest", truncation=True, max_length=max_txt_len, return_tensors="pt"))
    emb = sd[ANSWERER_PREFIX + "shared.weight"][ids]
    return torch.cat([f, emb], dim=1), torch.cat([torch.ones(f.shape[:2], dtype=torch.long, device=f.device), m], dim=1)


def forward_qa(sd, d, tok, samples, use_localizer=False, n_frames=4, post_process=None, frame_token_aggregation=None,
               max_txt_len=200, drop=None):
    """forward_QA, blip2_mr.py:309-431 -> dict(loss, logits, relevant_moments).  Stage 1 without gradients; the localizer's
    generate runs the eval-mode arithmetic (the product's choice), `drop` applies to the Q-Former and the answerer."""
    samples = dict(samples)
    samples["relevant_windows"], samples["query_id"] = [[0, 0]], samples["question_id"]
    with torch.no_grad():
        moments, rel = _qa_relevant_frames(sd, d, tok, samples, use_localizer, n_frames, post_process, frame_token_aggregation)
        f, _ = frame_tokens(sd, d, rel, frame_token_aggregation, drop=drop)
    inputs, atts = _qa_inputs(sd, tok, f, samples["qa_input"], max_txt_len)
    a_ids, a_m = _on(inputs.device, tok(samples["qa_output"], padding="longest", truncation=True, max_length=max_txt_len,
                                        return_tensors="pt"))
    labels = a_ids.masked_fill(a_ids == tok.pad_token_id, -100)
    out = _t5.t5_forward(sd, d, inputs, atts, labels, a_m, prefix=ANSWERER_PREFIX, drop=drop)
    out.update(relevant_moments=moments, labels=labels)
    return out


def videoqa_answer(sd, d, tok, samples, rel, frame_token_aggregation=None, max_txt_len=200, min_length=8):
    """videoQA_answer, blip2_mr.py:1233-1314: greedy answerer, arg-max over the letter ids of the second position's scores."""
    with torch.no_grad():
        f, _ = frame_tokens(sd, d, rel, frame_token_aggregation)
        inputs, atts = _qa_inputs(sd, tok, f, samples["qa_input"], max_txt_len)
        enc = _t5.t5_encoder(sd, d, inputs, atts, prefix=ANSWERER_PREFIX)
        ids = torch.full((inputs.shape[0], 1), tok.pad_token_id, dtype=torch.long, device=inputs.device)
        for _ in range(2):
            dec = _t5.t5_decoder(sd, d, ids, enc, atts, prefix=ANSWERER_PREFIX)
            logits = _t5.t5_logits(sd, d, dec[:, -1], prefix=ANSWERER_PREFIX)
            if min_length > 1:
                logits[:, tok.eos_token_id] = float("-inf")
            ids = torch.cat([ids, logits.argmax(-1, keepdim=True)], dim=1)
    return logits[:, ANSWER_IDS].argmax(-1).tolist(), logits[:, ANSWER_IDS]
