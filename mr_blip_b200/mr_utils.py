"""Host-side string/timestamp helpers of the moment-retrieval path.

Behavioural restatements of lavis/models/blip2_mr_models/utils.py: post_process (:18-83),
moment_str_to_list (:300-341), get_timestamps_as_seconds_integers (:388-434) and the sibling
input_time_format variants (:437-529).  Checked against the reference's functions on the string
corpus in tests/golden/mr_utils_golden.json.
"""
import ast
import re

import torch

_NESTED = re.compile(r"\[\[.*\]\]")
ERROR_WINDOW = "[[-1, -1]]"


def post_process(pred: str) -> str:
    """Normalise a decoded moment string into "[[s, e], [s, e]]" (reference utils.py:18-83)."""
    pred = pred.split("</s>")[0]
    if not _NESTED.match(pred):
        return ERROR_WINDOW
    body = pred[1:-1]
    fixed = []
    for w in re.split(r"\s+(?=\[)", body):
        w = re.sub(r",+$", "", w)                 # trailing commas
        w = re.sub(r"(\d) (\d)", r"\1, \2", w)    # missing comma between two numbers
        w = re.sub(r",+", ",", w)                 # doubled commas
        nums = re.findall(r"\d+", w)
        if len(nums) == 2 and int(nums[0]) > int(nums[1]):
            w = "[" + nums[1] + ", " + nums[0] + "]"
        fixed.append(w)
    return "[" + ", ".join(fixed) + "]"


def moment_str_to_list(m: str):
    """"[[0, 1], [4, 7]]" -> [[0, 1], [4, 7]]; anything malformed -> [[-1, -1]] (utils.py:300-341)."""
    if m == ERROR_WINDOW or not _NESTED.match(m):
        return [[-1, -1]]
    try:
        val = ast.literal_eval(m)
    except Exception:
        return [[-1, -1]]
    if not isinstance(val, list):
        return [[-1, -1]]
    for i in range(len(val)):
        if len(val[i]) != 2:
            val[i] = [-1, -1]
    return val


def _replace(n, table):
    return table[n] if (table and n in table) else n


def get_timestamps_as_seconds_integers(timestamps, durations, annoying_numbers_replacement_dict=None):
    """Round each frame timestamp / clip duration to whole seconds, substituting integers the
    tokenizer splits in two (utils.py:388-434).  -> (list[LongTensor[T]], list[int], list[str])."""
    table = annoying_numbers_replacement_dict or {}
    new_ts, new_durs, prompts = [], [], []
    for t, d in zip(timestamps, durations):
        secs = [int(_replace(round(float(x)), table)) for x in t]
        dur = _replace(round(float(d)), table)
        prompts.append(">" + ">".join(str(s) for s in secs) + ">" + str(dur))
        new_ts.append(torch.tensor(secs))
        new_durs.append(dur)
    return new_ts, new_durs, prompts


def get_timestamps_as_relative_integers(timestamps, durations, annoying_numbers_replacement_dict=None):
    """Timestamps as integer percent of the duration; durations pass through (utils.py:437-461)."""
    new_ts, prompts = [], []
    for t, d in zip(timestamps, durations):
        dur = float(d)
        rel = [int(round(float(x) / dur, 2) * 100) for x in t]
        prompts.append(">".join(str(s) for s in rel) + ">" + str(round(dur)))
        new_ts.append(torch.tensor(rel))
    return new_ts, durations, prompts


def get_timestamps_as_framenumbers(timestamps, durations, annoying_numbers_replacement_dict=None):
    """Frame indices 0..T-1 instead of seconds (intent of utils.py:512-529, whose string concat of
    `d.item()` raises TypeError as shipped)."""
    new_ts, prompts = [], []
    for t, d in zip(timestamps, durations):
        prompts.append(">".join(str(i) for i in range(len(t))) + ">" + str(float(d)))
        new_ts.append(torch.arange(len(t)))
    return new_ts, durations, prompts


def get_timestamps_as_seconds_floats(timestamps, durations, annoying_numbers_replacement_dict=None):
    """Timestamps in seconds rounded to 2 decimals (utils.py:464-484).  The float32 tensor the reference builds is what gets
    tokenised later (str(t.item()), blip2_mr.py:1576-1578), e.g. 149.6 -> "149.60000610351562"."""
    new_ts, prompts = [], []
    for t, d in zip(timestamps, durations):
        vals = [round(float(x), 2) for x in t]
        prompts.append(">".join(str(v) for v in vals) + ">" + str(round(float(d))))
        new_ts.append(torch.tensor(vals))
    return new_ts, durations, prompts


def get_timestamps_as_relative_floats(timestamps, durations, annoying_numbers_replacement_dict=None):
    """Timestamps as fractions of the duration, 2 decimals (utils.py:487-512): the prompt string drops the last frame and
    the tensor carries the rounded duration as an extra element, exactly as the reference does."""
    new_ts, prompts = [], []
    for t, d in zip(timestamps, durations):
        dur = float(d)
        rel = [round(float(x) / dur, 2) for x in t]
        prompts.append(">".join(str(v) for v in rel[:-1]) + ">" + str(round(dur)))
        new_ts.append(torch.tensor(rel + [round(dur)]))
    return new_ts, durations, prompts


def convert_to_absolute_time(prediction, duration, input_time_format="relative_integers"):
    """Relative windows (percent or fraction of the duration) back to seconds, 2 decimals (utils.py:242-297)."""
    assert input_time_format in ("relative_integers", "relative_floats"), "This function is only used for relative timestamps"
    div = 100.0 if input_time_format == "relative_integers" else 1.0
    out = []
    for pred, dur in zip(prediction, duration):
        dur = float(dur)
        out.append(str([[round((float(s_) / div) * dur, 2), round((float(e_) / div) * dur, 2)] if s_ != -1 and e_ != -1
                        else [-1, -1] for s_, e_ in moment_str_to_list(pred)]))
    return out


def find_annoying_numbers(tokenizer, range_end=200):
    """Integers the tokenizer encodes as more than one token (blip2_mr.py:1497-1534)."""
    annoying, annoying_space = [], []
    for i in range(range_end):
        ids = tokenizer(str(i), add_special_tokens=False)["input_ids"]
        if ids and isinstance(ids[0], list):
            ids = ids[0]
        if len(ids) > 1:
            (annoying_space if ids[0] == 3 else annoying).append(i)
    return annoying, annoying_space


def find_annoying_numbers_replacement_dict(annoying):
    """Nearest integer that is a single token, searching +j before -j (blip2_mr.py:1536-1559)."""
    s, table = set(annoying), {}
    for i in annoying:
        for j in range(100):
            if (i + j) not in s:
                table[i] = i + j
                break
            if (i - j) not in s:
                table[i] = i - j
                break
    return table
