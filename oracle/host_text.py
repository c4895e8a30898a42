"""The oracle's OWN restatement of the host-side string / index helpers the path needs outside the interleave branch, so that
no oracle comparison routes through product code (mr_blip_b200/ is never imported from oracle/).  Test infrastructure only.

  * video_prompt():  the "timestamps as text" string of the non-interleaved prompt design
                     (lavis/models/blip2_mr_models/utils.py:388-529, third return value of get_timestamps_as_*)
  * parse_moments(): utils.py:300-341 moment_str_to_list
  * qa_window() / qa_frame_index() / qa_frames(): blip2_mr.py:1101-1165 get_relevant_frames / extract_frames

Pinned against the reference's own functions through tests/golden/time_formats_golden.json, mr_utils_golden.json and
qa_frames_golden.json (tests/test_oracle_golden.py::test_oracle_host_text_*).
"""
import ast
import re

import torch


def _num(x):
    return x.item() if torch.is_tensor(x) else x


def video_prompt(fmt, timestamps, durations, table=None):
    """One prompt string per clip: the frame times joined by '>' and closed by the clip duration."""
    table = table or {}
    out = []
    for clip_times, clip_dur in zip(timestamps, durations):
        dur = _num(clip_dur)
        times = [_num(t) for t in clip_times]
        if fmt == "seconds_integers":                       # utils.py:388-434 (leading '>' only in this format)
            words = []
            for t in times:
                r = round(t)
                words.append(str(table[r]) if r in table else str(int(r)))
            rd = round(dur)
            tail = table[rd] if rd in table else rd
            out.append(">" + ">".join(words) + ">" + str(tail))
        elif fmt == "relative_integers":                    # utils.py:437-461
            out.append(">".join(str(int(round(t / dur, 2) * 100)) for t in times) + ">" + str(round(dur)))
        elif fmt == "seconds_floats":                       # utils.py:464-484
            out.append(">".join(str(round(t, 2)) for t in times) + ">" + str(round(dur)))
        elif fmt == "relative_floats":                      # utils.py:487-512 (the last frame is left out of the string)
            out.append(">".join(str(round(t / dur, 2)) for t in times[:-1]) + ">" + str(round(dur)))
        elif fmt == "framenumbers":                         # utils.py:515-529 (intent; as shipped the str + float concat raises)
            out.append(">".join(str(i) for i in range(len(times))) + ">" + str(float(dur)))
        else:
            raise ValueError(fmt)
    return out


_BAD = [[-1, -1]]


def parse_moments(text):
    """'[[0, 1], [4, 7]]' -> [[0, 1], [4, 7]]; whatever does not parse as a list of pairs -> [[-1, -1]] (utils.py:300-341)."""
    if text == "[[-1, -1]]" or re.match(r"\[\[.*\]\]", text) is None:
        return [[-1, -1]]
    try:
        val = ast.literal_eval(text)
    except Exception:
        return [[-1, -1]]
    if not isinstance(val, list):
        return [[-1, -1]]
    return [w if len(w) == 2 else [-1, -1] for w in val]


def qa_window(prediction, duration):
    """get_relevant_frames, blip2_mr.py:1104-1116: first predicted window, whole clip when unparsable, end clipped."""
    dur = _num(duration)
    wins = parse_moments(prediction)
    win = [0, dur] if wins == _BAD else list(wins[0])
    if win[1] > dur:
        win[1] = round(dur)
    return win


def qa_frame_index(timestamps, duration, start, end, n):
    """extract_frames, blip2_mr.py:1128-1159, as positions into one clip's sampled frames."""
    if start >= end:
        end = _num(duration)
    ts = torch.as_tensor(timestamps).float().cpu()
    first = int(torch.argmin((ts - start).abs()))
    last = int(torch.argmin((ts - end).abs()))
    pos = list(range(first, last + 1))
    assert pos, "No frames found for the relevant moment."
    if len(pos) < n:
        pos = pos + [pos[-1]] * (n - len(pos))
    elif len(pos) > n:
        pos = [pos[i] for i in torch.linspace(0, len(pos) - 1, n).long().tolist()]
    return pos


def qa_frames(samples, windows, n):
    video = samples["video"]
    picks = [video[i][torch.tensor(qa_frame_index(samples["timestamps"][i], samples["duration"][i], s, e, n), device=video.device)]
             for i, (s, e) in enumerate(windows)]
    return torch.stack(picks)
