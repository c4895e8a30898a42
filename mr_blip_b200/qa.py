"""Host-side pieces of the two-stage video-QA branch (lavis/models/blip2_mr_models/blip2_mr.py:309-431 forward_QA, :990-1099
videoQA_generate, :1233-1314 videoQA_answer): turning the localizer's moment strings into one [start, end] window per clip and
picking `num_frames_for_answer` of the already-sampled frames inside it.  Pure index arithmetic on the sample dict -- the frames
stay where they are (host or device); the kernels see them through BLIP2_MR.get_frame_embeddings_and_attentions.
Pinned against the reference's own methods in tests/golden/qa_frames_golden.json (make_golden_qa.py).
"""
import torch

from . import mr_utils

# token ids of the answer letters A B C D E in the FlanT5 vocabulary (blip2_mr.py:1297)
ANSWER_IDS = [71, 272, 205, 309, 262]


def _item(x):
    return x.item() if torch.is_tensor(x) else x


def relevant_moments_from_predictions(predictions, durations):
    """get_relevant_frames, blip2_mr.py:1101-1118: the first predicted window of each clip; an unparsable prediction
    ([[-1, -1]]) means the whole video; an end past the video is clipped to round(duration)."""
    out = []
    for i, text in enumerate(predictions):
        m = mr_utils.moment_str_to_list(text)
        dur = _item(durations[i])
        m = [0, dur] if m == [[-1, -1]] else m[0]
        if m[1] > dur:
            m[1] = round(dur)
        out.append(m)
    return out


def frame_indices(timestamps, duration, start, end, n):
    """extract_frames, blip2_mr.py:1128-1159, for one clip: indices into the clip's sampled frames -- the frames whose timestamps
    are closest to start / end and everything between, padded with the last one or thinned uniformly to n."""
    if start >= end:
        end = _item(duration)
    ts = torch.as_tensor(timestamps).detach().float().cpu()
    s = torch.argmin(torch.abs(ts - start)).item()
    e = torch.argmin(torch.abs(ts - end)).item()
    idx = torch.arange(s, e + 1)
    assert idx.numel() > 0, "No frames found for the relevant moment."
    if idx.numel() < n:
        idx = torch.cat([idx, idx[-1:].expand(n - idx.numel())])
    elif idx.numel() > n:
        idx = idx[torch.linspace(0, idx.numel() - 1, n).long()]
    return idx


def extract_frames(samples, relevant_moments, n):
    """-> [b, n, c, h, w] (same device / dtype as samples["video"])."""
    video = samples["video"]
    out = []
    for i, (start, end) in enumerate(relevant_moments):
        idx = frame_indices(samples["timestamps"][i], samples["duration"][i], start, end, n)
        out.append(video[i][idx.to(video.device)])
    return torch.stack(out)


def relevant_frames_resampled(samples, relevant_moments, processor):
    """get_relevant_frames_resampled, blip2_mr.py:1167-1231: the answerer's frames are RE-DECODED from samples["video_path"]
    inside the proposed window by the answerer's eval video processor (already set to num_frames_for_answer frames) instead of
    being picked among the clip's sampled frames.  relevant_moments: the localizer's strings, or [start, end] pairs.
    -> (moments, frames [b, n, 3, H, W] on the device of samples["video"])."""
    if isinstance(relevant_moments[0], str):
        moments = []
        for i, text in enumerate(relevant_moments):
            m = mr_utils.moment_str_to_list(text)
            dur = _item(samples["duration"][i])
            m = [0, round(dur)] if m == [[-1, -1]] else m[0]
            if m[1] > dur:
                m[1] = round(dur)
            moments.append(m)
    else:
        moments = relevant_moments
    assert len(moments) == samples["video"].shape[0]
    out = []
    for i, (start, end) in enumerate(moments):
        if start >= end:
            end = _item(samples["duration"][i])
        frames, _, _ = processor(samples["video_path"][i], clip_proposal=[start, end])      # c, n, h, w
        out.append(frames.permute(1, 0, 2, 3))
    return moments, torch.stack(out).to(samples["video"].device)
