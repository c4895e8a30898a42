"""Golden vectors for the QA branch's frame selection: the reference's own get_relevant_frames / extract_frames
(lavis/models/blip2_mr_models/blip2_mr.py:1101-1165) -- the two methods are cut out of the class source (the module itself cannot
be imported here: peft, omegaconf, ... are absent) and executed unmodified with the reference's moment_str_to_list.

Run here:   python tests/golden/make_golden_qa.py      -> tests/golden/qa_frames_golden.json
"""
import ast
import json
import os
import sys
import textwrap

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

SRC = "/root/reference/lavis/models/blip2_mr_models/blip2_mr.py"


def reference_methods():
    src = open(SRC).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "BLIP2_MR")
    ns = {"torch": torch, "moment_str_to_list": ref_shim.load_reference_utils().moment_str_to_list}
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("get_relevant_frames", "extract_frames", "get_relevant_frames_resampled"):
            exec(textwrap.dedent(ast.get_source_segment(src, fn)), ns)
    return ns["get_relevant_frames"], ns["extract_frames"], ns["get_relevant_frames_resampled"]


def cases():
    g = torch.Generator().manual_seed(4)
    out = []
    for T, dur in ((20, 31.4), (60, 150.0), (8, 12.0), (5, 3.2)):
        ts = torch.linspace(0.5 * dur / T, dur - 0.5 * dur / T, T)
        ts = (ts * 100).round() / 100
        out.append((T, dur, ts))
    return out


def main():
    get_relevant_frames, extract_frames, get_resampled = reference_methods()

    class Self:
        pass
    self = Self()
    self.extract_frames = lambda s, m, n: extract_frames(self, s, m, n)
    preds = ["[[3, 9]]", "[[10, 40], [60, 70]]", "garbage", "[[5, 400]]", "[[7, 7]]", "[[9, 2]]", "[[0, 1]]", "[[2.5, 2.6]]"]
    gold = []
    for T, dur, ts in cases():
        for n in (1, 4, 7):
            for pred in preds:
                video = torch.arange(T, dtype=torch.float32).view(1, T, 1, 1, 1)      # frame i carries the value i
                samples = {"video": video, "timestamps": ts[None], "duration": torch.tensor([dur])}
                moments, frames = get_relevant_frames(self, samples, [pred], n)
                gold.append({"T": T, "duration": dur, "timestamps": ts.tolist(), "n": n, "prediction": pred,
                             "moment": [float(x) for x in moments[0]], "frames": frames.view(-1).long().tolist()})
    # resample_frames=True: which windows the answerer's video processor is asked to decode (a recording stub stands for it)
    calls = []

    def processor(path, clip_proposal=None):
        calls.append([path, [float(x) for x in clip_proposal]])
        return torch.full((3, 2, 1, 1), float(len(calls))), None, None
    self.video_processor_answerer_eval = processor
    resampled = []
    for dur in (31.4, 150.0, 3.2):
        for m in preds + [[4, 9], [9, 4], [2, 400]]:
            del calls[:]
            samples = {"video": torch.zeros(1, 5, 3, 1, 1), "duration": torch.tensor([dur]), "video_path": ["v%d.mp4" % len(resampled)]}
            moments, frames = get_resampled(self, samples, [m], 2)
            resampled.append({"duration": dur, "moment_in": m, "video_path": samples["video_path"][0],
                              "moment": [float(x) for x in moments[0]], "calls": [list(c) for c in calls], "shape": list(frames.shape)})
    with open(os.path.join(HERE, "qa_frames_golden.json"), "w") as f:
        json.dump({"selected": gold, "resampled": resampled}, f)
    print(len(gold), "+", len(resampled), "cases")


if __name__ == "__main__":
    main()
