"""Golden vectors for the input side of the path: frame-index sampling of the REFERENCE's load_video
(lavis/datasets/data_utils.py:30-85, run with a stub decord.VideoReader of a given length / fps) and the sample dict of its
MomentRetrievalDataset.__getitem__ (lavis/datasets/datasets/moment_retrieval_dataset.py:17-60), both loaded by file path.
Run in the build container (needs /root/reference): python tests/golden/make_golden_data.py -> tests/golden/data_golden.json"""
import importlib.util
import json
import os
import random
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MRB_REFERENCE_ROOT", "/root/reference")

INDEX_CASES = [  # vlen, fps, n_frms, sampling, clip_proposal, seed
    (4500, 30.0, 60, "uniform", None, 0),
    (4500, 29.97, 60, "random", None, 1),
    (911, 25.0, 20, "random", None, 2),
    (911, 25.0, 20, "uniform", [3.5, 21.25], 3),
    (300, 30.0, 60, "random", [2.0, 4.0], 4),        # 60 frames wanted from a 60-frame window: empty ranges
    (40, 24.0, 60, "uniform", None, 5),              # shorter than n_frms: n_frms clamps to vlen
    (40, 24.0, 60, "random", None, 6),
    (3600, 30.0, 120, "uniform", [-5.0, 500.0], 7),  # proposal clipped to the video
    (1000, 30.0, 8, "headtail", None, 8),
    (7201, 23.976, 60, "random", [10.0, 160.0], 9),
]


class _FakeReader:
    vlen, fps = 0, 0.0

    def __init__(self, uri, height=-1, width=-1):
        self.h, self.w = (height if height > 0 else 8), (width if width > 0 else 8)

    def __len__(self):
        return _FakeReader.vlen

    def get_avg_fps(self):
        return _FakeReader.fps

    def get_batch(self, indices):
        # T, H, W, C with the frame index written into every pixel (mod 256) so that the dataset's permute can be checked
        t = torch.tensor(indices, dtype=torch.float32).remainder(256).view(-1, 1, 1, 1)
        return t.expand(len(indices), self.h, self.w, 3).to(torch.uint8)


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    for name in ("lavis", "lavis.common", "lavis.datasets", "lavis.datasets.datasets", "webdataset"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules.setdefault(name, m)
    dec = types.ModuleType("decord")
    dec.VideoReader = _FakeReader
    dec.bridge = types.SimpleNamespace(set_bridge=lambda *_: None)
    sys.modules["decord"] = dec
    reg = types.ModuleType("lavis.common.registry")
    reg.registry = types.SimpleNamespace(get=lambda k: {"MAX_INT": sys.maxsize}[k])
    sys.modules["lavis.common.registry"] = reg
    _load("lavis.datasets.datasets.base_dataset", "lavis/datasets/datasets/base_dataset.py")
    du = _load("lavis.datasets.data_utils", "lavis/datasets/data_utils.py")
    mrd = _load("lavis.datasets.datasets.moment_retrieval_dataset", "lavis/datasets/datasets/moment_retrieval_dataset.py")

    out = {"indices": [], "samples": []}
    for vlen, fps, n, sampling, clip, seed in INDEX_CASES:
        _FakeReader.vlen, _FakeReader.fps = vlen, fps
        random.seed(seed)
        frms, idx, f = du.load_video("x.mp4", n_frms=n, height=4, width=4, sampling=sampling, clip_proposal=clip)
        assert tuple(frms.shape) == (3, len(idx), 4, 4)
        out["indices"].append({"vlen": vlen, "fps": fps, "n_frms": n, "sampling": sampling, "clip": clip, "seed": seed,
                               "indices": [int(i) for i in idx], "fps_out": f})

    anns = [
        {"qid": 11, "video": "vidA", "query": "A man is cooking pasta in the kitchen.", "duration": 150, "relevant_windows": [[14, 36]]},
        {"qid": 7, "video": "vidB", "query": "the dog runs", "duration": 126.4, "relevant_windows": [[0, 4], [88, 102]]},
        {"qid": 3, "video": "vidC", "query": "clip proposal", "duration": 30.0, "relevant_windows": [[2.5, 9.1]], "start": 12.0, "end": 42.0},
    ]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "train.json")
        json.dump(anns, open(path, "w"))

        def vis(video_path, clip_proposal=None):
            frms, idx, f = du.load_video(video_path, n_frms=6, height=4, width=4, sampling="uniform", clip_proposal=clip_proposal)
            return frms, idx, f

        ds = mrd.MomentRetrievalDataset(vis, None, "/data/videos", [path])
        _FakeReader.vlen, _FakeReader.fps = 4500, 29.97
        for i in range(len(ds)):
            s = ds[i]
            rec = {k: v for k, v in s.items() if k not in ("video", "timestamps", "duration")}
            rec["timestamps"] = [float(x) for x in s["timestamps"]]
            rec["timestamps_dtype"] = str(s["timestamps"].dtype)
            rec["duration"] = float(s["duration"])
            rec["duration_dtype"] = str(s["duration"].dtype)
            rec["video_shape"] = list(s["video"].shape)
            rec["video_dtype"] = str(s["video"].dtype)
            rec["video_frame_values"] = [float(x) for x in s["video"][:, 0, 0, 0]]
            out["samples"].append(rec)
        batch = ds.collater([ds[0], ds[1]])
        out["collated"] = {"video_shape": list(batch["video"].shape), "timestamps_shape": list(batch["timestamps"].shape),
                           "duration": [float(x) for x in batch["duration"]], "duration_dtype": str(batch["duration"].dtype),
                           "query_id": batch["query_id"].tolist(), "query_id_dtype": str(batch["query_id"].dtype),
                           "relevant_windows": list(batch["relevant_windows"])}
    out["annotations"] = anns
    json.dump(out, open(os.path.join(HERE, "data_golden.json"), "w"), indent=0)
    print("wrote", len(out["indices"]), "index cases,", len(out["samples"]), "samples")


if __name__ == "__main__":
    main()
