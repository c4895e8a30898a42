"""Input side of the path for stand-alone use: frame sampling, the sample dict BLIP2_MR.forward / generate consume, and
batching -- the data format of lavis/datasets/data_utils.py:30-85 (load_video), lavis/datasets/datasets/
moment_retrieval_dataset.py:17-60 (MomentRetrievalDataset.__getitem__) and BaseDataset.collater.  Inside LAVIS the runner
keeps its own dataset classes; nothing here is on the device path.

Frames can stay uint8 ([T,3,H,W], `uint8=True`): BLIP2_MR then normalises inside the patch-extraction kernel and the
host->device copy is 4x smaller.  Decoding uses OpenCV (the reference's decord is not in this image)."""
import json
import os
import random as _random

import numpy as np
import torch

VIDEO_PROMPT_END = "<extra_id_0>"
TASK_PROMPT = "Given the video and the query, find the relevant windows.\nRelevant windows: "
PIXEL_MEAN = (0.48145466, 0.4578275, 0.40821073)       # blip_processors.py:24-27 (CLIP statistics)
PIXEL_STD = (0.26862954, 0.26130258, 0.27577711)


def sample_frame_indices(vlen, fps, n_frms, sampling="uniform", clip_proposal=None, rng=None):
    """Which frames of a vlen-frame video a clip is made of (data_utils.py:38-80).  The window [start, end) -- the whole
    video, or clip_proposal seconds x fps clipped to it -- is cut into n_frms equal integer segments: "uniform" takes each
    segment's middle frame, "random" one frame drawn from each (rng.choice, so the draws reproduce the reference's under the
    same seed), "headtail" n/2 sorted draws from each half of the video.  Fewer than n_frms (short video) repeats the last."""
    rng = rng or _random
    n = min(int(n_frms), vlen)
    lo, hi = 0, vlen
    if clip_proposal is not None:
        lo, hi = max(int(clip_proposal[0] * fps), 0), min(int(clip_proposal[1] * fps), vlen)
    cuts = np.linspace(start=lo, stop=hi, num=n + 1).astype(int)
    segs = list(zip(cuts[:-1].tolist(), cuts[1:].tolist()))
    if sampling == "uniform":
        picks = [min((a + b) // 2, vlen - 1) for a, b in segs]
    elif sampling == "random":
        picks = [a if a == b else rng.choice(range(a, b)) for a, b in segs]
    elif sampling == "headtail":
        half = vlen // 2
        picks = sorted(rng.sample(range(half), n // 2)) + sorted(rng.sample(range(half, vlen), n // 2))
    else:
        raise NotImplementedError(sampling)
    return picks + [picks[-1]] * (n - len(picks))


def frame_timestamps(indices, fps):
    """Seconds of each sampled frame, rounded to 2 decimals, float32 (moment_retrieval_dataset.py:43-47)."""
    return torch.tensor([round(float(i / fps), 2) for i in indices])


class Cv2VideoReader:
    """len / get_avg_fps / get_batch(indices) -> uint8 [T,H,W,3] RGB, the part of decord.VideoReader load_video uses.
    Frames are reached by sequential grab() from the smallest wanted index (container seeks are not frame-exact)."""

    def __init__(self, uri, height=-1, width=-1):
        import cv2
        self._cv2 = cv2
        self.uri = uri
        cap = cv2.VideoCapture(uri)
        if not cap.isOpened():
            raise RuntimeError("cannot open video %s" % uri)
        self._n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self._fps = float(cap.get(cv2.CAP_PROP_FPS))
        cap.release()
        self.size = (width, height) if height > 0 and width > 0 else None

    def __len__(self):
        return self._n

    def get_avg_fps(self):
        return self._fps

    def get_batch(self, indices):
        cv2 = self._cv2
        want = sorted(set(int(i) for i in indices))
        cap = cv2.VideoCapture(self.uri)
        got, pos = {}, 0
        if want and want[0] > 0:
            cap.set(cv2.CAP_PROP_POS_FRAMES, 0)
        for target in want:
            while pos < target:
                if not cap.grab():
                    break
                pos += 1
            ok, frame = cap.read()
            pos += 1
            if not ok:
                if not got:
                    raise RuntimeError("cannot decode frame %d of %s" % (target, self.uri))
                frame = None
            else:
                frame = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
                if self.size is not None and (frame.shape[1], frame.shape[0]) != self.size:
                    frame = cv2.resize(frame, self.size, interpolation=cv2.INTER_LINEAR)
            got[target] = frame if frame is not None else got[max(got)]     # truncated stream: repeat the last good frame
        cap.release()
        return torch.from_numpy(np.stack([got[int(i)] for i in indices]))


def random_resized_crop(clip, size, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), mode="bicubic"):
    """The train processor's augmentation (blip_processors.py:302-315 -> transforms_video.py:53-83): one random window for
    the whole clip -- area fraction in `scale`, log-uniform aspect in `ratio`, drawn from torch's global generator by
    torchvision's RandomResizedCrop.get_params -- resized to size x size with torch's bicubic kernel, then truncated to uint8
    as ToUint8 does.  clip: [3,T,H,W] (uint8 or float) -> uint8 [3,T,size,size]."""
    from torchvision.transforms import RandomResizedCrop
    clip = clip.float()
    i, j, h, w = RandomResizedCrop.get_params(clip, list(scale), list(ratio))
    out = torch.nn.functional.interpolate(clip[..., i:i + h, j:j + w], size=(size, size), mode=mode, align_corners=False)
    return out.to(torch.uint8)


class VideoProcessor:
    """Callable (video_path, clip_proposal) -> (frames, indices, fps), the contract of Blip2VideoTrainProcessor /
    BlipVideoEvalProcessor (blip_processors.py:287-393) as the dataset uses it.  Frames are decoded at image_size x image_size;
    augment=True applies the train processor's RandomResizedCrop (scale 0.5-1, bicubic).  uint8=False returns the reference
    layout -- float32 [3,T,H,W], scaled to [0,1] and CLIP-normalised; uint8=True returns the uint8 [3,T,H,W] the reference has
    just before ToTensorVideo, for the fused normalisation on the device (same values, a quarter of the bytes)."""

    def __init__(self, image_size=224, n_frms=60, sampling="uniform", uint8=False, reader=Cv2VideoReader, rng=None,
                 mean=PIXEL_MEAN, std=PIXEL_STD, augment=False, min_scale=0.5, max_scale=1.0):
        self.image_size, self.n_frms, self.sampling, self.uint8 = image_size, n_frms, sampling, uint8
        self.reader, self.rng = reader, rng
        self.augment, self.scale = augment, (min_scale, max_scale)
        self.mean = torch.tensor(mean).view(3, 1, 1, 1)
        self.std = torch.tensor(std).view(3, 1, 1, 1)

    def __call__(self, video_path, clip_proposal=None):
        vr = self.reader(video_path, height=self.image_size, width=self.image_size)
        fps = vr.get_avg_fps()
        indices = sample_frame_indices(len(vr), fps, self.n_frms, self.sampling, clip_proposal, self.rng)
        frames = vr.get_batch(indices).permute(3, 0, 1, 2)                  # T,H,W,C -> C,T,H,W
        if self.augment:
            frames = random_resized_crop(frames, self.image_size, self.scale)
        if not self.uint8:
            frames = (frames.float() / 255.0 - self.mean) / self.std        # ToTensorVideo + NormalizeVideo
        return frames, indices, fps


class MomentRetrievalDataset(torch.utils.data.Dataset):
    """Annotation records {"qid", "video", "query", "duration", "relevant_windows"[, "start", "end"]} -> the sample dict
    of moment_retrieval_dataset.py:17-60 (same keys, prompts, dtypes)."""

    def __init__(self, vis_processor, text_processor=None, vis_root="", ann_paths=()):
        self.vis_processor, self.text_processor, self.vis_root = vis_processor, text_processor, vis_root
        self.annotation = []
        for path in ann_paths:
            if ".json" not in path:
                raise AttributeError("Undefined data type")
            with open(path) as f:
                self.annotation.extend(json.load(f))
        for i, ann in enumerate(self.annotation):                          # base_dataset.py:56-61
            if not isinstance(ann, str):
                ann["instance_id"] = str(i)

    def __len__(self):
        return len(self.annotation)

    def __getitem__(self, index):
        ann = self.annotation[index]
        clip = [float(ann["start"]), float(ann["end"])] if "start" in ann else None
        frames, indices, fps = self.vis_processor(os.path.join(self.vis_root, ann["video"] + ".mp4"), clip_proposal=clip)
        return {"video": frames.permute(1, 0, 2, 3),                       # C,T,H,W -> T,C,H,W
                "duration": torch.tensor(ann["duration"]),
                "query_id": ann["qid"],
                "timestamps": frame_timestamps(indices, fps),
                "video_prompt_end": VIDEO_PROMPT_END,
                "query_prompt": "Query: " + ann["query"] + "\n",
                "task_prompt": TASK_PROMPT,
                "relevant_windows": str(ann["relevant_windows"])}

    @staticmethod
    def collater(samples):
        return torch.utils.data.dataloader.default_collate(samples)
