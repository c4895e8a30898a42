"""Golden output of the REFERENCE's train-time video transform (lavis/processors/transforms_video.py RandomResizedCropVideo
+ blip_processors.py ToTHWC / ToUint8, loaded by file path) on a smooth synthetic clip with a seeded torch generator.
Run in the build container (needs /root/reference): python tests/golden/make_golden_crop.py -> tests/golden/crop_golden.npz"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MRB_REFERENCE_ROOT", "/root/reference")


def synth_clip(T=5, H=48, W=64):
    """Smooth mid-range frames (no bicubic overshoot past 0 / 255): float32 [3,T,H,W] holding integer values."""
    y, x = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    frames = []
    for t in range(T):
        chans = [120 + 60 * torch.sin(0.11 * x + 0.3 * t + c) + 35 * torch.cos(0.17 * y - 0.2 * c) for c in range(3)]
        frames.append(torch.stack(chans))
    return torch.stack(frames, 1).round()


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    for name in ("lavis", "lavis.processors"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules.setdefault(name, m)
    fv = _load("lavis.processors.functional_video", "lavis/processors/functional_video.py")
    sys.modules["lavis.processors"].functional_video = fv
    tv = _load("lavis.processors.transforms_video", "lavis/processors/transforms_video.py")
    clip = synth_clip()
    out = {}
    for k, (seed, size, scale) in enumerate([(0, 32, (0.5, 1.0)), (7, 24, (0.5, 1.0)), (3, 32, (0.2, 0.6))]):
        torch.manual_seed(seed)
        t = tv.RandomResizedCropVideo(size, scale=scale, interpolation_mode="bicubic")
        y = t(clip)                                           # C,T,h,w float
        y = y.permute(1, 2, 3, 0).to(torch.uint8)             # ToTHWC, ToUint8
        out["case%d" % k] = y.permute(3, 0, 1, 2).numpy()     # back to C,T,H,W for comparison
        out["meta%d" % k] = np.array([seed, size, scale[0] * 1000, scale[1] * 1000], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "crop_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
