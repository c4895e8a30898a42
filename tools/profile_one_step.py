"""One full-size QVH training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import sys

import torch

sys.path.insert(0, ".")
from mr_blip_b200.blip2_mr import BLIP2_MR  # noqa: E402
from mr_blip_b200.dims import FULL, init_state_dict  # noqa: E402
from oracle import synth  # noqa: E402

sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
model = BLIP2_MR(dims=FULL, state_dict=sd).cuda().train()
del sd
samples = synth.make_samples(batch=4, frames=60, query_words=32, seed=100)
samples["video"] = samples["video"].cuda()
def zero():                             # as optimizer.zero_grad(set_to_none=True) does in the real loop: without it autograd ADDS
    for p in model.parameters():        # the handed-over gradients to the old ones (868 ATen add launches in round 1's list)
        p.grad = None


for _ in range(3):                      # eager, capture, replay: the profiled step replays the CUDA graph (ncu profiles its kernel nodes)
    zero()
    model(samples)["loss"].backward()
zero()
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(samples)["loss"].backward()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
