"""Full-size (ViT-g + Q-Former + FlanT5-XL widths AND depths, BASELINE.json configs[1]: batch 4, 60 frames) checks through
size-independent properties -- the fp32 CPU oracle needs minutes per clip at this size:
  * permutation of the clips in the batch leaves the loss and every gradient unchanged (clips are independent; the loss
    is a mean over all target tokens),
  * a replayed CUDA graph of the step reproduces the eager step, and two replays agree,
  * gradient hand-over is linear in the incoming gradient (GradScaler factor).
Tolerances: only the order of fp32 atomic adds differs between the compared runs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _relfro(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def full_model():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from mr_blip_b200.dims import FULL, init_state_dict
    sd = init_state_dict(FULL, seed=1234, lora_b_std=0.02, device="cuda")
    m = BLIP2_MR(dims=FULL, state_dict=sd).cuda().train()
    del sd
    yield m
    del m
    torch.cuda.empty_cache()


def _step(model, samples, scale=1.0):
    for p in model.parameters():
        p.grad = None
    loss = model(samples)["loss"]
    (loss * scale).backward()
    return loss.item(), model.flat_grads().clone()


def test_qvh_full_size_properties(full_model):
    from oracle import synth
    model = full_model
    s = synth.make_samples(batch=4, frames=60, query_words=32, seed=100)
    s["video"] = s["video"].cuda()
    model.cuda_graphs = False
    loss_e, g_e = _step(model, s)
    assert loss_e == loss_e and 5.0 < loss_e < 15.0          # ln(32128) = 10.4 for an untrained head
    assert torch.isfinite(g_e).all() and g_e.abs().max().item() > 0
    # clip permutation
    perm = [2, 0, 3, 1]
    sp = {k: (v[perm] if torch.is_tensor(v) else [v[i] for i in perm]) for k, v in s.items()}
    loss_p, g_p = _step(model, sp)
    assert abs(loss_p - loss_e) < 2e-5 * abs(loss_e)
    assert _relfro(g_p, g_e) < 2e-3
    # graph replay == eager, replay == replay, hand-over is linear in the incoming gradient
    model.cuda_graphs = True
    model.reset_graphs()
    outs = [_step(model, s) for _ in range(4)]               # eager, capture + replay, replay, replay
    for loss_g, g_g in outs:
        assert abs(loss_g - loss_e) < 2e-5 * abs(loss_e)
        assert _relfro(g_g, g_e) < 2e-3
    assert list(model._steps.values())[0].graph is not None
    loss_s, g_s = _step(model, s, scale=8.0)
    assert _relfro(g_s, 8.0 * outs[-1][1]) < 2e-3
    model.reset_graphs()
