"""The two-stage video-QA branch (blip2_mr.py:309-431, 990-1099, 1233-1314) on the device against the oracle restatement.
Its host logic is also checked on the CPU over op stand-ins (tests/test_host_logic.py::test_video_qa_branch_host_logic_with_emulated_ops);
first run on a B200 in round 2 (profiles/r02_call1.md: green)."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]

from mr_blip_b200.dims import ANSWERER_PREFIX, TINY, add_answerer  # noqa: E402


def _relfro(got, want):
    got, want = torch.as_tensor(got).float().cpu(), torch.as_tensor(want).float().cpu()
    return ((got - want).norm() / want.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("task", ["qformer_freeze_lora_QA", "qformer_freeze_lora_QA_with_localizer"])
def test_video_qa_branch_vs_oracle(tiny_sd, task):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mr_blip_b200 import mr_utils
    from mr_blip_b200.blip2_mr import BLIP2_MR
    from oracle import blip2_mr as ob
    from test_host_logic import _qa_samples
    sd = add_answerer(dict(tiny_sd), TINY, seed=1234, lora_b_std=0.02)
    model = BLIP2_MR(dims=TINY, state_dict=sd, task=task, num_frames_for_answer=3, train_dropout=False).cuda().train()
    samples = _qa_samples()
    res = model(dict(samples))
    res["loss"].backward()
    osd = dict(sd)
    leaves = {k: osd[k].clone().requires_grad_(True) for k in osd if "lora_" in k and k.startswith(ANSWERER_PREFIX)}
    osd.update(leaves)
    o = ob.forward_qa(osd, TINY, model.t5_tokenizer, samples, use_localizer="with_localizer" in task, n_frames=3,
                      post_process=mr_utils.post_process)
    o["loss"].backward()
    assert abs(res["loss"].item() - o["loss"].item()) < 5e-3
    for k, leaf in leaves.items():
        assert _relfro(model._get(k).grad, leaf.grad) < 4e-2, k
    assert all(p.grad is None for n, p in model.named_parameters() if not n.startswith(ANSWERER_PREFIX))
    model.eval()
    out = model.videoQA_generate(dict(samples))
    moments, rel = ob._qa_relevant_frames(sd, TINY, model.t5_tokenizer, dict(samples, relevant_windows=[[0, 0]], query_id=samples["question_id"]),
                                          "with_localizer" in task, 3, mr_utils.post_process, None)
    want, scores = ob.videoqa_answer(sd, TINY, model.t5_tokenizer, samples, rel)
    assert out["relevant_moments"] == [moments] and out["output_text"] == want
    assert _relfro(out["answer_scores"], scores) < 2e-2
