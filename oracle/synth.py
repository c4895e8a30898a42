"""Seeded synthetic `samples` dicts of the shape MomentRetrievalDataset yields
(lavis/datasets/datasets/moment_retrieval_dataset.py:17-60).  Test infrastructure."""
import torch

TASK_PROMPT = " Given the video and the query, find the relevant windows.\nRelevant windows: "
VIDEO_PROMPT_END = "<extra_id_0>\n"
_WORDS = ("a person is cooking pasta in the kitchen while the dog watches and then they walk outside to "
          "the garden where two friends talk about the weather before driving to town").split()


def make_samples(batch, frames, query_words=8, duration=150.0, seed=0, img=224, spread=7.0):
    """spread: seconds by which clip i is shorter than clip i-1 (ragged prompts); needs duration - spread * (batch-1) > 20."""
    g = torch.Generator().manual_seed(seed)
    video = torch.randn(batch, frames, 3, img, img, generator=g)
    durs = torch.tensor([duration - spread * i for i in range(batch)])
    ts = torch.stack([torch.linspace(0.5 * d / frames, d - 0.5 * d / frames, frames) for d in durs.tolist()])
    ts = (ts * 100).round() / 100
    queries, answers = [], []
    for i in range(batch):
        idx = torch.randint(0, len(_WORDS), (query_words,), generator=g).tolist()
        queries.append("Query: " + " ".join(_WORDS[j] for j in idx))
        s = int(torch.randint(0, int(durs[i]) - 20, (1,), generator=g))
        e = s + int(torch.randint(2, 18, (1,), generator=g))
        answers.append("[[%d, %d]]" % (s, e) if i % 2 == 0 else "[[%d, %d], [%d, %d]]" % (s, e, e + 1, e + 2))
    return {"video": video, "timestamps": ts, "duration": durs,
            "query_id": list(range(batch)), "video_prompt_end": [VIDEO_PROMPT_END] * batch,
            "query_prompt": queries, "task_prompt": [TASK_PROMPT] * batch, "relevant_windows": answers}
