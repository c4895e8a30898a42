"""Restatement of HuggingFace transformers==4.46.1 beam search (GenerationMixin._beam_search +
BeamSearchScorer / BeamHypotheses) for the arguments BLIP2_MR.generate passes
(blip2_mr.py:883-899: num_beams=5, max_new_tokens=50, min_length=1, length_penalty=1.0,
repetition_penalty=1.0, do_sample=False, early_stopping=False (config default), eos=1, pad=0,
decoder_start=0).  transformers 4.46.1 is a third-party dependency pinned in requirements.txt:56
and NOT vendored under /root/reference.  Pinned instead against the INSTALLED transformers (5.5.0):
tests/test_oracle_golden.py::test_beam_search_matches_transformers_generate runs
T5ForConditionalGeneration.generate on random tiny T5s (beams 1-5, min_length, length penalties, padded
encoder rows, finished and length-capped hypotheses) and requires token-for-token agreement up to the first
eos -- 160 sequences, no mismatch.  Against 4.46.1 itself: unpinned.  Test infrastructure only.
"""
import torch


class _Hyps:
    def __init__(self, num_beams, length_penalty):
        self.num_beams, self.lp = num_beams, length_penalty
        self.beams, self.worst = [], 1e9

    def add(self, hyp, sum_logprobs, generated_len):
        score = sum_logprobs / (generated_len ** self.lp)
        if len(self.beams) < self.num_beams or score > self.worst:
            self.beams.append((score, hyp))
            if len(self.beams) > self.num_beams:
                srt = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[srt[0][1]]
                self.worst = srt[1][0]
            else:
                self.worst = min(score, self.worst)

    def is_done(self, best_sum_logprobs, cur_len, prompt_len):
        if len(self.beams) < self.num_beams:
            return False
        # early_stopping=False heuristic
        return self.worst >= best_sum_logprobs / (cur_len - prompt_len) ** self.lp


def beam_search(step_logits_fn, batch, num_beams=5, max_new_tokens=50, min_length=1, length_penalty=1.0,
                eos_id=1, pad_id=0, start_id=0):
    """step_logits_fn(input_ids [batch*num_beams, cur]) -> next-token logits [batch*num_beams, V] (fp32).
    Returns LongTensor [batch, <= max_new_tokens+1] (starts with decoder_start, eos-terminated, pad-filled)."""
    nb = num_beams
    ids = torch.full((batch * nb, 1), start_id, dtype=torch.long)
    beam_scores = torch.zeros(batch, nb)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    hyps = [_Hyps(nb, length_penalty) for _ in range(batch)]
    done = [False] * batch
    max_length = max_new_tokens + 1
    prompt_len = 1
    while True:
        logits = step_logits_fn(ids).float()
        scores = torch.log_softmax(logits, dim=-1)
        cur = ids.shape[-1]
        if cur < min_length:                               # MinLengthLogitsProcessor
            scores[:, eos_id] = -float("inf")
        scores = scores + beam_scores[:, None]
        V = scores.shape[-1]
        scores = scores.view(batch, nb * V)
        top_s, top_i = torch.topk(scores, 2 * nb, dim=1, largest=True, sorted=True)
        src_beam, tok = top_i // V, top_i % V
        # BeamSearchScorer.process
        cur_len = cur + 1
        nxt_scores = torch.zeros(batch, nb)
        nxt_tok = torch.zeros(batch, nb, dtype=torch.long)
        nxt_idx = torch.zeros(batch, nb, dtype=torch.long)
        for b in range(batch):
            if done[b]:
                nxt_scores[b], nxt_tok[b], nxt_idx[b] = 0, pad_id, 0
                continue
            k = 0
            for rank in range(2 * nb):
                t, s, bi = tok[b, rank].item(), top_s[b, rank].item(), b * nb + src_beam[b, rank].item()
                if t == eos_id:
                    if rank >= nb:
                        continue
                    hyps[b].add(ids[bi].clone(), s, cur_len - prompt_len)
                else:
                    nxt_scores[b, k], nxt_tok[b, k], nxt_idx[b, k] = s, t, bi
                    k += 1
                if k == nb:
                    break
            done[b] = done[b] or hyps[b].is_done(top_s[b].max().item(), cur_len, prompt_len)
        beam_scores = nxt_scores.view(-1)
        ids = torch.cat([ids[nxt_idx.view(-1)], nxt_tok.view(-1, 1)], dim=-1)
        if all(done) or ids.shape[-1] >= max_length:
            break
    # finalize
    for b in range(batch):
        if done[b]:
            continue
        for j in range(nb):
            bi = b * nb + j
            hyps[b].add(ids[bi], beam_scores[bi].item(), ids.shape[-1] - prompt_len)
    best = [sorted(h.beams, key=lambda x: x[0])[-1][1] for h in hyps]
    sent_max = min(max(len(x) for x in best) + 1, max_length)
    out = torch.full((batch, sent_max), pad_id, dtype=torch.long)
    for b, hyp in enumerate(best):
        out[b, :len(hyp)] = hyp
        if len(hyp) < sent_max:
            out[b, len(hyp)] = eos_id
    return out
