"""Beam search for BLIP2_MR.generate with the semantics of transformers==4.46.1
GenerationMixin._beam_search / BeamSearchScorer for the arguments blip2_mr.py:883-899 passes
(do_sample False, early_stopping False, length_penalty, min_length, max_new_tokens; eos 1, pad 0,
decoder_start 0), on top of T5Engine's cached-K/V incremental decoder.  The device does the decoder
step and the top-2k selection; hypothesis bookkeeping (a few integers per beam) stays on the host.
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
                order = sorted((s, i) for i, (s, _) in enumerate(self.beams))
                del self.beams[order[0][1]]
                self.worst = order[1][0]
            else:
                self.worst = min(score, self.worst)

    def is_done(self, best_sum_logprobs, cur_len, prompt_len):
        if len(self.beams) < self.num_beams:
            return False
        return self.worst >= best_sum_logprobs / (cur_len - prompt_len) ** self.lp


@torch.no_grad()
def beam_search(t5, inputs_embeds, attention_mask, num_beams=5, max_new_tokens=50, min_length=1, length_penalty=1.0,
                eos_id=1, pad_id=0, start_id=0):
    B, Le, _ = inputs_embeds.shape
    nb = num_beams
    dev = inputs_embeds.device                                           # bookkeeping tensors follow the engine's device
    enc_ext, kmask = t5.encode(inputs_embeds, attention_mask)
    max_length = max_new_tokens + 1
    st = t5.init_decode(enc_ext, B, Le, nb, max_length)
    ids = [[start_id] for _ in range(B * nb)]
    beam_scores = torch.zeros((B, nb), dtype=torch.float32, device=dev)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    hyps = [_Hyps(nb, length_penalty) for _ in range(B)]
    done = [False] * B
    tokens = torch.full((B * nb,), start_id, dtype=torch.long, device=dev)
    cur = 1
    reorder = None                                                       # beam index each cache row continues from
    while True:
        logits = t5.decode_step(st, tokens, cur - 1, kmask, beam_idx=reorder)   # [B*nb, V] fp32 (static buffer)
        scores = torch.log_softmax(logits, dim=-1)
        if cur < min_length:
            scores[:, eos_id] = -float("inf")
        V = scores.shape[-1]
        scores = (scores + beam_scores[:, None]).view(B, nb * V)
        top_s, top_i = torch.topk(scores, 2 * nb, dim=1, largest=True, sorted=True)
        top_s_c, top_i_c = top_s.cpu(), top_i.cpu()
        src_beam, tok = top_i_c // V, top_i_c % V
        cur_len = cur + 1
        nxt_scores = torch.zeros(B, nb)
        nxt_tok = torch.zeros(B, nb, dtype=torch.long)
        nxt_idx = torch.zeros(B, nb, dtype=torch.long)
        for b in range(B):
            if done[b]:
                nxt_tok[b] = pad_id
                nxt_idx[b] = b * nb
                continue
            k = 0
            for rank in range(2 * nb):
                t, s, bi = tok[b, rank].item(), top_s_c[b, rank].item(), b * nb + src_beam[b, rank].item()
                if t == eos_id:
                    if rank >= nb:
                        continue
                    hyps[b].add(list(ids[bi]), s, cur_len - 1)
                else:
                    nxt_scores[b, k], nxt_tok[b, k], nxt_idx[b, k] = s, t, bi
                    k += 1
                if k == nb:
                    break
            done[b] = done[b] or hyps[b].is_done(top_s_c[b].max().item(), cur_len, 1)
        flat_idx = nxt_idx.view(-1)
        ids = [ids[i] + [t] for i, t in zip(flat_idx.tolist(), nxt_tok.view(-1).tolist())]
        beam_scores = nxt_scores.view(-1).to(dev)
        tokens = nxt_tok.view(-1).to(dev)
        cur += 1
        if all(done) or cur >= max_length:
            break
        reorder = flat_idx.to(dev)
    for b in range(B):
        if done[b]:
            continue
        for j in range(nb):
            bi = b * nb + j
            hyps[b].add(ids[bi], beam_scores[bi].item(), len(ids[bi]) - 1)
    best = [sorted(h.beams, key=lambda x: x[0])[-1][1] for h in hyps]
    sent_max = min(max(len(x) for x in best) + 1, max_length)
    out = torch.full((B, sent_max), pad_id, dtype=torch.long)
    for b, hyp in enumerate(best):
        out[b, :len(hyp)] = torch.tensor(hyp)
        if len(hyp) < sent_max:
            out[b, len(hyp)] = eos_id
    return out
