"""T5 tokenizer access for the BLIP2_MR surface (reference: blip2_mr.py:143 T5TokenizerFast).

`load_t5_tokenizer` returns the real HuggingFace tokenizer when its files are available locally and
otherwise a deterministic stand-in with the same call surface.  The build/bench environment has no
network and no `google/flan-t5-xl` files, so parity and throughput runs use `SyntheticT5Tokenizer`:
same vocabulary size and special ids (pad 0, eos 1, unk 2, <extra_id_k> = 32099-k), every integer
0..999 is ONE token (so the reference's "annoying number" replacement table, blip2_mr.py:1497-1559,
is empty), punctuation is one token per character, and other words hash into [1200, 32000).
"""
import re
import zlib

import logging

import torch

_PUNCT = "[](),.:;?!<>/\\-+=*&%$#@'\"_{}|~^`"
_NUM_BASE = 100          # ids 100..1099  <-> integers 0..999
_PUNCT_BASE = 10         # ids 10..(10+len(_PUNCT))
_WORD_LO, _WORD_HI = 1200, 32000
_PIECE = re.compile(r"<extra_id_\d+>|</s>|<pad>|\d+|[A-Za-z]+|[^\sA-Za-z\d]")


class BatchEncoding(dict):
    """dict with attribute access and .to(device), like transformers.BatchEncoding."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, device):
        return BatchEncoding({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.items()})


class SyntheticT5Tokenizer:
    vocab_size = 32100
    pad_token_id = 0
    eos_token_id = 1
    unk_token_id = 2
    pad_token = "<pad>"
    eos_token = "</s>"

    def __init__(self):
        self._words = {}

    # -- pieces <-> ids ------------------------------------------------------------------------------
    def _piece_to_ids(self, p):
        if p == "</s>":
            return [self.eos_token_id]
        if p == "<pad>":
            return [self.pad_token_id]
        if p.startswith("<extra_id_"):
            return [32099 - int(p[10:-1])]
        if p.isdigit():
            out, s = [], p
            # integers below 1000 without leading zeros are one token; anything else splits greedily
            while s:
                take = s[:3] if (len(s) >= 3 and s[0] != "0") else (s[:2] if (len(s) >= 2 and s[0] != "0") else s[:1])
                out.append(_NUM_BASE + int(take))
                s = s[len(take):]
            return out
        if len(p) == 1 and p in _PUNCT:
            return [_PUNCT_BASE + _PUNCT.index(p)]
        if p.isalpha():
            i = _WORD_LO + zlib.crc32(p.encode("utf-8")) % (_WORD_HI - _WORD_LO)
            self._words.setdefault(i, p)
            return [i]
        return [self.unk_token_id]

    def _id_to_piece(self, i):
        if i == self.eos_token_id:
            return "</s>"
        if i == self.pad_token_id:
            return "<pad>"
        if _NUM_BASE <= i < _NUM_BASE + 1000:
            return str(i - _NUM_BASE)
        if _PUNCT_BASE <= i < _PUNCT_BASE + len(_PUNCT):
            return _PUNCT[i - _PUNCT_BASE]
        if 32000 <= i <= 32099:
            return "<extra_id_%d>" % (32099 - i)
        if i in self._words:
            return self._words[i]
        return "<unk>" if i == self.unk_token_id else "<w%d>" % i

    def convert_tokens_to_ids(self, tok):
        if isinstance(tok, (list, tuple)):
            return [self.convert_tokens_to_ids(t) for t in tok]
        ids = self._piece_to_ids(tok)
        return ids[0] if len(ids) == 1 else self.unk_token_id

    def encode(self, text, add_special_tokens=True):
        ids = []
        for p in _PIECE.findall(text):
            ids.extend(self._piece_to_ids(p))
        if add_special_tokens:
            ids.append(self.eos_token_id)
        return ids

    def __call__(self, text, padding=False, truncation=False, max_length=None, add_special_tokens=True,
                 return_tensors=None, **_):
        single = isinstance(text, str)
        rows = [self.encode(t, add_special_tokens) for t in ([text] if single else text)]
        if truncation and max_length is not None:
            rows = [r[:max_length - 1] + [self.eos_token_id] if (len(r) > max_length and add_special_tokens)
                    else r[:max_length] for r in rows]
        masks = [[1] * len(r) for r in rows]
        if padding in (True, "longest"):
            L = max((len(r) for r in rows), default=0)
            masks = [m + [0] * (L - len(m)) for m in masks]
            rows = [r + [self.pad_token_id] * (L - len(r)) for r in rows]
        if return_tensors == "pt":
            return BatchEncoding(input_ids=torch.tensor(rows, dtype=torch.long).reshape(len(rows), -1),
                                 attention_mask=torch.tensor(masks, dtype=torch.long).reshape(len(rows), -1))
        if single:
            rows, masks = rows[0], masks[0]
        return BatchEncoding(input_ids=rows, attention_mask=masks)

    def decode(self, ids, skip_special_tokens=False):
        if torch.is_tensor(ids):
            ids = ids.reshape(-1).tolist()
        if isinstance(ids, int):
            ids = [ids]
        out, prev_word = [], False
        for i in ids:
            if skip_special_tokens and i in (self.pad_token_id, self.eos_token_id):
                continue
            p = self._id_to_piece(int(i))
            is_word = p[:1].isalnum() or p.startswith("<")
            if out and ((is_word and prev_word) or out[-1] == ","):
                out.append(" ")
            out.append(p)
            prev_word = is_word
        return "".join(out)

    def batch_decode(self, seqs, skip_special_tokens=False):
        if torch.is_tensor(seqs):
            seqs = seqs.tolist()
        return [self.decode(s, skip_special_tokens) for s in seqs]


def load_t5_tokenizer(name="google/flan-t5-xl", allow_synthetic=True):
    """Real tokenizer when its files are cached locally (never touches the network).  transformers 5.x returns an EMPTY tokenizer
    instead of raising when files are missing, so the result is validated before it is trusted.  When it is not there:
    allow_synthetic=False raises (what training / evaluation from a recipe uses: a run on hashed word ids is garbage), otherwise the
    synthetic stand-in is returned WITH a warning (parity tests and the benchmark, which have no tokenizer files offline)."""
    why = "not validated"
    try:
        from transformers import T5TokenizerFast
        tok = T5TokenizerFast.from_pretrained(name, local_files_only=True)
        ids = tok("the video shows 25 seconds")["input_ids"]
        if len(tok) >= 32000 and tok.unk_token_id not in ids[:-1] and len(ids) >= 5:
            return tok
        why = "the local files give an empty / wrong vocabulary (%d entries)" % len(tok)
    except Exception as e:
        why = "%s: %s" % (type(e).__name__, str(e)[:120])
    if not allow_synthetic:
        raise RuntimeError("T5 tokenizer %r is not available locally (%s); refusing to fall back to the synthetic tokenizer. "
                           "Point model.t5_model at a local FlanT5 directory, or set model.allow_synthetic: true for dry runs" % (name, why))
    logging.getLogger(__name__).warning("T5 tokenizer %r not available locally (%s): using the SYNTHETIC tokenizer (hashed word ids) -- "
                                        "fine for parity tests / benchmarks, useless for real training or evaluation", name, why)
    return SyntheticT5Tokenizer()
