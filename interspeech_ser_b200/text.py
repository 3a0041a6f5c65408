"""Text branch of the reference's feature extraction (preprocessing/preprocess_roberta.py): tokenizer + RobertaModel
surface, backed by libserenc's SERENC_ARCH_TEXT encoder.

    tokenizer = RobertaTokenizer.from_pretrained(SSL_TYPE)                         preprocess_roberta.py:103
    text_model = RobertaModel.from_pretrained(SSL_TYPE); .eval(); .to(device)       :104-106
    encoding = tokenizer(text, padding="max_length", truncation=True, max_length=80, return_tensors="pt").to(device)   :49-55
    outputs = model(**encoding, output_hidden_states=True); outputs.hidden_states   :57-62
    model(**encoding).last_hidden_state                                             :70

The tokenizer is GPT-2 / RoBERTa byte-level BPE (vocab.json + merges.txt of the checkpoint directory; there is no
network here, so the files have to be local). The encoder is a post-LN stack at T = max_len: the same tcgen05 GEMM,
LayerNorm and attention kernels as the speech encoders, plus an embedding-gather kernel and a key-length mask.
"""
from __future__ import annotations

import json
import os
from functools import lru_cache
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .configs import EncoderConfig
from .engine import REDUCE_NONE
from .modeling import Extracted, ModelOutput, _Base


# --------------------------------------------------------------------------------------------------
# byte-level BPE (the algorithm of GPT-2's encoder.py, which RobertaTokenizer follows)
# --------------------------------------------------------------------------------------------------
@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """Reversible byte -> printable unicode character table used by byte-level BPE vocabularies."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


_PRETOKENIZE = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""


class BatchEncoding(dict):
    """dict with attribute access and .to(device), like transformers.BatchEncoding."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def to(self, device):
        return BatchEncoding({k: (v.to(device) if hasattr(v, "to") else v) for k, v in self.items()})


class RobertaTokenizer:
    """RobertaTokenizer replacement: `<s> byte-level-BPE(text) </s>`, truncation and max_length padding."""

    def __init__(self, vocab: Union[str, Dict[str, int]], merges: Union[str, Sequence[str]], bos_token="<s>", eos_token="</s>",
                 unk_token="<unk>", pad_token="<pad>", **kw):
        import regex

        if isinstance(vocab, str):
            with open(vocab, encoding="utf-8") as fh:
                vocab = json.load(fh)
        if isinstance(merges, str):
            with open(merges, encoding="utf-8") as fh:
                merges = [ln for ln in fh.read().split("\n") if ln and not ln.startswith("#version")]
        self.encoder: Dict[str, int] = dict(vocab)
        self.bpe_ranks = {tuple(m.split()): i for i, m in enumerate(merges)}
        self.byte_encoder = bytes_to_unicode()
        self.pat = regex.compile(_PRETOKENIZE)
        self.cache: Dict[str, Tuple[str, ...]] = {}
        self.bos_token_id = self.encoder[bos_token]
        self.eos_token_id = self.encoder[eos_token]
        self.pad_token_id = self.encoder[pad_token]
        self.unk_token_id = self.encoder[unk_token]
        self.cls_token_id, self.sep_token_id = self.bos_token_id, self.eos_token_id

    @classmethod
    def from_pretrained(cls, name_or_path: str, **kw) -> "RobertaTokenizer":
        cand = [name_or_path]
        root = os.environ.get("SERENC_WEIGHTS_DIR")
        if root:
            cand += [os.path.join(root, name_or_path), os.path.join(root, name_or_path.split("/")[-1])]
        for c in cand:
            v, m = os.path.join(c, "vocab.json"), os.path.join(c, "merges.txt")
            if os.path.isfile(v) and os.path.isfile(m):
                return cls(v, m, **kw)
        raise OSError(f"No tokenizer files (vocab.json + merges.txt) found for '{name_or_path}' (looked in {cand}); there is no "
                      "network in this environment. Point SERENC_WEIGHTS_DIR at a directory of HF checkpoints.")

    def __len__(self):
        return len(self.encoder)

    def _bpe(self, token: str) -> Tuple[str, ...]:
        hit = self.cache.get(token)
        if hit is not None:
            return hit
        word = tuple(token)
        while len(word) > 1:
            pairs = {(word[i], word[i + 1]) for i in range(len(word) - 1)}
            best = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if best not in self.bpe_ranks:
                break
            a, b = best
            out, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == a and word[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = tuple(out)
        self.cache[token] = word
        return word

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for tok in self.pat.findall(text):
            out.extend(self._bpe("".join(self.byte_encoder[b] for b in tok.encode("utf-8"))))
        return out

    def encode(self, text: str, max_length: Optional[int] = None, truncation: bool = False) -> List[int]:
        ids = [self.encoder.get(t, self.unk_token_id) for t in self.tokenize(text)]
        if truncation and max_length is not None:
            ids = ids[: max(0, max_length - 2)]
        return [self.bos_token_id] + ids + [self.eos_token_id]

    def __call__(self, text: Union[str, Sequence[str]], padding: Union[bool, str] = False, truncation: bool = False,
                 max_length: Optional[int] = None, return_tensors: Optional[str] = None, **kw) -> BatchEncoding:
        single = isinstance(text, str)
        texts = [text] if single else list(text)
        rows = [self.encode(str(t), max_length, truncation) for t in texts]
        if padding == "max_length":
            if max_length is None:
                raise ValueError("padding='max_length' needs max_length")
            width = max_length
        elif padding in (True, "longest"):
            width = max(len(r) for r in rows)
        else:
            width = None
        if width is None and return_tensors and len({len(r) for r in rows}) > 1:
            raise ValueError("rows of different length need padding to be returned as tensors")
        ids, mask = [], []
        for r in rows:
            w = len(r) if width is None else max(width, len(r))
            ids.append(r + [self.pad_token_id] * (w - len(r)))
            mask.append([1] * len(r) + [0] * (w - len(r)))
        if return_tensors == "pt":
            return BatchEncoding(input_ids=torch.tensor(ids, dtype=torch.long), attention_mask=torch.tensor(mask, dtype=torch.long))
        if return_tensors == "np":
            return BatchEncoding(input_ids=np.asarray(ids, dtype=np.int64), attention_mask=np.asarray(mask, dtype=np.int64))
        if single:
            return BatchEncoding(input_ids=ids[0], attention_mask=mask[0])
        return BatchEncoding(input_ids=ids, attention_mask=mask)


# --------------------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------------------
def _valid_lengths(input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], pad_id: int) -> List[int]:
    """Lengths of a RIGHT-padded batch, the only pattern the reference produces (tokenizer(padding='max_length')).
    HF derives position ids from `input_ids != pad` and key masking from attention_mask; both have to describe the
    same prefix here."""
    ids = input_ids.detach().to("cpu")
    not_pad = ids.ne(pad_id)
    n_ids = not_pad.sum(-1)
    T = ids.shape[1]
    ar = torch.arange(T)[None, :]
    if not torch.equal(not_pad, ar < n_ids[:, None]):
        raise NotImplementedError("pad tokens inside a sequence: only right-padded batches (tokenizer(padding='max_length')) are supported")
    if attention_mask is None:
        if int(n_ids.min()) != T:
            raise NotImplementedError("padded input_ids without attention_mask (HF would attend to the pad tokens): pass the tokenizer's mask")
        return [T] * ids.shape[0]
    m = attention_mask.detach().to("cpu").ne(0)
    if not torch.equal(m, not_pad):
        raise NotImplementedError("attention_mask has to mark exactly the non-pad tokens of a right-padded batch")
    lens = [int(v) for v in n_ids.tolist()]
    if min(lens) < 1:
        raise ValueError("empty sequence (HF tokenizers always emit <s> </s>)")
    return lens


class RobertaModel(_Base):
    """RobertaModel replacement (HF modeling_roberta.py): embeddings + post-LN encoder; the pooler is not on this path
    (`pooler_output` is None: the reference reads hidden_states / last_hidden_state only)."""

    @classmethod
    def from_pretrained(cls, name_or_path: str, **kw) -> "RobertaModel":
        """RobertaModel.from_pretrained(SSL_TYPE) (preprocess_roberta.py:104): same resolution rules as AutoModel."""
        from .modeling import AutoModel

        model = AutoModel.from_pretrained(name_or_path, **kw)
        if not isinstance(model, cls):
            raise OSError(f"'{name_or_path}' is not a RoBERTa checkpoint")
        return model

    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, token_type_ids=None,
                position_ids=None, output_hidden_states: Optional[bool] = None, output_attentions: Optional[bool] = None,
                return_dict: Optional[bool] = None, **kw) -> ModelOutput:
        if output_attentions:
            raise NotImplementedError("attention probabilities are never materialised by the flash-style kernel")
        if position_ids is not None:
            raise NotImplementedError("explicit position_ids are not supported (they follow input_ids, as in HF's default)")
        if token_type_ids is not None and bool(torch.as_tensor(token_type_ids).ne(0).any()):
            raise NotImplementedError("token types other than 0 are not supported (RoBERTa has a single type)")
        if input_ids.dim() == 1:
            input_ids = input_ids[None]
        B, T = input_ids.shape
        L = self.cfg.num_hidden_layers
        layers = range(L + 1) if output_hidden_states else [L]
        frames, _, idx, _ = self._encode(input_ids, attention_mask, layers, REDUCE_NONE, None, True, False)
        hs = tuple(frames[i].view(B, T, -1) for i in range(len(idx)))
        return ModelOutput(last_hidden_state=hs[-1], pooler_output=None, hidden_states=hs if output_hidden_states else None)

    __call__ = forward

    def _encode(self, input_ids, attention_mask, layers, reduce, layer_weights, want_frames, want_pooled):
        lens = _valid_lengths(input_ids, attention_mask, self.cfg.pad_token_id)
        lo, hi = int(input_ids.min()), int(input_ids.max())
        if lo < 0 or hi >= self.cfg.vocab_size:
            raise IndexError(f"index out of range in self (token id {lo if lo < 0 else hi}, vocabulary {self.cfg.vocab_size})")
        ids = input_ids.to(self.device, torch.int32).contiguous()
        frames, pooled, idx = self.engine.encode_text(ids, lens, layers=layers, reduce=reduce, want_frames=want_frames,
                                                      want_pooled=want_pooled, layer_weights=layer_weights)
        return frames, pooled, idx, lens

    @torch.no_grad()
    def extract_tokens(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, layer: int = -1, average: bool = False,
                       want_frames: bool = True, want_pooled: bool = False, layer_weights=None, layers=None) -> Extracted:
        """Batched form of preprocess_roberta.py:57-70: hidden_states[layer] (or the mean of the last four) for every
        row of a padded [B, max_len] batch. frames[b] is [max_len, d] - pad positions included, as the reference saves
        them; pooled[b] is the mean over the non-pad tokens."""
        sel, reduce, lw = self._select(layer, average, layer_weights, layers)
        B, T = input_ids.shape
        frames, pooled, _, lens = self._encode(input_ids, attention_mask, sel, reduce, lw, want_frames, want_pooled)
        per, f2, ranges = None, None, None
        if frames is not None:
            f2 = frames if reduce != REDUCE_NONE else frames[0]
            ranges = [(b * T, (b + 1) * T) for b in range(B)]
            per = [f2[a:e] for a, e in ranges]
        if pooled is not None and reduce == REDUCE_NONE:
            pooled = pooled[0]
        return Extracted(per, pooled, lens, f2, ranges)
