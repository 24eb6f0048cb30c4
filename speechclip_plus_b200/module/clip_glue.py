"""N3 -- the step right after the vector quantiser on the cascaded path (reference: ``ClipModel.encode_keywords``,
avssl/module/clip_official.py:222-279, and ``get_keypadding_mask``, avssl/util/data_utils.py:6-22).

The reference assembles the CLIP text-transformer input with an id tensor + embedding lookup, a Python loop over the
batch that slice-assigns each utterance's keywords (one small kernel per sample, clip_official.py:261-265) and the
positional-embedding add.  ``splice_keywords`` produces the same ``(B,77,D)`` tensor and the EOT gather index in one
CUDA pass (csrc/scp_splice.cu); ``encode_keywords`` is the drop-in method body (the frozen CLIP transformer,
``ln_final`` and ``text_projection`` that follow are the reference's own modules, called unchanged).
"""
from __future__ import annotations

from typing import Tuple, Union

import torch

from .. import _lib

__all__ = ["splice_keywords", "encode_keywords", "get_keypadding_mask", "ReducedVocab", "reduce_subword_embedding"]


class ReducedVocab:
    """N2 -- the vocabulary-reduction state ``ClipModel.__init__`` builds from a ``text_clip_vocab_usage_byfreq.npy``
    file (avssl/module/clip_official.py:63-108; file format: (V',2) int64 rows ``[original_token_id, count]`` sorted by
    frequency, avssl/data/{flickr,coco}_stat/).  Attribute names follow the reference so that code written against
    ``ClipModel`` (``selected_text_emb_ids``, ``original2Reduced``, ``startOfTxt_reduced`` ...) reads the same."""

    def __init__(self, usage, sot_token: int, eot_token: int):
        import numpy as np
        _data = np.load(usage) if isinstance(usage, (str, bytes)) or hasattr(usage, "__fspath__") else np.asarray(usage)
        if _data.ndim != 2 or _data.shape[1] != 2:
            raise ValueError(f"vocabulary usage table must be (V',2) [token id, count], got {_data.shape}")
        self.selected_text_emb_ids = _data[:, 0]                                           # :71
        dist = _data[:, 1]
        self.selected_text_emb_ids_dist = torch.from_numpy(dist / np.sum(dist))           # :72-76
        self.original2Reduced = {int(old): new for new, old in enumerate(self.selected_text_emb_ids)}   # :94-97
        self.reducedl2Original = {new: int(old) for new, old in enumerate(self.selected_text_emb_ids)}  # :98-101
        self.startOfTxt_reduced = self.original2Reduced[int(sot_token)]                   # :103-105 (KeyError if absent)
        self.endOfTxt_reduced = self.original2Reduced[int(eot_token)]                     # :107-109

    def __len__(self) -> int:
        return len(self.selected_text_emb_ids)


def reduce_subword_embedding(token_embedding: torch.nn.Embedding, usage, sot_token: int, eot_token: int,
                             trainable: bool = False):
    """Replacement for the ``reduce_subword_embbedding`` branch of ``ClipModel.__init__`` (clip_official.py:63-108):
    returns ``(reduced nn.Embedding, ReducedVocab, original weight)``.  The reduced table is the frozen (V',D) matrix the
    VQ runs against; its fp16 unit-norm copy, transpose, norms and mean are built lazily -- once per table version -- by
    ``TokenTableCache`` (scp_vq_prepare_table), which is where the reference recomputes ``||e_v||`` K times per step.
    With the by-frequency files row 0 is the pad token ``!``, rows 2 and 3 are SOT / EOT: the ``prob_msk=[0,2,3]`` default
    of the quantiser (my_vector_quantizer.py:64)."""
    vocab = ReducedVocab(usage, sot_token, eot_token)
    original = token_embedding.weight
    ids = torch.as_tensor(vocab.selected_text_emb_ids, dtype=torch.long, device=original.device)
    reduced = torch.nn.Embedding.from_pretrained(original.detach()[ids], freeze=not trainable)      # :84-92
    return reduced, vocab, original


class _SpliceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, keywords, table, pos_emb, kw_num, fixed_num, sot, eot):
        lib = _lib.load()
        _lib.require_cuda(keywords, "splice_keywords")
        B, Kmax, D = keywords.shape
        L = pos_emb.shape[0]
        dev = keywords.device
        kw = keywords.detach()
        if kw.dtype != torch.float32 or not kw.is_contiguous():
            kw = kw.float().contiguous()
        tab = table.detach()
        pos = pos_emb.detach().to(tab.dtype).contiguous()
        if not tab.is_contiguous():
            tab = tab.contiguous()
        x = torch.empty((B, L, D), dtype=tab.dtype, device=dev)
        eot_index = torch.empty(B, dtype=torch.int64, device=dev)
        num = None
        if kw_num is not None:
            num = kw_num.to(device=dev, dtype=torch.int64).contiguous()
        with torch.cuda.device(dev):
            st = lib.scp_kw_splice_fwd(_lib.ptr(kw), _lib.ptr(num), int(fixed_num), _lib.ptr(tab), _lib.ptr(pos),
                                       _lib.dtype_code(tab.dtype), B, Kmax, D, L, int(sot), int(eot), _lib.ptr(x),
                                       _lib.ptr(eot_index), _lib.stream_ptr(dev))
        _lib.check(st, "scp_kw_splice_fwd")
        ctx.num = num
        ctx.fixed_num = int(fixed_num)
        ctx.shape = (B, Kmax, D, L)
        ctx.in_dtype = keywords.dtype
        ctx.mark_non_differentiable(eot_index)
        return x, eot_index

    @staticmethod
    def backward(ctx, g_x, _g_idx):
        lib = _lib.load()
        B, Kmax, D, L = ctx.shape
        g = g_x.contiguous()
        g_kw = torch.empty((B, Kmax, D), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            st = lib.scp_kw_splice_bwd(_lib.ptr(g), _lib.dtype_code(g.dtype), _lib.ptr(ctx.num), ctx.fixed_num, B, Kmax,
                                       D, L, _lib.ptr(g_kw), _lib.stream_ptr(g.device))
        _lib.check(st, "scp_kw_splice_bwd")
        return g_kw.to(ctx.in_dtype), None, None, None, None, None, None


def splice_keywords(keywords: torch.Tensor, keyword_num: Union[int, torch.Tensor], token_table: torch.Tensor,
                    positional_embedding: torch.Tensor, sot_token: int, eot_token: int
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(B,Kmax,D) keywords -> ((B,L,D) text-transformer input incl. positional embedding, (B,) EOT positions).

    ``keyword_num`` is the reference's argument: an int (every utterance has that many keywords and
    ``keywords.shape[1]`` must equal it, clip_official.py:266-267) or a (B,) tensor of per-utterance counts (:252-265).
    """
    if token_table.requires_grad or positional_embedding.requires_grad:
        raise _lib.ScpError("splice_keywords: the token table and the positional embedding must be frozen "
                            "(text_encoder_trainable=False, clip_official.py:110-123)")
    if isinstance(keyword_num, torch.Tensor):
        return _SpliceFn.apply(keywords, token_table, positional_embedding, keyword_num, 0, sot_token, eot_token)
    if keywords.shape[1] != int(keyword_num):
        raise RuntimeError(f"shape mismatch: {keywords.shape[1]} keywords per utterance but keyword_num={keyword_num}")
    return _SpliceFn.apply(keywords, token_table, positional_embedding, None, int(keyword_num), sot_token, eot_token)


def encode_keywords(self, keywords: torch.Tensor, keyword_num: Union[int, torch.Tensor]) -> torch.Tensor:
    """Drop-in body of ``ClipModel.encode_keywords`` (clip_official.py:222-279); ``self`` is the reference ClipModel."""
    if not isinstance(keywords, torch.Tensor):
        raise TypeError(f"Unknown keywords type {type(keywords)}")
    if self.selected_text_emb_ids is None:
        sot_token, eot_token = self.tokenizer.encoder["<|startoftext|>"], self.tokenizer.encoder["<|endoftext|>"]
    else:
        sot_token, eot_token = self.startOfTxt_reduced, self.endOfTxt_reduced
    x, index = splice_keywords(keywords, keyword_num, self.model.token_embedding.weight,
                               self.model.positional_embedding, sot_token, eot_token)
    x = x.permute(1, 0, 2)  # NLD -> LND
    x = self.model.transformer(x)
    x = x.permute(1, 0, 2)  # LND -> NLD
    x = self.model.ln_final(x)
    # take features from the eot embedding
    return x[torch.arange(x.shape[0], device=x.device), index] @ self.model.text_projection


def get_keypadding_mask(max_length: int, data_lens: torch.Tensor) -> torch.Tensor:
    """bool (B, max_length), True marks padding (data_utils.py:6-22) -- built on the device of ``data_lens``."""
    lib = _lib.load()
    _lib.require_cuda(data_lens, "get_keypadding_mask")
    lens = data_lens.to(torch.int64).contiguous()
    B = lens.shape[0]
    mask = torch.empty((B, max_length), dtype=torch.uint8, device=lens.device)
    with torch.cuda.device(lens.device):
        st = lib.scp_keypadding_mask(_lib.ptr(lens), B, int(max_length), _lib.ptr(mask), _lib.stream_ptr(lens.device))
    _lib.check(st, "scp_keypadding_mask")
    return mask.view(torch.bool)
