"""S1' -- the caller tail of the reference's HuBERT wrappers, fused (reference:
``FairseqSpeechEncoder_Hubert.forward`` avssl/module/speech_encoder_plus.py:572-622 and
``S3prlSpeechEncoderPlus.forward`` :292-311).

After the (frozen, out-of-scope) upstream model has produced its ``layer_results`` the reference
  1. optionally rescales every layer in a Python loop (``normalize_hiddenstates`` with ``normalize_type`` "method1" /
     "method2", :572-592; "s3prl" is instead the LayerNorm flag of the ``WeightedSumLayer``, :472-476),
  2. computes ``feat_len = clamp_max(round(len / downsample_rate), T)`` on the host (:600-611),
  3. calls ``self.weightedsum_layer(hidden_states)`` (:619-622).
``fuse_upstream_features`` does 1 + 3 in one pass over the layer tensors (csrc/scp_wsum.cu, norm modes
SCP_NORM_L2_FRAME / SCP_NORM_UTT_MEAN) and 2 with the reference's exact host arithmetic.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .. import _lib
from .weighted_sum import NORM_MODES, WeightedSumLayer, _WeightedSumFn

__all__ = ["upstream_feat_len", "fuse_upstream_features"]


def upstream_feat_len(wav_len: Sequence[int], downsample_rate: int, max_frames: int,
                      device: Optional[torch.device] = None) -> torch.Tensor:
    """``clamp_max(LongTensor([round(l / rate)]), T)`` (speech_encoder_plus.py:604-611 / :292-296).  ``round`` is
    Python's (half to even), exactly as in the reference."""
    feat_len = torch.LongTensor([round(int(l) / downsample_rate) for l in wav_len])
    if device is not None:
        feat_len = feat_len.to(device)
    return torch.clamp_max(feat_len, max_frames)


def fuse_upstream_features(layer_results: Sequence[torch.Tensor], weightedsum_layer: WeightedSumLayer,
                           normalize_hiddenstates: bool = False, normalize_type: str = "s3prl",
                           wav_len: Optional[Sequence[int]] = None, downsample_rate: int = 320
                           ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Fused replacement of speech_encoder_plus.py:572-622 for ``feat_select_idx == "weighted_sum"``.

    ``layer_results``: the L upstream hidden states, each ``(B,T,D)`` (the ``(T,B,D)``-storage transposed views are
    consumed in place).  Returns ``(features (B,T,D) fp32, feat_len (B,) int64 | None)``.
    """
    assert normalize_type in ("s3prl", "method1", "method2"), normalize_type  # :377
    assert len(layer_results) == weightedsum_layer.n_weights, len(layer_results)  # weighted_sum.py:36
    if normalize_hiddenstates and normalize_type.startswith("method"):
        # the reference builds the layer with normalize_features = False in this case (:472-476) and rescales here
        if weightedsum_layer.normalize_features:
            raise _lib.ScpError("normalize_type=method1/method2 excludes the LayerNorm flag of the WeightedSumLayer "
                                "(speech_encoder_plus.py:472-476)")
        mode = NORM_MODES[normalize_type]
    elif weightedsum_layer.normalize_features:
        mode = NORM_MODES[weightedsum_layer.normalize_type]
    else:
        mode = _lib.SCP_NORM_NONE
    feats = _WeightedSumFn.apply(weightedsum_layer.weights, mode, *layer_results)
    feat_len = None
    if wav_len is not None:
        feat_len = upstream_feat_len(wav_len, downsample_rate, layer_results[0].shape[1], layer_results[0].device)
    return feats, feat_len
