"""Method bodies that ``install()`` places on the reference's model classes so that the WHOLE hot path -- not only the
three leaf modules -- runs through this package without editing the reference's source:

  * ``kwclip_compute_loss``   -> ``KWClip_GeneralTransformer.compute_loss`` (avssl/model/kwClip.py:999-1040), reached
                                 from ``training_step_end`` / ``validation_step_end`` (kwClip.py:149-193, :248-285):
                                 L2-normalise + pack + (all-gather) + masked InfoNCE + quantity loss (N0, G0, C0, S3).
  * ``kwclip_forward``        -> ``KWClip_GeneralTransformer.forward`` (kwClip.py:839-960): the same plumbing around the
                                 towers and branches, minus the three ``f / f.norm()`` launches chains (kwClip.py:857,
                                 :905-907, :913-915), which move into the pack kernel of ``kwclip_compute_loss``.
  * ``fused_upstream_forward``-> wraps ``FairseqSpeechEncoder_Hubert.forward`` / ``S3prlSpeechEncoderPlus.forward``
                                 (avssl/module/speech_encoder_plus.py:520-640, :240-311): the per-layer ``method1`` /
                                 ``method2`` rescale loop (:572-592) is skipped and folded into the weighted-sum kernel
                                 (S1').
  * ``clipmodel_init``        -> wraps ``ClipModel.__init__`` (avssl/module/clip_official.py:30-108): the
                                 ``reduce_subword_embbedding`` branch (:63-108) goes through ``reduce_subword_embedding``
                                 (N2).
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Dict

import torch
import torch.distributed as dist

from .. import _lib
from ..module.clip_glue import reduce_subword_embedding
from ..module.weighted_sum import NORM_MODES, WeightedSumLayer
from . import kw_glue

logger = logging.getLogger(__name__)

FEAT_SELECT_IDX_WEIGHTED_SUM_MODE = "weighted_sum"  # avssl/module/speech_encoder_plus.py:26


# ---------------------------------------------------------------------------------------------------------------------
# C0 / N0 / G0: the loss
# ---------------------------------------------------------------------------------------------------------------------
def kwclip_compute_loss(self, inputDict: dict) -> Dict[str, torch.Tensor]:
    """``self`` is the reference's ``KWClip_GeneralTransformer``.  Same keys in and out as kwClip.py:999-1040.

    Features may arrive normalised (the reference's own ``forward``) or not (``kwclip_forward``): the pack kernel
    normalises, which is the identity on unit rows.  Under ``strategy: dp`` the dict already holds the whole batch
    (DataParallel gathered it, kwClip.py:149-169); with one process per GPU (``torch.distributed`` initialised) the
    features are all-gathered here, the loss is the global one, gradients flow into the local rows, and the returned
    ``loss`` carries ``world_size`` times its own gradient so that DDP's gradient averaging yields the reference's
    single-loss gradient (its VALUE is unchanged: the logged numbers are the reference's)."""
    assert isinstance(inputDict, dict)
    required_keys = {"id", "image_feat"}
    assert required_keys.issubset(set(inputDict.keys())), f"required: {required_keys}, input: {inputDict.keys()}"
    settings = self.config.model_settings
    cw = float(getattr(settings, "cascaded_objective_weight", 0.0))
    pw = float(getattr(settings, "parallel_objective_weight", 0.0))
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    feats, rows = kw_glue.gather_loss_feats(inputDict, None)
    out = kw_glue.compute_loss(feats, self.criterion, cascaded_objective_weight=cw, parallel_objective_weight=pw,
                               quantity_loss_weight=float(getattr(self, "quantity_loss_weight", 0.0)),
                               quantity_loss_criteria=getattr(self, "quantity_loss_criteria", None),
                               local_rows=rows if world > 1 else None)
    if world > 1 and torch.is_tensor(out["loss"]):
        loss = out["loss"]
        out["loss"] = loss.detach() + kw_glue.ddp_grad_scale(world) * (loss - loss.detach())
    return out


def kwclip_forward(self, batch: dict):
    """``KWClip_GeneralTransformer.forward`` (kwClip.py:839-960) without its three feature normalisations while training
    (they are fused into the pack kernel of ``kwclip_compute_loss``); evaluation keeps them, because the validation /
    retrieval code consumes the third output directly."""
    wav, wav_len, image, sample_id = batch["wav"], batch["wav_len"], batch["image"], batch["id"]
    self.clip.update_device(self.device)
    audio_feat, audio_feat_len = self.forward_audio(wav, wav_len, return_hidden_states=False)
    image_feat = self.forward_image(image)
    if self.img_enc_proj_net is not None:
        image_feat = self.img_enc_proj_net(image_feat)
    keep_norm = not self.training

    def unit(f):
        return f / f.norm(dim=-1, keepdim=True) if keep_norm else f

    image_feat = unit(image_feat)
    output = None
    if self.cascaded_branch is not None:
        other = None
        if any(c.__name__ == "KW_CascadedBranchPlus" for c in type(self.cascaded_branch).__mro__):
            # the "+" branches down-sample with CIF and need a target length (kwClip.py:861-877)
            other = {"global_step": self.global_step}
            if getattr(self.cascaded_branch, "using_gt_len", False):
                assert "text" in batch, f"Text captions are required, {batch.keys()}"
                other["target_len"] = torch.LongTensor(
                    [(t.squeeze().tolist().index(49407) - 1) for t in batch["text"]]).to(wav.device)
            else:
                other["target_len"] = (audio_feat_len / 20).round().long()
        output = self.cascaded_branch(audio_feat=audio_feat, audio_feat_len=audio_feat_len, otherInputs=other)
    if self.parallel_branch is not None:
        output = self.parallel_branch(audio_feat=audio_feat, audio_feat_len=audio_feat_len)
    parallel_audio_feat, cascaded_audio_feat = output["parallel_audio_feat"], output["cascaded_audio_feat"]
    vq_results, keywords, dsample_results = output["vq_results"], output["keywords"], output["dsample_results"]
    keywords_len = dsample_results["dsample_feats_length"] if dsample_results is not None else None

    losses = {"id": sample_id, "image_feat": image_feat}
    if cascaded_audio_feat is not None:
        if self.c_branch_proj_net is not None:
            cascaded_audio_feat = self.c_branch_proj_net(cascaded_audio_feat)
        cascaded_audio_feat = unit(cascaded_audio_feat)
        losses["cascaded_audio_feat"] = cascaded_audio_feat
    if parallel_audio_feat is not None:
        if self.p_branch_proj_net is not None:
            parallel_audio_feat = self.p_branch_proj_net(parallel_audio_feat)
        parallel_audio_feat = unit(parallel_audio_feat)
        losses["parallel_audio_feat"] = parallel_audio_feat
    if self.cascaded_branch is not None and getattr(self.cascaded_branch, "downsampling_type", None) == "cif":
        assert "target_len" in dsample_results and "quantity_out" in dsample_results, f"{dsample_results.keys()}"
        losses["cif_quantity_out"] = dsample_results["quantity_out"]
        losses["cif_target_len"] = dsample_results["target_len"]

    log_metrics = {"cl_temp": self.criterion.current_temperature}
    if vq_results is not None:
        log_metrics["softmax_temp"] = vq_results["temp"]
    if self.cascaded_branch is not None:
        if dsample_results is not None and "dsample_len_diff" in dsample_results:
            log_metrics["dsample_len_diff"] = dsample_results["dsample_len_diff"]
        log_keys = ["temp", "code_perplexity", "prob_perplexity", "ent_per_t"]
        assert set(log_keys).issubset(set(vq_results.keys())), f"log keys: {log_keys}, result: {vq_results.keys()}"
        log_metrics.update({k: vq_results[k] for k in log_keys})
    others = {"id": sample_id, "image_feat": image_feat, "parallel_audio_feat": parallel_audio_feat,
              "cascaded_audio_feat": cascaded_audio_feat, "vq_results": vq_results, "keywords": keywords,
              "dsample_results": dsample_results, "keywords_len": keywords_len}
    return losses, log_metrics, others


# ---------------------------------------------------------------------------------------------------------------------
# S1': the caller tail of the upstream wrappers
# ---------------------------------------------------------------------------------------------------------------------
def fused_upstream_forward(original_forward: Callable) -> Callable:
    """Wrap a reference speech-encoder ``forward``: when it would rescale every hidden state in Python (``method1`` /
    ``method2``, speech_encoder_plus.py:572-592) and then take the weighted sum, skip the loop and let the weighted-sum
    kernel apply the rescale per frame / per utterance while it reads the layers (SCP_NORM_L2_FRAME / SCP_NORM_UTT_MEAN).
    Anything else (other ``feat_select_idx`` values, ``return_hidden_states`` -- which hands the rescaled states to the
    caller --, the ``s3prl`` LayerNorm type, a foreign weighted-sum layer) runs the original code unchanged."""

    def forward(self, wav, wav_len=[], feat_select_idx=None, return_hidden_states: bool = False):  # noqa: B006
        layer = getattr(self, "weightedsum_layer", None)
        select = self.feat_select_idx if feat_select_idx is None else feat_select_idx
        fuse = (getattr(self, "normalize_hiddenstates", False) and str(getattr(self, "normalize_type", "")).startswith("method")
                and select == FEAT_SELECT_IDX_WEIGHTED_SUM_MODE and not return_hidden_states
                and isinstance(layer, WeightedSumLayer))
        if not fuse:
            return original_forward(self, wav, wav_len, feat_select_idx, return_hidden_states)
        self.normalize_hiddenstates = False
        layer.upstream_norm_mode = NORM_MODES[self.normalize_type]
        try:
            return original_forward(self, wav, wav_len, feat_select_idx, return_hidden_states)
        finally:
            layer.upstream_norm_mode = None
            self.normalize_hiddenstates = True

    forward.__wrapped__ = original_forward
    return forward


# ---------------------------------------------------------------------------------------------------------------------
# N2: vocabulary reduction inside ClipModel.__init__
# ---------------------------------------------------------------------------------------------------------------------
def apply_reduced_vocab(clip_model, usage_path: str) -> None:
    """What the ``reduce_subword_embbedding`` branch of ``ClipModel.__init__`` leaves behind (clip_official.py:63-108),
    built by ``reduce_subword_embedding``: same attribute names, same reduced ``nn.Embedding``."""
    if not os.path.exists(usage_path):
        logger.error(f"File not found {usage_path}")
        raise SystemExit(1)  # the reference calls exit(1) here (:65-67)
    enc = clip_model.tokenizer.encoder
    reduced, vocab, original = reduce_subword_embedding(
        clip_model.model.token_embedding, usage_path, enc["<|startoftext|>"], enc["<|endoftext|>"],
        trainable=bool(clip_model.text_encoder_trainable))
    logger.warning("Reduce text embedding to size of {}".format(len(vocab)))
    clip_model.selected_text_emb_ids = vocab.selected_text_emb_ids
    clip_model.selected_text_emb_ids_dist = vocab.selected_text_emb_ids_dist
    clip_model.original_text_emb_weight = original
    clip_model.model.token_embedding = reduced
    clip_model.original2Reduced = vocab.original2Reduced
    clip_model.reducedl2Original = vocab.reducedl2Original
    clip_model.startOfTxt_reduced = vocab.startOfTxt_reduced
    clip_model.endOfTxt_reduced = vocab.endOfTxt_reduced


def clipmodel_init(original_init: Callable) -> Callable:
    def __init__(self, name, device="cpu", image_encoder_trainable=False, text_encoder_trainable=False,
                 reduce_subword_embbedding=None, **kwargs):
        original_init(self, name, device, image_encoder_trainable, text_encoder_trainable, None, **kwargs)
        if reduce_subword_embbedding is not None:
            apply_reduced_vocab(self, reduce_subword_embbedding)

    __init__.__wrapped__ = original_init
    return __init__
