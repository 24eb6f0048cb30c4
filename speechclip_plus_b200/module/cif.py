"""Drop-in for ``avssl.module.cif.CIF`` (reference: avssl/module/cif.py:24-311) -- the continuous integrate-and-fire
down-sampler that produces the dynamic-length keyword sequence of the "+" branches.

Same constructor, sub-module names and ``state_dict`` keys (``conv.*`` / ``dense_proj.*``, ``weight_proj.*``,
``cif_output_proj.*``), same ``forward(input_dict, target_lengths) -> dict`` keys.  The weight generator (Conv1d / Linear +
sigmoid) is a handful of library layers and stays as in the reference; ``integrate_and_fire`` -- cumsum, 2 + N
``scatter_add_`` passes over (B,S,C) temporaries, an ``.item()``-driven Python loop and the inference tail handling -- runs
in csrc/scp_cif.cu (one sequential-cumsum kernel, one single-pass integrate kernel, one tail kernel; backward included).
"""
from __future__ import annotations

import logging
from typing import Optional

import torch
from torch import nn

from .. import _lib
from .clip_glue import get_keypadding_mask

logger = logging.getLogger(__name__)

MAX_FEAT_LEN = 75  # cif.py:11

__all__ = ["CIF", "integrate_and_fire", "MAX_FEAT_LEN"]


class _CifFn(torch.autograd.Function):
    """(out[:, :T_out], feat_len, fire_mask) = integrate_and_fire(input, alpha)."""

    @staticmethod
    def forward(ctx, x, alpha, threshold, tail_mode, firing_threshold):
        # tail_mode: 0 = slice off the tail row (training, or apply_tail_handling False), 1 = inference tail handling
        lib = _lib.load()
        _lib.require_cuda(x, "CIF.integrate_and_fire")
        B, S, C = x.shape
        dev = x.device
        xc = x.detach()
        if xc.dtype != torch.float32 or not xc.is_contiguous():
            xc = xc.float().contiguous()
        al = alpha.detach().float().contiguous()
        csum = torch.empty((B, S), dtype=torch.float32, device=dev)
        feat_len = torch.empty(B, dtype=torch.int64, device=dev)
        stream = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(lib.scp_cif_plan(_lib.ptr(al), B, S, float(threshold), MAX_FEAT_LEN, _lib.ptr(csum),
                                        _lib.ptr(feat_len), stream), "scp_cif_plan")
            T = int(feat_len.max())  # the output shape depends on it (the reference synchronises here too, cif.py:189)
            out = torch.empty((B, T + 1, C), dtype=torch.float32, device=dev)
            fire_mask = torch.empty((B, S), dtype=torch.uint8, device=dev)
            tail_w = torch.empty(B, dtype=torch.float32, device=dev) if tail_mode == 1 else None
            _lib.check(lib.scp_cif_fire_fwd(_lib.ptr(xc), _lib.ptr(al), _lib.ptr(csum), _lib.ptr(feat_len), B, S, C,
                                            float(threshold), T, _lib.ptr(out), _lib.ptr(fire_mask), _lib.ptr(tail_w),
                                            stream), "scp_cif_fire_fwd")
            keep_len = scale_row = None
            T_out = T
            if tail_mode == 1:
                feat_len_new = torch.empty_like(feat_len)
                _lib.check(lib.scp_cif_tail(_lib.ptr(out), B, T + 1, C, _lib.ptr(feat_len), _lib.ptr(tail_w),
                                            float(threshold), float(firing_threshold), MAX_FEAT_LEN,
                                            _lib.ptr(feat_len_new), stream), "scp_cif_tail")
                T_out = int(feat_len_new.max())
                keep_len, scale_row = feat_len_new, feat_len
                # the reference marks the extra fire in `fired_marks` with a (B,B)-broadcast column update (cif.py:281-283)
                # -- `fire_mask[:, feat_lengths - 1] = fire_mask[:, feat_lengths - 1] + extend_mask` ORs extend[k] into
                # column feat_lengths[k]-1 of EVERY row; with duplicate columns the last k wins (sequential CPU
                # semantics; CUDA index_put_ with duplicates is unordered, so it is spelt out deterministically here)
                extend = tail_w >= firing_threshold
                cols = feat_len + extend.long() - 1                       # before the clip to MAX_FEAT_LEN (:280-284)
                last = torch.full((S,), -1, dtype=torch.long, device=dev)
                last = last.scatter_reduce(0, cols.clamp(0, S - 1), torch.arange(B, device=dev), "amax")
                col_or = (last >= 0) & extend[last.clamp(min=0)]         # all-False when nothing extends (:271)
                fire_mask = (fire_mask.bool() | col_or[None, :]).to(torch.uint8)
        ctx.save_for_backward(xc, al, csum)
        ctx.aux = (keep_len, scale_row, tail_w)
        ctx.cfg = (float(threshold), float(firing_threshold), T, T_out)
        ctx.in_dtypes = (x.dtype, alpha.dtype)
        out_len = keep_len if keep_len is not None else feat_len
        ctx.mark_non_differentiable(out_len, fire_mask)
        return out[:, :T_out], out_len, fire_mask

    @staticmethod
    def backward(ctx, g_out, _g_len, _g_mask):
        lib = _lib.load()
        xc, al, csum = ctx.saved_tensors
        keep_len, scale_row, tail_w = ctx.aux
        threshold, firing_threshold, T, T_out = ctx.cfg
        B, S, C = xc.shape
        dev = xc.device
        g = g_out.float().contiguous()
        g_x = torch.empty_like(xc)
        g_alpha = torch.empty_like(al)
        ws = torch.empty(2 * B * S, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.scp_cif_fire_bwd(_lib.ptr(g), T_out, _lib.ptr(xc), _lib.ptr(al), _lib.ptr(csum), B, S, C,
                                            threshold, T, _lib.ptr(keep_len), _lib.ptr(scale_row), _lib.ptr(tail_w),
                                            firing_threshold, _lib.ptr(g_x), _lib.ptr(g_alpha), _lib.ptr(ws),
                                            ws.numel() * 4, _lib.stream_ptr(dev)), "scp_cif_fire_bwd")
        return g_x.to(ctx.in_dtypes[0]), g_alpha.to(ctx.in_dtypes[1]), None, None, None


def integrate_and_fire(input: torch.Tensor, alpha: torch.Tensor, cif_threshold: float = 1.0,
                       target_lengths: Optional[torch.Tensor] = None, apply_tail_handling: bool = True,
                       tail_handling_firing_threshold: float = 0.5) -> dict:
    """``CIF.integrate_and_fire`` (cif.py:157-311) as a function: same result dict."""
    B, S, C = input.size()
    assert tuple(alpha.size()) == (B, S), f"{alpha.size()} != {(B, S)}"
    tail_mode = 1 if (apply_tail_handling and target_lengths is None) else 0
    output, feat_lengths, fire_mask = _CifFn.apply(input, alpha, cif_threshold, tail_mode,
                                                   tail_handling_firing_threshold)
    return {
        "dsample_feats_pad_mask": get_keypadding_mask(output.shape[1], feat_lengths),
        "dsample_feats": output,
        "dsample_feats_length": feat_lengths,
        "alpha": alpha,
        "fired_marks": fire_mask.bool(),
    }


class CIF(nn.Module):
    def __init__(self, cif_threshold=1.0, cif_output_dim=768, encoder_embed_dim=768, produce_weight_type="conv",
                 num_layer=1, conv_cif_width=3, conv_cif_dropout=0.1, apply_scaling=True, apply_tail_handling=True,
                 tail_handling_firing_threshold=0.5, scaling_step=-1, **config):
        super().__init__()
        # Load configurations (cif.py:42-53)
        self.cif_threshold = cif_threshold
        self.cif_output_dim = cif_output_dim
        self.encoder_embed_dim = encoder_embed_dim
        self.produce_weight_type = produce_weight_type
        self.conv_cif_width = conv_cif_width
        self.conv_cif_dropout = conv_cif_dropout
        self.apply_scaling = apply_scaling
        self.apply_tail_handling = apply_tail_handling
        self.tail_handling_firing_threshold = tail_handling_firing_threshold
        self.scaling_step = scaling_step
        self.num_layer = num_layer
        if self.apply_scaling:
            logger.info(f"Apply scaling strategy step: {self.scaling_step}")
        # weight generator (cif.py:58-87): library layers, unchanged
        if self.produce_weight_type == "dense":
            self.dense_proj = nn.Sequential(nn.Linear(self.encoder_embed_dim, self.encoder_embed_dim), nn.ReLU())
        elif self.produce_weight_type == "conv":
            conv_list = []
            for _ in range(self.num_layer):
                conv_list += [
                    nn.Conv1d(self.encoder_embed_dim, self.encoder_embed_dim, self.conv_cif_width, stride=1,
                              padding=int(self.conv_cif_width / 2), dilation=1, groups=1, padding_mode="zeros"),
                    nn.Dropout(),
                    nn.ReLU(),
                ]
            self.conv = nn.Sequential(*conv_list)
        else:
            raise NotImplementedError(self.produce_weight_type)
        self.weight_proj = nn.Sequential(nn.Dropout(), nn.Linear(self.encoder_embed_dim, 1), nn.Sigmoid())
        if self.cif_output_dim != self.encoder_embed_dim:
            logger.info(f"Built projection layer to match the dimension of input {self.encoder_embed_dim} and output "
                        f"{self.cif_output_dim}")
            self.cif_output_proj = nn.Linear(self.encoder_embed_dim, self.cif_output_dim, bias=False)

    def forward(self, input_dict, target_lengths=None, eps=1e-5):
        input_feats = input_dict["audio_feat"]  # B x T x D
        input_feats_pad_mask = input_dict["audio_feat_pad_mask"].bool()  # B x T
        original_length = (~input_feats_pad_mask).sum(-1).long()  # B
        if self.scaling_step >= 0:
            if self.apply_scaling and input_dict["global_step"] >= self.scaling_step:
                self.apply_scaling = False
        # Produce weights for integration (cif.py:106-129)
        if self.produce_weight_type == "dense":
            proj_out = self.dense_proj(input_feats)
            alpha = self.weight_proj(proj_out)  # B x T x 1  (the reference continues with this shape and fails, :108-109)
            raise NotImplementedError("produce_weight_type='dense' is broken in the reference (alpha keeps its last "
                                      "dim and is never masked, cif.py:106-109); no shipped recipe uses it")
        conv_input = input_feats.permute(0, 2, 1)
        proj_input = self.conv(conv_input).permute(0, 2, 1)
        logits = self.conv_dropout(proj_input) if hasattr(self, "conv_dropout") else proj_input
        alpha = self.weight_proj(logits).clip(min=0.0, max=1.0).float().squeeze(-1)  # B x T
        alpha = alpha.masked_fill(input_feats_pad_mask, 0.0)  # (the reference assigns in place: alpha[mask] = 0.0)
        orig_alpha = alpha
        alpha_sum = alpha.sum(1)
        assert (alpha_sum > 0).any(), f"alphas are all zero:\n{alpha_sum}"  # cif.py:124
        if self.apply_scaling and target_lengths is not None:
            desired_sum = self.cif_threshold * target_lengths.type_as(alpha) + eps
            alpha = alpha * (desired_sum / alpha_sum).unsqueeze(1)
        result_dict = {
            "quantity_out": alpha_sum,
            "orig_alpha": orig_alpha,
            "original_length": original_length,
            "target_len": target_lengths,
        }
        dsmaple_dict = self.integrate_and_fire(input_feats, alpha, target_lengths=target_lengths)
        dsmaple_dict["input_feats_pad_mask"] = input_feats_pad_mask
        result_dict = {**result_dict, **dsmaple_dict}
        if self.cif_output_dim != self.encoder_embed_dim:
            result_dict["dsample_feats"] = self.cif_output_proj(result_dict["dsample_feats"])
            result_dict["dsample_feats"] = result_dict["dsample_feats"] * result_dict["dsample_feats_pad_mask"]
        return result_dict

    def integrate_and_fire(self, input: torch.Tensor, alpha: torch.Tensor,
                           target_lengths: Optional[torch.Tensor] = None) -> dict:
        return integrate_and_fire(input, alpha, self.cif_threshold, target_lengths, self.apply_tail_handling,
                                  self.tail_handling_firing_threshold)
