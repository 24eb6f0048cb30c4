"""Drop-in for ``avssl.module.weighted_sum.WeightedSumLayer`` (reference: avssl/module/weighted_sum.py:10-45).

Same constructor, parameter name (``weights``, zeros-initialised, shape ``(n_weights,)``), ``state_dict`` key and
``forward(x: List[Tensor]) -> Tensor`` contract; the arithmetic runs in the sm_100a kernels of csrc/scp_wsum.cu
(one pass over the L layer tensors, no ``torch.stack`` copy, no broadcast temporaries).

Extension (keyword-only, default = reference behaviour): ``normalize_type`` selects which per-layer normalisation is
fused into the sum -- ``"s3prl"`` (LayerNorm, weighted_sum.py:41-42) or the two rescales the reference's HuBERT wrapper
applies in a Python loop before calling the layer (speech_encoder_plus.py:572-592): ``"method1"`` (per-frame L2) and
``"method2"`` (per-utterance mean frame norm).  See ``speech_encoder_plus.fuse_upstream_features``.
"""
from __future__ import annotations

import ctypes
import logging
from typing import List, Sequence

import torch
from torch import nn

from .. import _lib

logger = logging.getLogger(__name__)

LN_EPS = 1e-5  # F.layer_norm default, weighted_sum.py:42
NORM_MODES = {None: _lib.SCP_NORM_NONE, "s3prl": _lib.SCP_NORM_LAYERNORM, "method1": _lib.SCP_NORM_L2_FRAME,
              "method2": _lib.SCP_NORM_UTT_MEAN}


def _uniform_views(layers: Sequence[torch.Tensor]):
    """The kernel reads every layer as x[b*sb + t*st + d].  Layers that already share such a layout (the HuBERT
    wrapper hands over (T,B,D) storage viewed as (B,T,D)) are used in place; anything else is made contiguous."""
    x0 = layers[0]
    ne = 4 if x0.dtype == torch.float32 else 8
    ok = all(
        l.dim() == 3 and l.shape == x0.shape and l.dtype == x0.dtype and l.device == x0.device and l.stride(2) == 1
        and l.stride() == x0.stride() and l.stride(0) % ne == 0 and l.stride(1) % ne == 0 and l.data_ptr() % 16 == 0
        for l in layers)
    if ok:
        return list(layers)
    return [l.contiguous() for l in layers]


class _WeightedSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights: torch.Tensor, norm_mode: int, *layers: torch.Tensor):
        lib = _lib.load()
        x0 = layers[0]
        _lib.require_cuda(x0, "WeightedSumLayer")
        lead_shape = x0.shape[:-1]
        D = x0.shape[-1]
        in_shapes = [l.shape for l in layers]  # the callers' shapes: layer gradients are returned in them
        if x0.dim() != 3:  # (..., D) -> (1, R, D)
            layers = tuple(l.reshape(1, -1, D) for l in layers)
        views = _uniform_views(layers)
        v0 = views[0]
        B, T, _ = v0.shape
        w = weights.detach().float().contiguous()
        y = torch.empty((B, T, D), dtype=torch.float32, device=v0.device)
        ptrs = _lib.ptr_array(views)
        utt_scale = None
        with torch.cuda.device(v0.device):
            if norm_mode == _lib.SCP_NORM_UTT_MEAN:  # statistics pre-pass: 1 / mean_t ||x_{l,b,t}||
                utt_scale = torch.empty((len(views), B), dtype=torch.float32, device=v0.device)
                st = lib.scp_wsum_utt_scale(ptrs, len(views), B, T, D, v0.stride(0), v0.stride(1),
                                            _lib.dtype_code(v0.dtype), _lib.ptr(utt_scale), _lib.stream_ptr(v0.device))
                _lib.check(st, "scp_wsum_utt_scale")
            st = lib.scp_wsum_fwd(ptrs, len(views), B, T, D, v0.stride(0), v0.stride(1), _lib.dtype_code(v0.dtype),
                                  _lib.ptr(w), int(norm_mode), LN_EPS, _lib.ptr(utt_scale), _lib.ptr(y), _lib.SCP_F32,
                                  _lib.stream_ptr(v0.device))
        _lib.check(st, "scp_wsum_fwd")
        ctx.norm_mode = int(norm_mode)
        ctx.utt_scale = utt_scale
        ctx.save_for_backward(w, *views)
        ctx.lead_shape = lead_shape
        ctx.in_shapes = in_shapes
        return y.reshape(*lead_shape, D)

    @staticmethod
    def backward(ctx, grad_y: torch.Tensor):
        lib = _lib.load()
        w, *views = ctx.saved_tensors
        v0 = views[0]
        B, T, D = v0.shape
        L = len(views)
        g = grad_y.reshape(B, T, D).float().contiguous()
        need_layers = any(ctx.needs_input_grad[2:])
        d_w = torch.empty(L, dtype=torch.float32, device=v0.device)
        g_layers = None
        g_ptrs = None
        if need_layers:
            g_layers = [torch.empty((B, T, D), dtype=torch.float32, device=v0.device) for _ in range(L)]
            g_ptrs = _lib.ptr_array(g_layers)
        ws_bytes = lib.scp_wsum_bwd_workspace_bytes(L, B, T, D)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=v0.device)
        with torch.cuda.device(v0.device):
            st = lib.scp_wsum_bwd(_lib.ptr_array(views), L, B, T, D, v0.stride(0), v0.stride(1),
                                  _lib.dtype_code(v0.dtype), _lib.ptr(w), ctx.norm_mode, LN_EPS,
                                  _lib.ptr(ctx.utt_scale), _lib.ptr(g), _lib.SCP_F32, _lib.ptr(d_w),
                                  g_ptrs if g_ptrs is not None else ctypes.cast(None, ctypes.POINTER(ctypes.c_void_p)),
                                  _lib.ptr(ws), ws_bytes, _lib.stream_ptr(v0.device))
        _lib.check(st, "scp_wsum_bwd")
        grads = [None] * L
        if need_layers:
            for i in range(L):
                if ctx.needs_input_grad[2 + i]:
                    grads[i] = g_layers[i].reshape(ctx.in_shapes[i]).to(views[i].dtype)
        return (d_w, None, *grads)


class WeightedSumLayer(nn.Module):
    def __init__(self, n_weights: int, normalize_features: bool = False, *, normalize_type: str = "s3prl"):
        """Softmax-weighted sum of ``n_weights`` hidden representations (weighted_sum.py:11-24)."""
        super().__init__()
        if normalize_type not in ("s3prl", "method1", "method2"):  # speech_encoder_plus.py:377
            raise AssertionError(normalize_type)
        self.normalize_type = normalize_type
        if n_weights > _lib.SCP_MAX_LAYERS:
            raise _lib.ScpError(f"n_weights={n_weights} > {_lib.SCP_MAX_LAYERS}")
        self.n_weights = n_weights
        self.weights = nn.Parameter(torch.zeros((n_weights,), dtype=torch.float))
        self.normalize_features = normalize_features
        if self.normalize_features:
            logger.info("Normalize feature before weighted sum")
        # set (transiently) by the wrapped upstream forward of install(): the caller's per-layer method1 / method2
        # rescale (speech_encoder_plus.py:572-592) is applied by this layer's kernel instead of a Python loop
        self.upstream_norm_mode = None

    def forward(self, x: List[torch.Tensor]) -> torch.Tensor:
        assert len(x) == self.n_weights, len(x)  # weighted_sum.py:36
        mode = NORM_MODES[self.normalize_type] if self.normalize_features else _lib.SCP_NORM_NONE
        forced = getattr(self, "upstream_norm_mode", None)
        if forced is not None:
            if self.normalize_features:  # the reference builds the layer without the LayerNorm flag in this case (:472-476)
                raise _lib.ScpError("normalize_type=method1/method2 excludes the LayerNorm flag of the WeightedSumLayer")
            mode = forced
        return _WeightedSumFn.apply(self.weights, mode, *x)
