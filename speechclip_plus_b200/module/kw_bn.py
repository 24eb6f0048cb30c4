"""Drop-in for ``avssl.module.speechclip_c_modules.kw_bn`` (reference: kw_bn.py:8-228): the keyword batch-norm layers
that sit between the keyword projection and the vector quantiser (``GeneralBranch.project_feats_to_CLIPspace``,
avssl/model/kw_branches.py:143-156).

Same constructors, sub-module names and therefore ``state_dict`` keys (``bn_layer.weight / bias / running_mean /
running_var / num_batches_tracked`` or ``bn_layers.<i>.*``): the ``nn.BatchNorm1d`` objects are kept as PARAMETER
HOLDERS only -- their forward is never called; the arithmetic runs in csrc/scp_kwbn.cu (scp_kwbn_fwd / scp_kwbn_bwd),
which addresses the parameters in place whatever their layout (``BatchNorm1d(D)``, the ``parallel`` variant's
``BatchNorm1d(D*K)`` with feature index ``d*K + k``, or K stacked layers).
"""
from __future__ import annotations

import logging
from typing import Optional

import torch
from torch import nn

from .. import _lib

logger = logging.getLogger(__name__)

__all__ = ["Kw_BatchNorm", "Kw_BatchNorm_dynamic"]


class _KwBnFn(torch.autograd.Function):
    """y = batch_norm(x) over keyword rows; x (M,D) fp32; parameters addressed as p[g*gstride + d*dstride]."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, n_groups, gstride, dstride, row_valid, training,
                momentum, eps):
        lib = _lib.load()
        _lib.require_cuda(x, "Kw_BatchNorm")
        M, D = x.shape
        dev = x.device
        xc = x.detach()
        if xc.dtype != torch.float32 or not xc.is_contiguous():
            xc = xc.float().contiguous()
        y = torch.empty_like(xc)
        save_mean = torch.empty((n_groups, D), dtype=torch.float32, device=dev)
        save_rstd = torch.empty((n_groups, D), dtype=torch.float32, device=dev)
        ws_bytes = lib.scp_kwbn_workspace_bytes(M, D, n_groups)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = lib.scp_kwbn_fwd(_lib.ptr(xc), M, D, n_groups, gstride, dstride, _lib.ptr(row_valid),
                                  _lib.ptr(gamma.detach()), _lib.ptr(beta.detach()), _lib.ptr(running_mean),
                                  _lib.ptr(running_var), int(training), float(momentum), float(eps), _lib.ptr(y),
                                  _lib.ptr(save_mean), _lib.ptr(save_rstd), _lib.ptr(ws), ws_bytes,
                                  _lib.stream_ptr(dev))
        _lib.check(st, "scp_kwbn_fwd")
        ctx.save_for_backward(xc, gamma.detach(), save_mean, save_rstd)
        ctx.row_valid = row_valid
        ctx.cfg = (n_groups, gstride, dstride, bool(training))
        ctx.param_shape = gamma.shape
        ctx.in_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, g_y):
        lib = _lib.load()
        xc, gamma, save_mean, save_rstd = ctx.saved_tensors
        n_groups, gstride, dstride, training = ctx.cfg
        M, D = xc.shape
        dev = xc.device
        g = g_y.float().contiguous()
        g_x = torch.empty_like(xc)
        need_params = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        g_gamma = torch.zeros(ctx.param_shape, dtype=torch.float32, device=dev) if need_params else None
        g_beta = torch.zeros(ctx.param_shape, dtype=torch.float32, device=dev) if need_params else None
        ws_bytes = lib.scp_kwbn_workspace_bytes(M, D, n_groups)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = lib.scp_kwbn_bwd(_lib.ptr(g), _lib.ptr(xc), M, D, n_groups, gstride, dstride, _lib.ptr(ctx.row_valid),
                                  _lib.ptr(gamma), _lib.ptr(save_mean), _lib.ptr(save_rstd), int(training),
                                  _lib.ptr(g_x), _lib.ptr(g_gamma), _lib.ptr(g_beta), _lib.ptr(ws), ws_bytes,
                                  _lib.stream_ptr(dev))
        _lib.check(st, "scp_kwbn_bwd")
        return (g_x.to(ctx.in_dtype), g_gamma if ctx.needs_input_grad[1] else None,
                g_beta if ctx.needs_input_grad[2] else None, None, None, None, None, None, None, None, None, None)


def _run_bn(bn: nn.BatchNorm1d, x2d: torch.Tensor, n_groups: int, gstride: int, dstride: int,
            row_valid: Optional[torch.Tensor], module_training: bool, n_stat_rows: int) -> torch.Tensor:
    """One BatchNorm1d parameter holder applied to (M,D) keyword rows with torch's train / eval semantics."""
    use_batch_stats = module_training or bn.running_mean is None
    if use_batch_stats and n_stat_rows <= 1:
        # same failure as torch.nn.functional.batch_norm
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {list(x2d.shape)}")
    if module_training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return _KwBnFn.apply(x2d, bn.weight, bn.bias, bn.running_mean if bn.track_running_stats else None,
                         bn.running_var if bn.track_running_stats else None, n_groups, gstride, dstride, row_valid,
                         use_batch_stats, momentum, bn.eps)


class Kw_BatchNorm(nn.Module):
    """BatchNorm layer for a fixed number of keywords (kw_bn.py:8-164)."""

    def __init__(self, kw_num: int, kw_dim: int, batchnorm_type: str, init_bias: torch.Tensor,
                 init_scale: torch.Tensor, std_scale: int = 1, learnable: bool = True, parallel: bool = False) -> None:
        super().__init__()
        self.batchnorm_type = batchnorm_type
        self.kw_num = kw_num
        self.kw_dim = kw_dim
        self.std_scale = std_scale
        self.learnable = learnable
        self.parallel = parallel
        if self.batchnorm_type == "eachKw":
            if self.parallel:
                self.bn_layer = nn.BatchNorm1d(kw_dim * self.kw_num)
            else:
                self.bn_layers = nn.ModuleList([nn.BatchNorm1d(kw_dim) for _ in range(self.kw_num)])
        elif self.batchnorm_type == "same":
            self.bn_layer = nn.BatchNorm1d(kw_dim)
        else:
            raise NotImplementedError()
        if not isinstance(self.std_scale, list):
            self.std_scale = [self.std_scale] * self.kw_num
        self.init_bn(init_bias, init_scale)
        logger.info("Initialize BatchNorm(%s) weight and bias learnable=(%s) with token embeddings w/ scale=%s, "
                    "parallel=(%s)", self.batchnorm_type, self.learnable, self.std_scale, self.parallel)

    def init_bn(self, init_bias: torch.Tensor, init_scale: torch.Tensor) -> None:
        """kw_bn.py:68-95."""
        if self.batchnorm_type == "eachKw":
            if self.parallel:
                self.bn_layer.weight.data.copy_((init_scale * self.std_scale[0]).repeat(self.kw_num))
                self.bn_layer.bias.data.copy_(init_bias.repeat(self.kw_num))
                self.bn_layer.weight.requires_grad = self.learnable
                self.bn_layer.bias.requires_grad = self.learnable
            else:
                for i, _bn_layer in enumerate(self.bn_layers):
                    _bn_layer.weight.data.copy_(init_scale * self.std_scale[i])
                    _bn_layer.bias.data.copy_(init_bias)
                    _bn_layer.weight.requires_grad = self.learnable
                    _bn_layer.bias.requires_grad = self.learnable
        elif self.batchnorm_type == "same":
            self.bn_layer.weight.data.copy_(init_scale * self.std_scale[0])
            self.bn_layer.bias.data.copy_(init_bias)
            self.bn_layer.weight.requires_grad = self.learnable
            self.bn_layer.bias.requires_grad = self.learnable

    def forward(self, keywords: torch.Tensor, seq_lens: torch.Tensor = None) -> torch.Tensor:
        assert keywords.dim() == 3
        assert keywords.shape[2] == self.kw_dim
        if seq_lens is None:
            assert keywords.shape[1] == self.kw_num
        bsz, n_kw, D = keywords.shape
        x2d = keywords.reshape(bsz * n_kw, D)
        if self.batchnorm_type == "eachKw":
            if self.parallel:
                # BatchNorm1d(D*K) on the (B, D*K) view of (B,D,K): feature index d*K + k  (kw_bn.py:119-127)
                y = _run_bn(self.bn_layer, x2d, self.kw_num, 1, self.kw_num, None, self.training, bsz)
            else:
                # K independent layers (kw_bn.py:128-140): each one normalises the rows of its keyword slot
                outs = []
                for i in range(self.kw_num):
                    outs.append(_run_bn(self.bn_layers[i], keywords[:, i].reshape(bsz, D), 1, 0, 1, None,
                                        self.training, bsz))
                return torch.stack(outs, dim=1)
        elif self.batchnorm_type == "same":
            if seq_lens is None:
                y = _run_bn(self.bn_layer, x2d, 1, 0, 1, None, self.training, bsz * n_kw)
            else:
                # statistics over the first seq_lens[b] keywords of every utterance only; the other rows pass through
                # (kw_bn.py:141-159; the reference writes the result back into `keywords` in place)
                assert seq_lens.dim() == 1
                lens = seq_lens.to(keywords.device)
                valid = (torch.arange(n_kw, device=keywords.device)[None, :] < lens[:, None]).reshape(-1)
                y = _run_bn(self.bn_layer, x2d, 1, 0, 1, valid.to(torch.uint8).contiguous(), self.training,
                            int(seq_lens.sum()))
        else:
            raise NotImplementedError()
        return y.reshape(bsz, n_kw, D)


class Kw_BatchNorm_dynamic(nn.Module):
    """BatchNorm layer for a dynamic number of keywords (kw_bn.py:167-228): one BatchNorm1d(D) over all (B,T') rows."""

    def __init__(self, kw_dim: int, init_bias: torch.Tensor, init_scale: torch.Tensor, std_scale: int = 1,
                 learnable: bool = True) -> None:
        super().__init__()
        self.kw_dim = kw_dim
        self.learnable = learnable
        assert std_scale > 0, f"std scale must > 0, but input std scale is {std_scale}"
        self.std_scale = std_scale
        self.bn_layer = nn.BatchNorm1d(kw_dim)
        self.init_bn(init_bias, init_scale)
        logger.info("Initialize BatchNorm weight and bias learnable=(%s) with token embeddings w/ scale=%s",
                    self.learnable, self.std_scale)

    def init_bn(self, init_bias: torch.Tensor, init_scale: torch.Tensor) -> None:
        self.bn_layer.weight.data.copy_(init_scale * self.std_scale)
        self.bn_layer.bias.data.copy_(init_bias)
        self.bn_layer.weight.requires_grad = self.learnable
        self.bn_layer.bias.requires_grad = self.learnable

    def forward(self, keywords: torch.Tensor) -> torch.Tensor:
        assert keywords.dim() == 3
        bsz, n_kw, D = keywords.shape
        y = _run_bn(self.bn_layer, keywords.reshape(bsz * n_kw, D), 1, 0, 1, None, self.training, bsz * n_kw)
        return y.reshape(bsz, n_kw, D)
