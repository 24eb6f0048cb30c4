"""Drop-in for ``avssl.module.losses.MaskedContrastiveLoss`` (reference: avssl/module/losses.py:129-245).

Same constructor kwargs, the same ``temperature`` attribute (0-d log-scale Parameter when trainable, python float
``1/temperature`` otherwise), the ``current_temperature`` property, the ``eye_mat / neg_eye_mat / eye_mat_fl``
buffers (kept only so that released checkpoints load; the kernels never read them) and
``forward(feat_A, feat_B, index=None) -> 0-d Tensor``.  Differences, all deliberate:
  * no batch-size limit (the reference raises IndexError for N > MAX_EYE = 256, losses.py:126);
  * log-sum-exp with max-subtraction instead of a raw ``exp`` (losses.py:232): identical value, no overflow;
  * no host synchronisation inside ``forward`` (the reference's boolean-mask indexing syncs twice per call).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from .. import _lib

MAX_EYE = 256  # losses.py:126 -- only the size of the compatibility buffers


class _MaskedNceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat_a, feat_b, log_scale, index, fixed_scale, margin, dcl, a2b, b2a, row_begin, row_end,
                group=None, sharded=False):
        lib = _lib.load()
        _lib.require_cuda(feat_a, "MaskedContrastiveLoss")
        N, D = feat_a.shape
        dev = feat_a.device
        a = feat_a.detach()
        b = feat_b.detach()
        if a.dtype != torch.float32 or not a.is_contiguous():
            a = a.float().contiguous()
        if b.dtype != torch.float32 or not b.is_contiguous():
            b = b.float().contiguous()
        ids = None
        if index is not None:
            ids = index.detach().reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
            assert ids.shape[0] == N, (ids.shape, feat_a.shape)  # losses.py:205
        ls = None
        if log_scale is not None:
            ls = log_scale.detach().reshape(1).float().contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        lse_row = torch.empty(N, dtype=torch.float32, device=dev)
        lse_col = torch.empty(N, dtype=torch.float32, device=dev)
        ws_bytes = lib.scp_nce_workspace_bytes(N, D)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        need_bwd = any(ctx.needs_input_grad[:3])
        if sharded:
            # SURVEY.md section 8(e) option B: this rank evaluates the denominators of its own rows / columns only and
            # the ranks exchange 12 bytes per sample (losses.py:224-243 restated per shard)
            n_local = int(row_end) - int(row_begin)
            world = N // n_local
            stats = torch.empty((3, n_local), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                st = lib.scp_nce_fwd_local(_lib.ptr(a), _lib.ptr(b), _lib.ptr(ids), N, D, _lib.ptr(ls),
                                           float(fixed_scale), float(margin), int(dcl), int(row_begin), int(row_end),
                                           int(need_bwd), _lib.ptr(stats), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
            _lib.check(st, "scp_nce_fwd_local")
            stats_all = all_gather_stats(stats, world, group)
            with torch.cuda.device(dev):
                st = lib.scp_nce_loss_from_stats(_lib.ptr(stats_all), world, n_local, _lib.ptr(ls), float(fixed_scale),
                                                 float(margin), int(a2b), int(b2a), _lib.ptr(loss), _lib.ptr(lse_row),
                                                 _lib.ptr(lse_col), _lib.stream_ptr(dev))
            _lib.check(st, "scp_nce_loss_from_stats")
        else:
            with torch.cuda.device(dev):
                st = lib.scp_nce_fwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(ids), N, D, _lib.ptr(ls), float(fixed_scale),
                                     float(margin), int(dcl), int(a2b), int(b2a), int(need_bwd), _lib.ptr(loss),
                                     _lib.ptr(lse_row), _lib.ptr(lse_col), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
            _lib.check(st, "scp_nce_fwd")
        # the workspace holds the split fp16 operands; keeping it alive lets the backward skip their re-computation
        ctx.save_for_backward(a, b, ids, ls, lse_row, lse_col, ws if need_bwd else None)
        ctx.cfg = (float(fixed_scale), float(margin), int(dcl), int(a2b), int(b2a), int(row_begin), int(row_end))
        ctx.in_dtypes = (feat_a.dtype, feat_b.dtype)
        # inputs gathered by gather_loss_feats: only rows [row_begin, row_end) of their gradient are ever read
        rows = (int(row_begin), int(row_end))
        ctx.rows_only = tuple(getattr(t, "_scp_local_rows", None) == rows for t in (feat_a, feat_b))
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g_loss):
        lib = _lib.load()
        a, b, ids, ls, lse_row, lse_col, ws = ctx.saved_tensors
        fixed_scale, margin, dcl, a2b, b2a, row_begin, row_end = ctx.cfg
        N, D = a.shape
        dev = a.device
        n_local = row_end - row_begin
        g = g_loss.detach().reshape(1).float().contiguous()
        need_b = ctx.needs_input_grad[1]
        need_t = ls is not None and ctx.needs_input_grad[2]
        # The kernel writes this rank's rows straight into the full-size gradient (rows of other ranks: zero, or left
        # unwritten when the producer of the input is known to read the local rows only) -- no zero-fill + copy afterwards.
        def grad_buffer(rows_only: bool, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
            if n_local == N or dtype != torch.float32:
                local = torch.empty((n_local, D), dtype=torch.float32, device=dev)
                return local, None
            # (anomaly detection inspects whole gradient tensors: give it zeros, not uninitialised memory)
            skip_fill = rows_only and not torch.is_anomaly_enabled()
            whole = (torch.empty if skip_fill else torch.zeros)((N, D), dtype=torch.float32, device=dev)
            return whole[row_begin:row_end], whole

        dA, dA_full = grad_buffer(ctx.rows_only[0], ctx.in_dtypes[0])
        dB, dB_full = grad_buffer(ctx.rows_only[1], ctx.in_dtypes[1]) if need_b else (None, None)
        dT = torch.empty(1, dtype=torch.float32, device=dev) if need_t else None
        ws_bytes = lib.scp_nce_workspace_bytes(N, D)
        state_valid = ws is not None
        if ws is None:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = lib.scp_nce_bwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(ids), N, D, _lib.ptr(ls), fixed_scale, margin,
                                 dcl, a2b, b2a, _lib.ptr(lse_row), _lib.ptr(lse_col), _lib.ptr(g), row_begin,
                                 row_end, int(state_valid), _lib.ptr(dA), _lib.ptr(dB), _lib.ptr(dT), _lib.ptr(ws),
                                 ws_bytes, _lib.stream_ptr(dev))
        _lib.check(st, "scp_nce_bwd")

        def full(local: Optional[torch.Tensor], whole: Optional[torch.Tensor], dtype) -> Optional[torch.Tensor]:
            if local is None:
                return None
            if whole is not None:
                return whole
            if n_local == N:
                return local.to(dtype)
            out = torch.zeros((N, D), dtype=dtype, device=dev)  # rows owned by other ranks get no gradient here
            out[row_begin:row_end] = local
            return out

        gA = full(dA, dA_full, ctx.in_dtypes[0]) if ctx.needs_input_grad[0] else None
        gB = full(dB, dB_full, ctx.in_dtypes[1])
        gT = dT.reshape(()) if dT is not None else None
        return gA, gB, gT, None, None, None, None, None, None, None, None, None, None


def all_gather_stats(stats: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """(3, n) per rank -> (world, 3, n), rank-major (the row order of ``gather_loss_feats``)."""
    if world == 1:
        return stats.reshape(1, *stats.shape)
    import torch.distributed as dist
    flat = stats.reshape(-1).contiguous()
    if flat.is_cuda and flat.numel() * 4 % 16 == 0:
        from ..model import peer_gather
        ctx = peer_gather.get(flat.numel() * 4, flat.device, group, tag="stats")
        if ctx is not None:
            return ctx.all_gather(flat).view(stats.dtype).reshape(world, *stats.shape)
    out = torch.empty(world * flat.numel(), dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.reshape(world, *stats.shape)


def _shard_min_n() -> int:
    import os
    return int(os.environ.get("SCP_NCE_SHARD_MIN_N", "2048"))


def _use_sharded_forward(N: int, begin: int, end: int, group) -> bool:
    """The sharded forward needs every rank to own an equal, rank-ordered slice of the gathered batch.  It costs one
    more (tiny) all-gather, ~25 us of NCCL latency inside a CUDA graph on NVLink, and saves (world-1)/world of the 2 N^2 D
    forward: measured on B200 it pays from N = 2048 (8 ranks x 256) on; below that every rank evaluates the global
    denominators redundantly (SCP_NCE_SHARD_MIN_N overrides the threshold)."""
    import torch.distributed as dist
    if N < _shard_min_n():
        return False
    if (begin, end) == (0, N) or group == "local" or not (dist.is_available() and dist.is_initialized()):
        return False
    world = dist.get_world_size(group)
    n = end - begin
    return world > 1 and n * world == N and begin == dist.get_rank(group) * n


class MaskedContrastiveLoss(nn.Module):
    def __init__(self, temperature: float = 0.07, temperature_trainable: bool = False, margin: float = 0.0,
                 dcl: bool = False, a2b: bool = True, b2a: bool = True):
        """Masked Contrastive Loss (losses.py:130-168)."""
        super().__init__()
        assert a2b or b2a, "Cannot set both `a2b` and `b2a` to False."
        self.temperature_trainable = temperature_trainable
        self.margin = margin
        self.dcl = dcl
        self.a2b = a2b
        self.b2a = b2a
        if temperature_trainable:
            self.temperature = nn.Parameter(torch.ones([]) * np.log(1 / temperature))
        else:
            self.temperature = 1 / temperature
        eye_mat = torch.eye(MAX_EYE, dtype=torch.bool)
        self.register_buffer("eye_mat", eye_mat)
        self.register_buffer("neg_eye_mat", ~eye_mat)
        self.register_buffer("eye_mat_fl", eye_mat.type(torch.float))

    @property
    def current_temperature(self) -> float:
        """losses.py:170-183 (a host read of the learnable scale; only used for logging)."""
        if self.temperature_trainable:
            temp = self.temperature.data.cpu().detach().float().exp().item()
        else:
            temp = self.temperature
        return float(temp)

    def forward(self, feat_A: torch.Tensor, feat_B: torch.Tensor, index: torch.LongTensor = None,
                local_rows: Optional[Tuple[int, int]] = None, group=None) -> torch.Tensor:
        """``local_rows=(begin, end)`` (extension for the one-process-per-GPU layout): the loss is the global one over
        all rows, gradients are produced only for this rank's rows -- see ``model.kw_glue.gather_loss_feats``.  When the
        rows are this rank's equal share of a ``torch.distributed`` group, the forward is sharded as well: each rank
        computes the denominators of its own rows / columns and one small all-gather completes the loss."""
        assert feat_A.shape == feat_B.shape, (feat_A.shape, feat_B.shape)  # losses.py:199
        N = feat_A.shape[0]
        begin, end = (0, N) if local_rows is None else local_rows
        sharded = _use_sharded_forward(N, begin, end, group)
        margin = self.margin if self.margin > 0 else 0.0  # losses.py:227: the margin applies only when positive
        if self.temperature_trainable:
            return _MaskedNceFn.apply(feat_A, feat_B, self.temperature, index, 0.0, margin, self.dcl,
                                      self.a2b, self.b2a, begin, end, group, sharded)
        return _MaskedNceFn.apply(feat_A, feat_B, None, index, float(self.temperature), margin, self.dcl,
                                  self.a2b, self.b2a, begin, end, group, sharded)
