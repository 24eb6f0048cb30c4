"""Tail of the data-parallel training step for the hot path's own trainable tensors.

The reference trains ``audio_encoder.weightedsum_layer.weights`` and ``criterion.temperature`` (plus the branch / HuBERT
parameters that are out of scope here) in ONE ``torch.optim.Adam`` group (avssl/model/kwClip.py:636-668; lr 1e-4,
weight_decay 1e-6, config/speechCLIP+/model_base/spchclip_c+.yaml:119-123) after ``nn.DataParallel`` has summed the
replicas' gradients onto GPU 0 (``reduce_add_coalesced``).  In the one-process-per-GPU layout every
rank back-propagates the global loss into its own rows, so the full gradient is the SUM over ranks:

    scp_grad_pack  ->  one all-reduce (SUM) of the packed buffer  ->  scp_adam_packed   (in place, every rank identical)

Three launches / one collective per step regardless of the number of registered tensors; capturable in a CUDA graph (the
Adam step counter lives on the device).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from .. import _lib


class PackedAdam:
    """``torch.optim.Adam(params, lr, betas, eps, weight_decay)`` for a handful of small fp32 tensors on one device."""

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, group=None, all_reduce: bool = True):
        self.params: List[torch.Tensor] = [p for p in params]
        if not 1 <= len(self.params) <= _lib.SCP_MAX_PACKED:
            raise _lib.ScpError(f"PackedAdam takes 1..{_lib.SCP_MAX_PACKED} tensors, got {len(self.params)}")
        p0 = self.params[0]
        for p in self.params:
            _lib.require_cuda(p, "PackedAdam")
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != p0.device:
                raise _lib.ScpError("PackedAdam: parameters must be contiguous fp32 tensors on one device")
        self.lr, self.betas, self.eps, self.weight_decay, self.group = float(lr), betas, float(eps), float(weight_decay), group
        self.all_reduce = all_reduce  # False: a purely local optimiser even inside an initialised process group
        self.sizes = [p.numel() for p in self.params]
        total = sum(self.sizes)
        self.total = total
        self.padded = (total + 3) // 4 * 4  # the peer-memory exchange moves 16-byte units
        dev = p0.device
        self.packed = torch.zeros(self.padded, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.padded, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.padded, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_device: Optional[torch.Tensor] = None  # set to a device scalar to drive the rate from a scheduler
        self._sizes_arr = (ctypes.c_int64 * len(self.sizes))(*self.sizes)
        self._param_ptrs = _lib.ptr_array(self.params)

    def world(self) -> int:
        if not self.all_reduce:
            return 1
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def pack_grads(self, grads: Optional[List[Optional[torch.Tensor]]] = None, scale: float = 1.0) -> torch.Tensor:
        lib = _lib.load()
        grads = [p.grad for p in self.params] if grads is None else grads
        arr = (ctypes.c_void_p * len(grads))()
        keep = []
        for i, (g, p) in enumerate(zip(grads, self.params)):
            if g is None:
                arr[i] = None
                continue
            g = g.detach()
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            assert g.numel() == p.numel(), (g.shape, p.shape)
            keep.append(g)
            arr[i] = g.data_ptr()
        dev = self.packed.device
        with torch.cuda.device(dev):
            st = lib.scp_grad_pack(arr, self._sizes_arr, len(grads), float(scale), _lib.ptr(self.packed),
                                   _lib.stream_ptr(dev))
        _lib.check(st, "scp_grad_pack")
        return self.packed

    def step(self, grads: Optional[List[Optional[torch.Tensor]]] = None, grad_scale: float = 1.0) -> None:
        """Pack this rank's gradients, SUM them over the group, apply Adam in place."""
        lib = _lib.load()
        self.pack_grads(grads)
        src, n_shards, stride = self.packed, 1, 0
        if self.world() > 1:
            from . import peer_gather
            ctx = peer_gather.get(self.padded * 4, self.packed.device, self.group, tag="grads")
            if ctx is not None:  # every rank receives every rank's packed gradients and sums them in rank order
                src = ctx.all_gather(self.packed).view(torch.float32)
                n_shards, stride = self.world(), self.padded
                self.gathered = src
            else:
                dist.all_reduce(self.packed, op=dist.ReduceOp.SUM, group=self.group)
        dev = self.packed.device
        with torch.cuda.device(dev):
            st = lib.scp_adam_packed(self._param_ptrs, self._sizes_arr, len(self.params), _lib.ptr(src), n_shards, stride,
                                     float(grad_scale), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                     _lib.ptr(self.step_count), _lib.ptr(self.lr_device), self.lr, float(self.betas[0]),
                                     float(self.betas[1]), self.eps, self.weight_decay, _lib.stream_ptr(dev))
        _lib.check(st, "scp_adam_packed")

    def unpacked_grads(self) -> List[torch.Tensor]:
        """Views of the (all-reduced) packed gradient buffer, one per registered tensor."""
        out, off = [], 0
        for p, n in zip(self.params, self.sizes):
            out.append(self.packed[off:off + n].view_as(p))
            off += n
        return out
