"""One-shot all-gather over NVLink peer memory (csrc/scp_p2p.cu) behind the three small exchanges of the step:

  * the packed, normalised loss features + ids                (G0: avssl/model/kwClip.py:149-169),
  * the (3, n) InfoNCE statistics of the sharded forward      (losses.py:224-243 per shard),
  * the packed gradients of the path's trainable tensors      (the reference's DataParallel reduce_add_coalesced).

``torch.distributed`` is used once, to allocate and exchange the symmetric buffers
(``torch.distributed._symmetric_memory``); every exchange afterwards is one push kernel + one collect kernel on the
caller's stream.  If symmetric memory is unavailable (different nodes, no peer access) the callers fall back to NCCL.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from .. import _lib

logger = logging.getLogger(__name__)

_CONTEXTS: Dict[Tuple, Optional["PeerGather"]] = {}


def enabled() -> bool:
    return os.environ.get("SCP_P2P_GATHER", "1") != "0"


class PeerGather:
    """Symmetric gather buffer of one (group, device, capacity) + the device-side epoch state."""

    def __init__(self, capacity_bytes: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        lib = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.capacity = (int(capacity_bytes) + 255) // 256 * 256
        total = int(lib.scp_p2p_buffer_bytes(self.world, self.capacity))
        try:  # older releases need the group registered first; newer ones do it inside rendezvous
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                symm_mem.enable_symm_mem_for_group(self.group.group_name)
        except Exception:
            pass
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(self.buf, group=self.group)
        self.buf.zero_()
        self.state = torch.zeros(2, dtype=torch.int32, device=device)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self.buf.data_ptr(), (ptrs, self.buf.data_ptr())
        self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self._hdl = hdl
        torch.cuda.synchronize(device)
        dist.barrier(self.group)  # every rank's flags are zero before anyone pushes

    def all_gather(self, src: torch.Tensor) -> torch.Tensor:
        """src: contiguous tensor of <= capacity bytes (multiple of 16) -> (world, nbytes) uint8."""
        lib = _lib.load()
        flat = src.reshape(-1).view(torch.uint8)
        nbytes = flat.numel()
        out = torch.empty((self.world, nbytes), dtype=torch.uint8, device=flat.device)
        with torch.cuda.device(flat.device):
            st = lib.scp_p2p_allgather(_lib.ptr(flat), nbytes, _lib.ptr(self.peer_ptrs), _lib.ptr(self.buf), self.rank,
                                       self.world, self.capacity, _lib.ptr(self.state), _lib.ptr(out),
                                       _lib.stream_ptr(flat.device))
        _lib.check(st, "scp_p2p_allgather")
        return out


    def all_gather_segments(self, src: torch.Tensor, seg_bytes) -> torch.Tensor:
        """Like :meth:`all_gather`, but the result is SEGMENT-major: for every segment of the payload (``seg_bytes``:
        sizes in bytes, multiples of 16, adding up to the payload) the world copies follow each other in rank order, so
        each segment of the flat ``(world * nbytes,)`` result is one contiguous gathered array (no copies to unpack)."""
        import ctypes
        lib = _lib.load()
        flat = src.reshape(-1).view(torch.uint8)
        nbytes = flat.numel()
        segs = (ctypes.c_int64 * len(seg_bytes))(*[int(b) for b in seg_bytes])
        out = torch.empty(self.world * nbytes, dtype=torch.uint8, device=flat.device)
        with torch.cuda.device(flat.device):
            st = lib.scp_p2p_allgather_segments(_lib.ptr(flat), nbytes, _lib.ptr(self.peer_ptrs), _lib.ptr(self.buf),
                                                self.rank, self.world, self.capacity, _lib.ptr(self.state), segs,
                                                len(seg_bytes), _lib.ptr(out), _lib.stream_ptr(flat.device))
        _lib.check(st, "scp_p2p_allgather_segments")
        return out


def get(capacity_bytes: int, device: torch.device, group=None, tag: str = "") -> Optional[PeerGather]:
    """Cached context (created collectively on first use: every rank must reach this call).  None: use NCCL."""
    if not enabled() or not (dist.is_available() and dist.is_initialized()):
        return None
    key = (tag, id(group), device.index, (int(capacity_bytes) + 255) // 256 * 256)
    if key not in _CONTEXTS:
        if torch.cuda.is_current_stream_capturing():
            return None  # contexts are created eagerly (warm-up steps); never inside a capture
        try:
            _CONTEXTS[key] = PeerGather(capacity_bytes, device, group)
        except Exception as exc:  # no peer access / symmetric memory not available: NCCL does the exchange
            logger.warning("peer-memory all-gather unavailable (%s): falling back to NCCL", repr(exc)[:200])
            _CONTEXTS[key] = None
    return _CONTEXTS[key]
