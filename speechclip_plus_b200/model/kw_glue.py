"""Host-side glue of the hot path, mirroring the ~40 lines of reference glue around the three modules:

  * the L2 normalisation of image / cascaded-audio / parallel-audio features
    (avssl/model/kwClip.py:857, :905-907, :913-915),
  * the gather point of the data-parallel step (``training_step_end``, kwClip.py:149-193): the reference lets
    ``nn.DataParallel`` copy every replica's ``loss_feats`` to GPU 0; here each process packs its normalised features
    and ids into one buffer (scp_l2norm_pack) and a single NCCL all-gather over NVLink delivers the global batch to
    every rank, which then evaluates the global-negatives loss redundantly and back-propagates into its own rows,
  * ``compute_loss`` (kwClip.py:999-1040).

The collective itself is plain ``torch.distributed`` plumbing (one all-gather of <= 0.8 MB per rank: latency-bound).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import _lib

FEAT_KEYS = ("image_feat", "cascaded_audio_feat", "parallel_audio_feat")

# pass as `group` to keep a call purely local even though torch.distributed is initialised (e.g. a single-process
# reference evaluation of the concatenated batch on rank 0)
LOCAL = "local"


def group_world_size(group) -> int:
    if group is LOCAL or not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


# ---------------------------------------------------------------------------------------------------------------
# device-agnostic plumbing (exercised on CPU with gloo in tests/test_multiproc_gloo.py)
# ---------------------------------------------------------------------------------------------------------------
def shard_rows(n_local: int, rank: int) -> Tuple[int, int]:
    """Rows of the gathered global batch that belong to ``rank`` (every rank contributes ``n_local`` rows)."""
    return rank * n_local, (rank + 1) * n_local


def pack_nbytes(n_feats: int, n: int, D: int) -> int:
    return n_feats * n * D * 4 + n * 8


def all_gather_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather one packed uint8 buffer per rank -> (world, nbytes).  Works for NCCL (device) and gloo (host)."""
    world = group_world_size(group)
    if world == 1:
        return packed.reshape(1, -1)
    if packed.is_cuda and packed.numel() * packed.element_size() % 16 == 0:
        from . import peer_gather  # one push + one collect launch over NVLink peer memory instead of an NCCL collective
        ctx = peer_gather.get(packed.numel() * packed.element_size(), packed.device, group, tag="feats")
        if ctx is not None:
            return ctx.all_gather(packed).view(packed.dtype).reshape(world, -1)
    out = torch.empty(world * packed.numel(), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed.reshape(-1), group=group)  # flat output: accepted by NCCL and gloo
    return out.reshape(world, -1)


def gather_packed_segments(packed: torch.Tensor, n_feats: int, n: int, D: int, group=None):
    """The peer-memory exchange with a segment-major result (every feature block and the ids come out as one contiguous
    gathered array): ``(flat uint8 (world * nbytes,), world)``, or ``None`` when that path is unavailable (host tensors,
    no peer access, segment sizes that are not multiples of 16) and :func:`all_gather_packed` has to do it."""
    world = group_world_size(group)
    seg = [n * D * 4] * n_feats + [n * 8]
    if world == 1 or not packed.is_cuda or any(b % 16 for b in seg):
        return None
    from . import peer_gather
    ctx = peer_gather.get(packed.numel() * packed.element_size(), packed.device, group, tag="feats")
    if ctx is None:
        return None
    return ctx.all_gather_segments(packed, seg), world


def unpack_segments(flat: torch.Tensor, world: int, n_feats: int, n: int, D: int) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """Views into the segment-major result of :func:`gather_packed_segments` (no copies)."""
    fbytes = world * n * D * 4
    feats = [flat[f * fbytes:(f + 1) * fbytes].view(torch.float32).reshape(world * n, D) for f in range(n_feats)]
    ids = flat[n_feats * fbytes:n_feats * fbytes + world * n * 8].view(torch.int64)
    return feats, ids


def unpack_gathered(gathered: torch.Tensor, n_feats: int, n: int, D: int) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """(world, nbytes) uint8 -> ([ (world*n, D) float32 ] * n_feats, (world*n,) int64), rank-major row order."""
    world = gathered.shape[0]
    fbytes = n * D * 4
    feats = []
    for f in range(n_feats):
        blk = gathered[:, f * fbytes:(f + 1) * fbytes].contiguous().view(torch.float32)
        feats.append(blk.reshape(world * n, D))
    ids = gathered[:, n_feats * fbytes:n_feats * fbytes + n * 8].contiguous().view(torch.int64).reshape(world * n)
    return feats, ids


def ddp_grad_scale(world_size: int) -> float:
    """Every rank back-propagates the GLOBAL loss into its local rows only.  DDP then averages parameter gradients
    over ranks; multiplying the loss (or the gradients) by ``world_size`` makes the all-reduced result equal to the
    reference's single-process gradient of the same global loss."""
    return float(world_size)


# ---------------------------------------------------------------------------------------------------------------
# CUDA pieces
# ---------------------------------------------------------------------------------------------------------------
class _NormPackGatherFn(torch.autograd.Function):
    """(ids, feats...) -> (global ids, global normalised feats...); gradients flow to the local feature rows only."""

    @staticmethod
    def forward(ctx, ids: torch.Tensor, group, *feats: torch.Tensor):
        lib = _lib.load()
        f0 = feats[0]
        _lib.require_cuda(f0, "gather_loss_feats")
        n, D = f0.shape
        dev = f0.device
        # one storage type for the pack kernel: the features' own type when they agree, else fp32 (never round an fp32
        # audio feature to the image tower's half precision: the reference normalises each tensor in its own type and
        # calls .float() before the criterion, kwClip.py:1015-1028)
        common = f0.dtype if all(f.dtype == f0.dtype for f in feats) else torch.float32
        srcs = []
        for f in feats:
            assert f.shape == (n, D), (f.shape, (n, D))
            fd = f.detach()
            if fd.dtype != common:
                fd = fd.to(common)
            srcs.append(fd.contiguous())
        ids64 = ids.detach().reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
        nbytes = int(lib.scp_pack_bytes(len(srcs), n, D))
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        inv = torch.empty((len(srcs), n), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = lib.scp_l2norm_pack(_lib.ptr_array(srcs), len(srcs), n, D, _lib.dtype_code(common),
                                     _lib.ptr(ids64), _lib.ptr(packed), _lib.ptr(inv), _lib.stream_ptr(dev))
        _lib.check(st, "scp_l2norm_pack")
        seg = gather_packed_segments(packed, len(srcs), n, D, group)
        if seg is not None:  # one push + one segment-major collect: the gathered arrays are views, nothing to copy
            g_feats, g_ids = unpack_segments(seg[0], seg[1], len(srcs), n, D)
            world = seg[1]
        else:
            gathered = all_gather_packed(packed, group)
            g_feats, g_ids = unpack_gathered(gathered, len(srcs), n, D)
            world = gathered.shape[0]
        rank = dist.get_rank(group) if world > 1 else 0
        ctx.rows = shard_rows(n, rank)
        ctx.n_feats = len(srcs)
        ctx.in_dtypes = [f.dtype for f in feats]
        ctx.save_for_backward(inv, *[g[ctx.rows[0]:ctx.rows[1]] for g in g_feats])
        ctx.mark_non_differentiable(g_ids)
        ctx.set_materialize_grads(False)  # unused outputs (ids, features of an inactive branch) get None, not zeros
        return (g_ids, *g_feats)

    @staticmethod
    def backward(ctx, _g_ids, *g_feats):
        lib = _lib.load()
        inv, *f_hat = ctx.saved_tensors
        r0, r1 = ctx.rows
        grads = []
        for f in range(ctx.n_feats):
            g = g_feats[f]
            if g is None or not ctx.needs_input_grad[2 + f]:
                grads.append(None)
                continue
            g_loc = g[r0:r1].float().contiguous()
            fh = f_hat[f].contiguous()
            n, D = fh.shape
            out = torch.empty_like(fh)
            with torch.cuda.device(fh.device):
                st = lib.scp_l2norm_bwd(_lib.ptr(g_loc), _lib.ptr(fh), _lib.ptr(inv[f].contiguous()), n, D,
                                        _lib.ptr(out), _lib.stream_ptr(fh.device))
            _lib.check(st, "scp_l2norm_bwd")
            grads.append(out.to(ctx.in_dtypes[f]))
        return (None, None, *grads)


def gather_loss_feats(loss_feats: Dict[str, torch.Tensor], group=None) -> Tuple[Dict[str, torch.Tensor], Tuple[int, int]]:
    """Normalise (kwClip.py:857/:905/:913), pack and all-gather the UN-normalised ``image_feat`` /
    ``cascaded_audio_feat`` / ``parallel_audio_feat`` entries and ``id`` of ``loss_feats``.

    Returns the global ``loss_feats`` dict (same keys) and this rank's ``(row_begin, row_end)``."""
    keys = [k for k in FEAT_KEYS if loss_feats.get(k) is not None]
    feats = [loss_feats[k] for k in keys]
    outs = _NormPackGatherFn.apply(loss_feats["id"], group, *feats)
    g_ids, g_feats = outs[0], outs[1:]
    out = dict(loss_feats)
    out["id"] = g_ids
    n_rows = feats[0].shape[0]
    rows = shard_rows(n_rows, dist.get_rank(group) if g_ids.shape[0] > n_rows else 0)
    for k, g in zip(keys, g_feats):
        # the backward of this gather reads ONLY this rank's rows of the gradient: consumers that know their local rows
        # (MaskedContrastiveLoss with local_rows) may leave the other rows of the gradient unwritten
        g._scp_local_rows = rows
        out[k] = g
    n = feats[0].shape[0]
    world = g_ids.shape[0] // n
    rank = dist.get_rank(group) if world > 1 else 0
    return out, shard_rows(n, rank)


_SIDE_STREAMS = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def _overlap_enabled() -> bool:
    import os
    return os.environ.get("SCP_NCE_OVERLAP", "1") != "0"


def _world(group) -> int:
    return group_world_size(group)


def global_mean_with_local_grad(local_mean: torch.Tensor, group=None) -> torch.Tensor:
    """A per-rank mean -> a tensor whose VALUE is the mean over all ranks (what the reference's DataParallel gather +
    ``nn.L1Loss`` produces, kwClip.py:1031-1038) and whose gradient is ``d local_mean / world``: after the loss is
    multiplied by ``ddp_grad_scale(world)`` and DDP averages the parameter gradients over ranks, the result equals the
    gradient of the global mean.  Equal per-rank batch sizes are assumed (the data-parallel sampler guarantees them)."""
    world = _world(group)
    if world == 1:
        return local_mean
    g = local_mean.detach().clone()
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    g = g / world
    return g + (local_mean - local_mean.detach()) / world


def compute_loss(loss_feats: Dict[str, torch.Tensor], criterion, cascaded_objective_weight: float = 0.0,
                 parallel_objective_weight: float = 0.0, quantity_loss_weight: float = 0.0,
                 quantity_loss_criteria=None, local_rows: Optional[Tuple[int, int]] = None,
                 group=None) -> Dict[str, torch.Tensor]:
    """``KWClip_GeneralTransformer.compute_loss`` (kwClip.py:999-1040) with the same keys in and out.

    ``local_rows`` / ``group`` (one process per GPU): ``loss_feats`` holds the gathered global batch, the contrastive
    terms back-propagate into this rank's rows only, and the CIF quantity loss -- whose inputs stay rank-local -- is
    turned into the global mean (see :func:`global_mean_with_local_grad`)."""
    assert isinstance(loss_feats, dict)
    required_keys = {"id", "image_feat"}
    assert required_keys.issubset(set(loss_feats.keys())), f"required: {required_keys}, input: {loss_feats.keys()}"
    losses = {"loss": 0}
    image_feat = loss_feats["image_feat"].float()
    ids = loss_feats["id"]
    active = [(branch, weight) for branch, weight in (("cascaded", cascaded_objective_weight),
                                                      ("parallel", parallel_objective_weight)) if weight > 0.0]
    kwargs = {} if local_rows is None else {"local_rows": local_rows, "group": group}
    terms = {}
    # Hybrid models call the criterion twice on the same image features and ids (kwClip.py:1015-1028).  Each call is a
    # chain of short, latency-bound launches; the second chain runs on a helper stream next to the first (fork / join by
    # events: capturable), and autograd replays each backward on the stream of its forward, so both directions overlap.
    overlap = (len(active) == 2 and image_feat.is_cuda and _overlap_enabled())
    side = _side_stream(image_feat.device) if overlap else None
    def call(branch):
        key = f"{branch}_audio_feat"
        assert key in loss_feats, f"{loss_feats.keys()}"
        return criterion(feat_A=loss_feats[key].float(), feat_B=image_feat, index=ids, **kwargs)

    if side is not None:
        main = torch.cuda.current_stream(image_feat.device)
        side.wait_stream(main)                       # fork: everything issued so far (gather, normalise) is visible
        with torch.cuda.stream(side):
            terms[active[1][0]] = call(active[1][0])
        terms[active[0][0]] = call(active[0][0])     # the first chain on the caller's stream, next to the second
        main.wait_stream(side)                       # join (the next fork orders any reuse of the side stream's blocks)
    else:
        for branch, _ in active:
            terms[branch] = call(branch)
    for branch, weight in active:
        losses[f"{branch[0]}_cl_loss"] = terms[branch]
        # `loss += weight * term` of the reference, without the no-op kernels for `0 + x` and `1.0 * x`
        term = terms[branch]
        if weight != 1.0:
            term = weight * term
        losses["loss"] = term if isinstance(losses["loss"], int) else losses["loss"] + term
    if ("cif_quantity_out" in loss_feats and "cif_target_len" in loss_feats and quantity_loss_criteria is not None):
        q_loss = quantity_loss_criteria(loss_feats["cif_quantity_out"], loss_feats["cif_target_len"])
        if local_rows is not None:
            q_loss = global_mean_with_local_grad(q_loss, group)
        losses["quantity_loss"] = q_loss
        losses["loss"] = losses["loss"] + quantity_loss_weight * losses["quantity_loss"]
    return losses
