"""Drop-in for ``avssl.module.speechclip_c_modules.vector_quantizers`` (the getattr namespace of
avssl/model/kw_branches.py:75-91) and the fused keyword quantiser behind ``GeneralBranch.vq_audio_features``.

``SimpleVectorQuantizer`` keeps the reference constructor, the ``curr_temp`` buffer/parameter (state_dict key
``...vector_quantizer.curr_temp``), ``set_num_updates`` and ``forward(x, prob_msk, produce_targets) -> dict``
(reference: avssl/module/speechclip_c_modules/my_vector_quantizer.py:12-165).

Two entry points:
  * ``forward(x)``          -- the reference signature: x is the dense (B,K,V) cosine-score tensor.  Runs the
                               dense kernels (scp_vq_dense_*); masks x in place exactly like the reference.
  * ``quantize_keywords()`` -- the fused hot path: takes the keyword vectors and the frozen CLIP token table,
                               runs cosine + mask + arg-max + softmax statistics + lookup on the tensor cores without
                               ever materialising the (B,K,V) logits (scp_vq_fwd / scp_vq_bwd).
"""
from __future__ import annotations

import ast
import ctypes
import logging
import threading
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import _lib

logger = logging.getLogger(__name__)

__all__ = ["SimpleVectorQuantizer", "TokenTableCache", "fused_vq_audio_features"]


def _masked_array(prob_msk: Sequence[int], V: int):
    cols = [int(c) for c in prob_msk]
    if len(cols) > _lib.SCP_MAX_MASKED:
        raise _lib.ScpError(f"at most {_lib.SCP_MAX_MASKED} masked columns are supported, got {len(cols)}")
    for c in cols:
        if not 0 <= c < V:
            raise IndexError(f"prob_msk column {c} outside [0,{V})")  # the reference's x[:, i] would raise too
    arr = (ctypes.c_int32 * max(len(cols), 1))(*cols) if cols else (ctypes.c_int32 * 1)(0)
    return arr, len(cols)


class TableEntry:
    """Immutable per-device snapshot of the prepared token table (what one fused VQ call needs)."""
    __slots__ = ("key", "table", "hat", "hat_t", "norm", "mean", "V", "D", "Vp")

    def __init__(self, key, table, hat, hat_t, norm, mean, V, D, Vp):
        self.key, self.table, self.hat, self.hat_t, self.norm, self.mean = key, table, hat, hat_t, norm, mean
        self.V, self.D, self.Vp = V, D, Vp


class TokenTableCache:
    """fp16 unit-norm copy of the frozen CLIP token table (+ transpose, norms, mean), rebuilt only when the table
    tensor changes (data pointer / version counter / shape).  The table is frozen in the reference
    (kw_branches.py:194 asserts requires_grad == False), so in steady state this costs nothing per step.

    One entry PER DEVICE behind a lock: ``nn.DataParallel`` replicas are shallow copies that share this object and call
    ``get`` concurrently from one thread per GPU (SURVEY.md section 8(b) "Threading"); ``get`` returns an immutable
    :class:`TableEntry`, so a replica can never observe another device's pointers."""

    def __init__(self):
        self._entries = {}
        self._lock = threading.Lock()

    def get(self, table: torch.Tensor) -> TableEntry:
        _lib.require_cuda(table, "token table")
        key = (table.data_ptr(), table._version, tuple(table.shape), table.dtype, table.device)
        with self._lock:
            entry = self._entries.get(table.device)
            if entry is not None and entry.key == key:
                return entry
            lib = _lib.load()
            src = table.detach()
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.float().contiguous()
            V, D = src.shape
            Vp = int(lib.scp_vq_padded_vocab(V))
            dev = src.device
            hat = torch.empty((Vp, D), dtype=torch.float16, device=dev)
            hat_t = torch.empty((D, Vp), dtype=torch.float16, device=dev)
            norm = torch.empty(Vp, dtype=torch.float32, device=dev)
            mean = torch.empty(D + 1, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                st = lib.scp_vq_prepare_table(_lib.ptr(src), V, D, _lib.ptr(hat), _lib.ptr(hat_t), _lib.ptr(norm),
                                              _lib.ptr(mean), _lib.stream_ptr(dev))
            _lib.check(st, "scp_vq_prepare_table")
            entry = TableEntry(key, src, hat, hat_t, norm, mean, V, D, Vp)
            self._entries[table.device] = entry
            return entry


class _FusedVQFn(torch.autograd.Function):
    """keywords_out, idx, metrics, avg_probs, code_hist = f(keywords_in, tau)."""

    @staticmethod
    def forward(ctx, kw: torch.Tensor, tau: torch.Tensor, cache: TableEntry, prob_msk, training: bool,
                want_avg_probs: bool, save_probs: bool = False):
        lib = _lib.load()
        B, K, D = kw.shape
        M = B * K
        V, Vp = cache.V, cache.Vp
        assert D == cache.D, (D, cache.D)
        dev = kw.device
        kw2 = kw.detach().reshape(M, D)
        if kw2.dtype != torch.float32 or not kw2.is_contiguous():
            kw2 = kw2.float().contiguous()
        tau_f = tau.detach().reshape(-1)[:1].float().contiguous()
        masked, n_masked = _masked_array(prob_msk, V)
        Mp = (M + 127) // 128 * 128
        idx = torch.empty(M, dtype=torch.int64, device=dev)
        kw_out = torch.empty((M, D), dtype=torch.float32, device=dev)
        row_stats = torch.empty((M, 4), dtype=torch.float32, device=dev)
        code_hist = torch.empty(Vp, dtype=torch.float32, device=dev)
        avg_probs = torch.empty(Vp, dtype=torch.float32, device=dev) if want_avg_probs else None
        metrics = torch.empty(3 + K, dtype=torch.float32, device=dev)
        kw_hat = torch.empty((Mp, D), dtype=torch.float16, device=dev)
        # the soft-max numerators the backward pass would otherwise recompute take the place of the forward's (M,V) scratch;
        # they are owned by this call (never a shared workspace: a second forward before the backward must not overwrite them)
        saved = None
        if training and save_probs and ctx.needs_input_grad[0] and lib.scp_vq_bwd_saved_available(M, V, D):
            saved = torch.empty(lib.scp_vq_saved_probs_bytes(M, V), dtype=torch.uint8, device=dev)
        ws_bytes = (lib.scp_vq_fwd_save_workspace_bytes if saved is not None else lib.scp_vq_fwd_workspace_bytes)(M, V, D)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = lib.scp_vq_fwd_save(_lib.ptr(kw2), M, K, V, D, _lib.ptr(cache.hat), _lib.ptr(cache.norm),
                                     _lib.ptr(cache.table), masked, n_masked, _lib.ptr(tau_f), _lib.ptr(idx),
                                     _lib.ptr(kw_out), _lib.ptr(row_stats), _lib.ptr(code_hist), _lib.ptr(avg_probs),
                                     _lib.ptr(metrics), _lib.ptr(kw_hat), _lib.ptr(saved), _lib.ptr(ws), ws_bytes,
                                     _lib.stream_ptr(dev))
        _lib.check(st, "scp_vq_fwd_save")
        ctx.training = training
        ctx.set_materialize_grads(False)  # five of the six outputs carry no gradient: do not zero-fill them per step
        if training:
            ctx.saved_probs = saved
            ctx.save_for_backward(kw2, kw_hat, row_stats, tau_f)
            ctx.cache = cache
            ctx.prob_msk = list(prob_msk)
            ctx.shape = (B, K, D)
            ctx.in_dtype = kw.dtype
        out = kw_out.view(B, K, D)
        if avg_probs is None:
            avg_probs = torch.empty(0, device=dev)
        if training:
            ctx.mark_non_differentiable(idx, metrics, row_stats, code_hist, avg_probs)
        else:  # eval: subword_prob is the one-hot (my_vector_quantizer.py:138-139): no gradient path at all
            ctx.mark_non_differentiable(out, idx, metrics, row_stats, code_hist, avg_probs)
        return out, idx, metrics, row_stats, code_hist, avg_probs

    @staticmethod
    def backward(ctx, g_out, *unused):
        if not ctx.training:  # eval: subword_prob = hard one-hot, no gradient path (my_vector_quantizer.py:138-139)
            return None, None, None, None, None, None, None
        lib = _lib.load()
        kw2, kw_hat, row_stats, tau_f = ctx.saved_tensors
        cache = ctx.cache
        B, K, D = ctx.shape
        if g_out is None:
            return None, None, None, None, None, None, None
        M = B * K
        dev = kw2.device
        g = g_out.reshape(M, D).float().contiguous()
        masked, n_masked = _masked_array(ctx.prob_msk, cache.V)
        g_kw = torch.empty((M, D), dtype=torch.float32, device=dev)
        need_tau = ctx.needs_input_grad[1]
        g_tau = torch.empty(1, dtype=torch.float32, device=dev) if need_tau else None
        saved = None if need_tau else ctx.saved_probs
        ws_bytes = (lib.scp_vq_bwd_saved_workspace_bytes(M, cache.V, D, int(need_tau)) if saved is not None
                    else lib.scp_vq_bwd_workspace_bytes(M, cache.V, D))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = lib.scp_vq_bwd_saved(_lib.ptr(g), _lib.ptr(kw2), M, cache.V, D, _lib.ptr(kw_hat), _lib.ptr(cache.hat),
                                      _lib.ptr(cache.hat_t), _lib.ptr(cache.norm), _lib.ptr(cache.mean),
                                      _lib.ptr(row_stats), masked, n_masked, _lib.ptr(tau_f),
                                      _lib.ptr(saved), _lib.ptr(g_kw),
                                      _lib.ptr(g_tau), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev))
        _lib.check(st, "scp_vq_bwd_saved")
        return g_kw.view(B, K, D).to(ctx.in_dtype), g_tau, None, None, None, None, None


class _DenseVQFn(torch.autograd.Function):
    """subword_prob, idx, metrics = f(x, tau) for a dense score tensor x (B,K,V); masks x in place."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, tau: torch.Tensor, prob_msk, training: bool, hard: bool = True):
        lib = _lib.load()
        B, K, V = x.shape
        M = B * K
        dev = x.device
        x2 = x.detach().view(M, V)  # the reference masks the caller's tensor in place (:78-79): so do we
        tau_f = tau.detach().reshape(-1)[:1].float().contiguous()
        masked, n_masked = _masked_array(prob_msk, V)
        idx = torch.empty(M, dtype=torch.int64, device=dev)
        row_stats = torch.empty((M, 4), dtype=torch.float32, device=dev)
        code_hist = torch.empty(V, dtype=torch.float32, device=dev)
        avg_probs = torch.empty(V, dtype=torch.float32, device=dev)
        metrics = torch.empty(3 + K, dtype=torch.float32, device=dev)
        sub = torch.empty((M, V), dtype=torch.float32, device=dev)
        ws_bytes = lib.scp_vq_dense_workspace_bytes(M, V)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = lib.scp_vq_dense_fwd(_lib.ptr(x2), M, K, V, x2.stride(0), masked, n_masked, _lib.ptr(tau_f),
                                      int(training) | (0 if hard else 2), _lib.ptr(idx), _lib.ptr(row_stats), _lib.ptr(code_hist),
                                      _lib.ptr(avg_probs), _lib.ptr(metrics), _lib.ptr(sub), _lib.ptr(ws), ws_bytes,
                                      _lib.stream_ptr(dev))
        _lib.check(st, "scp_vq_dense_fwd")
        ctx.training = training
        if training:
            ctx.save_for_backward(x2, row_stats, tau_f)
            ctx.shape = (B, K, V)
        ctx.mark_non_differentiable(idx, metrics)
        return sub.view(B, K, V), idx, metrics

    @staticmethod
    def backward(ctx, g_sub, *unused):
        if not ctx.training:
            return None, None, None, None, None
        lib = _lib.load()
        x2, row_stats, tau_f = ctx.saved_tensors
        B, K, V = ctx.shape
        M = B * K
        dev = x2.device
        g = g_sub.reshape(M, V).float().contiguous()
        g_x = torch.empty((M, V), dtype=torch.float32, device=dev)
        need_tau = ctx.needs_input_grad[1]
        g_tau = torch.empty(1, dtype=torch.float32, device=dev) if need_tau else None
        with torch.cuda.device(dev):
            st = lib.scp_vq_dense_bwd(_lib.ptr(x2), _lib.ptr(g), M, V, x2.stride(0), V, _lib.ptr(row_stats),
                                      _lib.ptr(tau_f), _lib.ptr(g_x), _lib.ptr(g_tau), _lib.stream_ptr(dev))
        _lib.check(st, "scp_vq_dense_bwd")
        return g_x.view(B, K, V), g_tau, None, None, None


class SimpleVectorQuantizer(nn.Module):
    """SimpleVectorQuantizer (my_vector_quantizer.py:12-62 constructor semantics)."""

    def __init__(self, temp, groundTruthPerplexity=None, time_first=True, use_gumbel=False, hard=True):
        super().__init__()
        self.time_first = time_first
        self.use_gumbel = use_gumbel
        self.hard = hard
        if use_gumbel:
            raise NotImplementedError("use_gumbel=True is not used by any shipped recipe and has no CUDA path here")

        if isinstance(temp, str):
            if temp.startswith("learnable="):
                self.temp_type = "learnable"
                value = ast.literal_eval(temp.replace("learnable=", ""))
                self.curr_temp = nn.parameter.Parameter(torch.FloatTensor([value]))
                logger.info("Setting vq temp learnable (init={})".format(value))
            elif temp.startswith("fixed="):
                self.temp_type = "fixed"
                value = ast.literal_eval(temp.replace("fixed=", ""))
                self.register_buffer("curr_temp", torch.FloatTensor([value]))
                logger.info("Setting vq temp fixed={}".format(value))
            else:
                self.temp_type = "scheduled"
                sched = ast.literal_eval(temp)
                assert len(sched) == 3, f"{sched}, {len(sched)}"
                self.max_temp, self.min_temp, self.temp_decay = sched
                logger.info("Setting vq temp scheduled = ({},{},{})".format(*sched))
                # the reference keeps a python float here and then crashes on `.item()` (:123); we keep a buffer that
                # set_num_updates refreshes so that the scheduled mode is actually usable
                self.register_buffer("curr_temp", torch.FloatTensor([self.max_temp]), persistent=False)
        else:
            raise TypeError("temp must be a string: 'learnable=x', 'fixed=x' or '(max,min,decay)'")
        self.codebook_indices = None
        self.groundTruthPerplexity = groundTruthPerplexity
        if self.groundTruthPerplexity is not None:
            self.perplexity_criteria = nn.MSELoss()
        self._table_cache = TokenTableCache()
        self._temp_float_cache = {}  # device -> (key, value); DataParallel replicas share this dict

    def set_num_updates(self, num_updates):
        if self.temp_type == "scheduled":  # :58-62
            value = max(self.max_temp * self.temp_decay ** num_updates, self.min_temp)
            self.curr_temp.fill_(value)

    # -- helpers ---------------------------------------------------------------------------------------------
    def _temp_as_float(self) -> float:
        """``result["temp"]`` is a python float in the reference (``.item()``, a device sync every step, :123).  For a
        fixed / scheduled temperature the value is cached per tensor version, so no sync happens in steady state."""
        t = self.curr_temp
        key = (t.data_ptr(), t._version)
        hit = self._temp_float_cache.get(t.device)
        if self.temp_type == "learnable" or hit is None or hit[0] != key:
            hit = (key, float(t.detach().reshape(-1)[0].item()))
            self._temp_float_cache[t.device] = hit  # one atomic dict store: safe across replica threads
        return hit[1]

    def _finish(self, result: Dict, metrics: torch.Tensor, K: int) -> Dict:
        result["code_perplexity"] = metrics[0]
        result["ent_per_t"] = metrics[3:3 + K]
        result["prob_perplexity"] = metrics[1]
        result["temp"] = self._temp_as_float()
        if self.groundTruthPerplexity is not None:
            gt = torch.tensor(float(self.groundTruthPerplexity), device=metrics.device)
            result["diversity_loss"] = self.perplexity_criteria(metrics[1], gt) / (
                result["num_vars"] - self.groundTruthPerplexity) ** 2
        else:
            result["diversity_loss"] = metrics[2]
        return result

    # -- reference signature: dense scores ----------------------------------------------------------------------
    def forward(self, x, prob_msk=[0, 2, 3], produce_targets=True):
        _lib.require_cuda(x, "SimpleVectorQuantizer")
        if not self.time_first:
            x = x.transpose(1, 2)
        bsz, tsz, fsz = x.shape
        result = {"num_vars": fsz}
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        sub, idx, metrics = _DenseVQFn.apply(x, self.curr_temp, tuple(prob_msk), self.training, self.hard)
        self._finish(result, metrics, tsz)
        result["subword_prob"] = sub
        if produce_targets:
            result["targets"] = idx.view(bsz, tsz, 1)
        return result

    # -- fused hot path ---------------------------------------------------------------------------------------
    def quantize_keywords(self, keywords: torch.Tensor, table: torch.Tensor, prob_msk=(0, 2, 3),
                          produce_targets: bool = True, compute_prob_perplexity: bool = True
                          ) -> Tuple[Dict, torch.Tensor]:
        """Fused V1+V3+V4 of the reference: cosine(keywords, table) -> mask -> arg-max / softmax / straight-through ->
        ``subword_prob @ table``.  Returns ``(vq_results, keywords_out)`` like ``GeneralBranch.vq_audio_features``.
        ``vq_results["subword_prob"]`` is ``None``: its only consumer in the reference is the lookup matmul
        (kw_branches.py:195), which is fused here."""
        _lib.require_cuda(keywords, "quantize_keywords")
        if not self.hard:
            raise NotImplementedError("hard=False: the soft lookup softmax(x / tau) @ E has no fused kernel (no shipped recipe "
                                      "uses it); call forward() on the dense cosine scores -- fused_vq_audio_features does")
        cache = self._table_cache.get(table)
        B, K, _ = keywords.shape
        # fixed / scheduled temperature >= 0.1 (every shipped recipe: 0.1): the forward keeps the fp16 soft-max numerators
        # and the backward skips the k . E^T product and the exponentials (scp_vq_fwd_save / scp_vq_bwd_saved).  A learnable
        # temperature needs the logits for d/dtau; below 0.1 the numerators exp((c - 1)/tau + 10) of strongly negative
        # cosines (c < 1 - 26.6 tau) underflow fp16 and avg_probs, which is derived from them, would lose those columns
        # (tests/test_vq_saved_math.py): recompute path.
        save_probs = (self.training and self.temp_type != "learnable" and torch.is_grad_enabled()
                      and keywords.requires_grad and self._temp_as_float() >= 0.1 - 1e-6)
        out, idx, metrics, row_stats, code_hist, avg_probs = _FusedVQFn.apply(
            keywords, self.curr_temp, cache, tuple(prob_msk), self.training, compute_prob_perplexity, save_probs)
        result = {"num_vars": cache.V}
        self._finish(result, metrics, K)
        result["subword_prob"] = None
        if produce_targets:
            result["targets"] = idx.view(B, K, 1)
        result["row_stats"] = row_stats
        result["avg_probs"] = avg_probs[:cache.V] if avg_probs.numel() else None
        result["code_hist"] = code_hist[:cache.V]
        return result, out


def fused_vq_audio_features(branch, audio_feat: torch.Tensor) -> Tuple[dict, torch.Tensor]:
    """Replacement body for ``GeneralBranch.vq_audio_features`` (avssl/model/kw_branches.py:181-197): same inputs,
    same ``(vq_results, keywords)`` outputs; the projection / batch-norm prologue stays the branch's own."""
    audio_feat = branch.project_feats_to_CLIPspace(audio_feat)
    table = branch.clip.model.token_embedding.weight
    assert table.requires_grad == False  # noqa: E712  (kw_branches.py:194)
    vq = branch.vector_quantizer
    if not getattr(vq, "hard", True):
        # hard=False (no shipped recipe): the reference's own flow -- dense cosine scores (kw_branches.py:158-179), the dense
        # CUDA quantiser, the lookup matmul (:195)
        cos_score = branch.get_keyword_cosine_score(audio_feat)
        vq_results = vq(x=cos_score)
        return vq_results, vq_results["subword_prob"] @ table
    return vq.quantize_keywords(audio_feat, table)
