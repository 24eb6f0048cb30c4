"""CPU oracle for the SpeechCLIP+ data-parallel hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-torch (CPU, fp32 or fp64) restatement of the reference algorithm for
the three hot-path subsystems.  It is NOT the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the timed CPU baseline.  Nothing under
``speechclip_plus_b200/`` imports it; the product path raises if the CUDA library is missing.

Parity pin: the reference ships no golden vectors or known-answer tests for this path
(its ``test/`` directory does not touch these classes), so the oracle is pinned
differentially: ``tests/golden/make_golden.py`` imports the reference's own modules from
``/root/reference`` in the build container, runs them on seeded inputs and commits
inputs + outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays every
fixture through this file.  Each function cites the reference lines it restates
(paths relative to the reference root).

Conventions: every function is pure (no module state), takes/returns torch tensors,
and works in the dtype of its inputs (call with ``.double()`` for an fp64 re-evaluation,
used by the tests to classify arg-max ties).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

__all__ = [
    "wsum_forward",
    "wsum_grad_weights",
    "normalize_hidden_states",
    "upstream_feat_len",
    "upstream_tail",
    "kw_batchnorm",
    "splice_keywords",
    "cif_integrate_and_fire",
    "keypadding_mask",
    "cosine_scores_loop",
    "cosine_scores",
    "vq_forward",
    "vq_audio_features",
    "vq_keyword_grad",
    "l2_normalise",
    "nce_mask",
    "nce_forward",
    "nce_grads",
    "hybrid_loss",
]


# ----------------------------------------------------------------------------------------
# S1  upstream-feature fusion            avssl/module/weighted_sum.py:26-45
# ----------------------------------------------------------------------------------------
def wsum_forward(layers: Sequence[torch.Tensor], weights: torch.Tensor,
                 normalize_features: bool = False) -> torch.Tensor:
    """y = sum_l softmax(weights)_l * X_l, with X_l optionally layer-normalised first.

    Restates weighted_sum.py:38-43: the softmax over the raw ``weights`` (:38), the stack of
    the L layer tensors (:40), the optional *non-affine* LayerNorm over the last dim applied
    to every layer BEFORE the sum (:41-42, eps = torch default 1e-5) and the weighted
    reduction over the layer axis (:43).
    """
    assert len(layers) == weights.numel(), (len(layers), weights.numel())  # :36
    w = torch.softmax(weights, dim=0)
    acc = None
    for w_l, x_l in zip(w, layers):
        if normalize_features:
            x_l = F.layer_norm(x_l, (x_l.shape[-1],))
        term = w_l * x_l
        acc = term if acc is None else acc + term
    return acc


def wsum_grad_weights(layers: Sequence[torch.Tensor], weights: torch.Tensor, grad_y: torch.Tensor,
                      normalize_features: bool = False) -> torch.Tensor:
    """Closed-form d(loss)/d(weights) for S1:  w * (d - <w, d>),  d_l = <grad_y, X_l>.

    (Derivative of weighted_sum.py:38-43; verified against autograd in the tests.)
    """
    w = torch.softmax(weights, dim=0)
    d = []
    for x_l in layers:
        if normalize_features:
            x_l = F.layer_norm(x_l, (x_l.shape[-1],))
        d.append((grad_y * x_l).sum())
    d = torch.stack(d)
    return w * (d - (w * d).sum())


# ----------------------------------------------------------------------------------------
# S1' caller tail of the HuBERT wrapper   avssl/module/speech_encoder_plus.py:572-622
# ----------------------------------------------------------------------------------------
def normalize_hidden_states(layers: Sequence[torch.Tensor], normalize_type: str) -> List[torch.Tensor]:
    """Per-layer rescale applied BEFORE the weighted sum when ``normalize_hiddenstates`` is set and
    ``normalize_type`` is "method1" / "method2" (speech_encoder_plus.py:574-592):
      method1  x / (||x||_2 + 1e-8), norm over the feature axis, per frame                    (:578-583)
      method2  x / mean_t(||x_t||_2) reshaped (-1,1,1): one scalar per utterance and layer    (:586-590)
    "s3prl" leaves the layers untouched here (it is the LayerNorm flag of the WeightedSumLayer, :472-476).
    """
    assert normalize_type in ("s3prl", "method1", "method2"), normalize_type  # :377
    out = []
    for x in layers:
        if normalize_type == "method1":
            x = x / (torch.norm(x, dim=-1, keepdim=True) + 1e-8)
        elif normalize_type == "method2":
            x = x / torch.mean(torch.norm(x, dim=-1), dim=-1).view(-1, 1, 1)
        out.append(x)
    return out


def upstream_feat_len(wav_len: Sequence[int], downsample_rate: int, max_frames: int) -> torch.Tensor:
    """feat_len = clamp_max(LongTensor([round(l / rate)]), T)  (speech_encoder_plus.py:604-611; Python ``round``
    = half to even)."""
    feat_len = torch.LongTensor([round(int(l) / downsample_rate) for l in wav_len])
    return torch.clamp_max(feat_len, max_frames)


def upstream_tail(layers: Sequence[torch.Tensor], weights: torch.Tensor, normalize_hiddenstates: bool,
                  normalize_type: str, wav_len: Sequence[int], downsample_rate: int = 320
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """speech_encoder_plus.py:572-622 with feat_select_idx == "weighted_sum": optional rescale, feat_len, weighted sum
    (whose LayerNorm flag is ``normalize_hiddenstates and normalize_type == "s3prl"``, :472-476)."""
    if normalize_hiddenstates and normalize_type.startswith("method"):
        layers = normalize_hidden_states(layers, normalize_type)
    feat_len = upstream_feat_len(wav_len, downsample_rate, layers[0].shape[1])
    y = wsum_forward(layers, weights, normalize_features=normalize_hiddenstates and normalize_type == "s3prl")
    return y, feat_len


# ----------------------------------------------------------------------------------------
# N1  keyword batch-norm prologue         avssl/module/speechclip_c_modules/kw_bn.py:97-164, :216-228
# ----------------------------------------------------------------------------------------
def kw_batchnorm(keywords: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, running_mean: torch.Tensor,
                 running_var: torch.Tensor, batchnorm_type: str = "same", parallel: bool = False,
                 training: bool = True, seq_lens: Optional[Sequence[int]] = None, momentum: float = 0.1,
                 eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Kw_BatchNorm.forward / Kw_BatchNorm_dynamic.forward written out as explicit statistics.

    keywords (B,K,D).  Parameter / running-statistic layouts follow the reference modules:
      "same" (and the dynamic layer, kw_bn.py:216-228): (D,)  -- one BatchNorm1d(D) over all B*K rows (:141-143), or
          over the first seq_lens[b] rows of every utterance with the other rows left untouched (:144-159);
      "eachKw", parallel=True: (D*K,) with feature index d*K + k -- BatchNorm1d(D*K) on the (B, D*K) view of the
          (B,D,K) permutation (:119-127);
      "eachKw", parallel=False: (K,D) -- K independent BatchNorm1d(D) layers, one per keyword slot (:128-140).
    Training uses the biased batch variance for the normalisation and updates the running statistics with the
    unbiased one (torch.nn.BatchNorm1d); eval uses the running statistics.  Returns (y, running_mean', running_var').
    """
    B, K, D = keywords.shape
    if batchnorm_type == "eachKw":
        if parallel:
            w, b = weight.view(D, K).t(), bias.view(D, K).t()            # (K,D) views of index d*K + k
            rm, rv = running_mean.view(D, K).t(), running_var.view(D, K).t()
        else:
            w, b, rm, rv = weight, bias, running_mean, running_var        # (K,D)
        x = keywords                                                       # statistics over the batch axis
        if training:
            mean = x.mean(dim=0)
            var = x.var(dim=0, unbiased=False)
            new_rm = (1 - momentum) * rm + momentum * mean.detach()
            new_rv = (1 - momentum) * rv + momentum * x.var(dim=0, unbiased=True).detach()
        else:
            mean, var, new_rm, new_rv = rm, rv, rm, rv
        y = (x - mean) / torch.sqrt(var + eps) * w + b
        if parallel:
            new_rm, new_rv = new_rm.t().reshape(-1), new_rv.t().reshape(-1)
        return y, new_rm, new_rv
    assert batchnorm_type == "same", batchnorm_type
    if seq_lens is None:
        valid = torch.ones(B, K, dtype=torch.bool)
    else:
        valid = torch.arange(K)[None, :] < torch.as_tensor(list(seq_lens))[:, None]
    rows = keywords[valid]                                                 # (n_valid, D)
    if training:
        mean = rows.mean(dim=0)
        var = rows.var(dim=0, unbiased=False)
        new_rm = (1 - momentum) * running_mean + momentum * mean.detach()
        new_rv = (1 - momentum) * running_var + momentum * rows.var(dim=0, unbiased=True).detach()
    else:
        mean, var, new_rm, new_rv = running_mean, running_var, running_mean, running_var
    normed = (keywords - mean) / torch.sqrt(var + eps) * weight + bias
    y = torch.where(valid[..., None], normed, keywords)
    return y, new_rm, new_rv


# ----------------------------------------------------------------------------------------
# V1  keyword-vs-vocabulary cosine        avssl/model/kw_branches.py:158-179
# ----------------------------------------------------------------------------------------
def cosine_scores_loop(keywords: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """Faithful (slow) form of get_keyword_cosine_score: one F.cosine_similarity call per
    keyword position on a (B, D, 1) x (1, D, V) broadcast (kw_branches.py:167-177), stacked
    on dim 1.  This is what the CPU baseline times; it materialises a B x D x V temporary per
    keyword exactly like the reference.
    """
    bsz, n_kw, dim = keywords.shape
    table_t = table.transpose(0, 1).unsqueeze(0)  # (1, D, V)
    per_kw = [F.cosine_similarity(keywords[:, i, :].reshape(bsz, dim, 1), table_t, dim=1)
              for i in range(n_kw)]
    return torch.stack(per_kw, dim=1)


def cosine_scores(keywords: torch.Tensor, table: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """Same numbers as :func:`cosine_scores_loop` as one normalise + GEMM.

    F.cosine_similarity divides each operand by max(||x||_2, eps) (eps = 1e-8) and then
    reduces the product; on (B,K,D) x (V,D) that is a row-normalised matmul.
    """
    k_hat = keywords / keywords.norm(dim=-1, keepdim=True).clamp_min(eps)
    e_hat = table / table.norm(dim=-1, keepdim=True).clamp_min(eps)
    return k_hat @ e_hat.t()


# ----------------------------------------------------------------------------------------
# V3  SimpleVectorQuantizer.forward       avssl/module/speechclip_c_modules/my_vector_quantizer.py:64-165
# ----------------------------------------------------------------------------------------
def vq_forward(cos: torch.Tensor, curr_temp: torch.Tensor | float, training: bool = True,
               prob_msk: Sequence[int] = (0, 2, 3), hard: bool = True,
               ground_truth_perplexity: Optional[float] = None) -> Dict[str, object]:
    """Vector-quantise cosine scores ``cos`` of shape (B, K, V).

    Follows my_vector_quantizer.py line by line (time_first=True, use_gumbel=False):
      * ``-inf`` added to the ``prob_msk`` columns (:78-79).  The reference does it IN PLACE on
        the caller's tensor; here a masked copy is made and returned as ``masked_scores``.
      * k = first arg-max over the vocabulary (:82), one-hot ``hard`` (:85-91)
      * code_perplexity = exp(-sum h log(h + 1e-7)), h = mean_m hard (:94-99)
      * avg_probs = mean_m softmax(x) at temperature ONE (:102); prob_perplexity likewise (:119-121)
      * ent_per_t[i] = mean_b( -sum_v p log(p + 1e-9) ), p = softmax(x) at temperature one (:104-116)
      * temp = curr_temp as a python float (:123)
      * train: p_tau = softmax(x / curr_temp); subword_prob = hard + p_tau - p_tau.detach() (:130-136)
        eval : subword_prob = hard (:138-139)
      * diversity_loss (:147-158), targets = argmax(subword_prob) (:160-163)
    """
    bsz, tsz, fsz = cos.shape
    x = cos.reshape(bsz * tsz, fsz).clone()
    for col in prob_msk:
        x[:, col] = x[:, col] + float("-inf")
    k = x.max(dim=-1).indices
    hard_x = torch.zeros_like(x).scatter_(-1, k.view(-1, 1), 1.0)

    hard_probs = hard_x.float().mean(dim=0)
    code_ppl = torch.exp(-(hard_probs * torch.log(hard_probs + 1e-7)).sum(dim=-1)).sum()

    p1 = torch.softmax(x.float() if x.dtype != torch.float64 else x, dim=-1)
    avg_probs = p1.mean(dim=0)
    p_bt = p1.view(bsz, tsz, fsz)
    ent_bt = -(p_bt * torch.log(p_bt + 1e-9)).sum(dim=-1)  # (B, K)
    ent_per_t = ent_bt.mean(dim=0)  # (K,)
    prob_ppl = torch.exp(-(avg_probs * torch.log(avg_probs + 1e-7)).sum(dim=-1)).sum()

    temp_val = float(curr_temp.item()) if torch.is_tensor(curr_temp) else float(curr_temp)
    if training:
        p_tau = torch.softmax(x / curr_temp, dim=-1).type_as(x)
        sub = hard_x + p_tau - p_tau.detach() if hard else p_tau
    else:
        sub = hard_x
    out: Dict[str, object] = {
        "num_vars": fsz,
        "code_perplexity": code_ppl,
        "ent_per_t": ent_per_t,
        "prob_perplexity": prob_ppl,
        "temp": temp_val,
        "subword_prob": sub.view(bsz, tsz, fsz),
        "masked_scores": x.view(bsz, tsz, fsz),
        "avg_probs": avg_probs,
        "hard_probs": hard_probs,
    }
    if ground_truth_perplexity is not None:
        gt = torch.tensor(float(ground_truth_perplexity)).type_as(x)
        out["diversity_loss"] = F.mse_loss(prob_ppl, gt) / (fsz - ground_truth_perplexity) ** 2
    else:
        out["diversity_loss"] = (fsz - prob_ppl) / fsz
    out["targets"] = sub.argmax(dim=-1).view(bsz, tsz, 1).detach()
    return out


# ----------------------------------------------------------------------------------------
# V4  GeneralBranch.vq_audio_features      avssl/model/kw_branches.py:181-197
#     (the Linear+BatchNorm prologue `project_feats_to_CLIPspace` :143-156 is out of scope; the
#      oracle starts from its output, the keyword vectors in CLIP space)
# ----------------------------------------------------------------------------------------
def vq_audio_features(keywords_in: torch.Tensor, table: torch.Tensor, curr_temp, training: bool = True,
                      prob_msk: Sequence[int] = (0, 2, 3), faithful_loop: bool = False, hard: bool = True
                      ) -> Tuple[Dict[str, object], torch.Tensor]:
    """cos = cosine(keywords, table) (:192); vq = quantiser(cos) (:193);
    keywords_out = subword_prob @ table (:195).  Returns (vq_results, keywords_out)."""
    cos = cosine_scores_loop(keywords_in, table) if faithful_loop else cosine_scores(keywords_in, table)
    vq = vq_forward(cos, curr_temp, training=training, prob_msk=prob_msk, hard=hard)
    kw_out = vq["subword_prob"] @ table
    return vq, kw_out


def vq_keyword_grad(keywords_in: torch.Tensor, table: torch.Tensor, curr_temp, grad_kw_out: torch.Tensor,
                    prob_msk: Sequence[int] = (0, 2, 3)) -> Tuple[torch.Tensor, torch.Tensor]:
    """Closed-form backward of V1+V3+V4 in training mode (straight-through estimator):

        g_p   = g_out E^T                      (through :195; value path `hard` has no grad)
        g_c   = p_tau * (g_p - <p_tau, g_p>) / tau            (softmax(x / tau), :130-131)
        g_khat= g_c Ehat ;  g_kw = (g_khat - <g_khat, khat> khat) / max(||k||, eps)
        g_tau = -sum_m <g_c[m], x[m]> / tau    (only meaningful for a learnable temperature)

    Returns (grad wrt keywords_in, grad wrt curr_temp).  Checked against autograd in tests.
    """
    B, K, D = keywords_in.shape
    kw = keywords_in.reshape(B * K, D)
    g_out = grad_kw_out.reshape(B * K, D)
    tau = curr_temp if torch.is_tensor(curr_temp) else torch.tensor(float(curr_temp), dtype=kw.dtype)
    tau = tau.reshape(()).to(kw.dtype)
    k_norm = kw.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    k_hat = kw / k_norm
    e_hat = table / table.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    x = k_hat @ e_hat.t()
    x[:, list(prob_msk)] = float("-inf")
    p_tau = torch.softmax(x / tau, dim=-1)
    g_p = g_out @ table.t()
    s = (p_tau * g_p).sum(dim=-1, keepdim=True)
    g_c = p_tau * (g_p - s) / tau
    g_khat = g_c @ e_hat
    g_kw = (g_khat - (g_khat * k_hat).sum(-1, keepdim=True) * k_hat) / k_norm
    x_fin = torch.where(torch.isinf(x), torch.zeros_like(x), x)
    g_tau = -(g_c * x_fin).sum() / tau
    return g_kw.view(B, K, D), g_tau


# ----------------------------------------------------------------------------------------
# N0  L2 normalisation of the loss features   avssl/model/kwClip.py:857, :905-907, :913-915
# ----------------------------------------------------------------------------------------
def l2_normalise(feat: torch.Tensor) -> torch.Tensor:
    """f / ||f||_2 over the last dim, no epsilon (kwClip.py:857)."""
    return feat / feat.norm(dim=-1, keepdim=True)


# ----------------------------------------------------------------------------------------
# N4  CIF integrate-and-fire              avssl/module/cif.py:157-311
# ----------------------------------------------------------------------------------------
def cif_integrate_and_fire(x: torch.Tensor, alpha: torch.Tensor, threshold: float = 1.0,
                           target_lengths: Optional[torch.Tensor] = None, apply_tail_handling: bool = True,
                           firing_threshold: float = 0.5, max_feat_len: int = 75
                           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """CIF.integrate_and_fire restated through an explicit (B,S,T+1) weight matrix ``W`` with ``out = W^T x``:
    source ``s`` spreads ``alpha_s`` over the output rows ``left_s .. right_s`` (rows = floor(cumsum/threshold), clipped
    to T, cif.py:193-201): ``right_w = csum - right*thr`` where it fires (:209-211), ``thr`` on each of the
    ``extra = max(fire_num-1,0)`` rows in between (:226-243), ``left_w = alpha - right_w - extra*thr`` (:219-221).  The
    indices carry no gradient (:194); the weights do.  Tail handling as cif.py:246-298.
    Returns (features (B,T',C), lengths (B,), fired marks (B,S))."""
    B, S, C = x.shape
    feat_len = (alpha.sum(1) / threshold).floor().clip(min=1, max=max_feat_len).long()   # :183-188
    T = int(feat_len.max())
    csum = alpha.cumsum(-1)
    with torch.no_grad():
        right = (csum / threshold).floor().long().clip(min=0, max=T)
        left = torch.cat([torch.zeros(B, 1, dtype=torch.long), right[:, :-1]], dim=1)
        fire = right - left
        extra = (fire - 1).clip(min=0)
    fire_mask = fire > 0
    right_w = torch.where(fire_mask, csum - right.to(alpha.dtype) * threshold, torch.zeros((), dtype=alpha.dtype))
    left_w = alpha - right_w - extra.to(alpha.dtype) * threshold
    W = torch.zeros(B, S, T + 1, dtype=x.dtype)
    W = W.scatter_add(2, right.unsqueeze(-1), right_w.to(x.dtype).unsqueeze(-1))
    W = W.scatter_add(2, left.unsqueeze(-1), left_w.to(x.dtype).unsqueeze(-1))
    for e in range(1, int(extra.max()) + 1):
        idx = (left + e).clip(max=T)
        W = W.scatter_add(2, idx.unsqueeze(-1), (threshold * (extra >= e)).to(x.dtype).unsqueeze(-1))
    out = torch.einsum("bst,bsc->btc", W, x)
    if apply_tail_handling and target_lengths is None:
        fl = feat_len.unsqueeze(1)
        tail_w = (torch.where(right == fl, right_w, torch.zeros(())) + torch.where(left == fl, left_w, torch.zeros(()))).sum(-1)
        extend = tail_w >= firing_threshold
        scale = torch.ones(B, T + 1, dtype=out.dtype)
        scale[torch.arange(B), feat_len] = torch.where(extend, threshold / tail_w, torch.ones(())).to(out.dtype).detach()
        out = out * scale.unsqueeze(-1)
        if bool(extend.any()):
            new_len = feat_len + extend.long()
            fire_mask = fire_mask.clone()
            fire_mask[:, new_len - 1] = fire_mask[:, new_len - 1] + extend      # the (B,B) broadcast of :281-283
            feat_len = new_len.clip(max=max_feat_len)
        T = int(feat_len.max())
        out = out[:, :T]
        keep = torch.arange(T)[None, :] < feat_len[:, None]
        out = out * keep.unsqueeze(-1)
    else:
        out = out[:, :T]
    return out, feat_len, fire_mask


# ----------------------------------------------------------------------------------------
# N3  text-transformer input splice       avssl/module/clip_official.py:240-267, avssl/util/data_utils.py:6-22
# ----------------------------------------------------------------------------------------
def splice_keywords(keywords: torch.Tensor, keyword_num, table: torch.Tensor, pos_emb: torch.Tensor,
                    sot_token: int, eot_token: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """The prologue of ClipModel.encode_keywords: a zero-initialised (B,77) id tensor with SOT at 0 (:240-249) and EOT at
    keyword_num + 1 (:251-258), the token-embedding lookup (:260), the keywords written over positions 1..n (:262-267)
    and the positional-embedding add (:269).  Returns (x (B,L,D), eot positions (B,))."""
    B = keywords.shape[0]
    L = pos_emb.shape[0]
    text = torch.zeros(B, L, dtype=torch.long)
    text[:, 0] = sot_token
    if torch.is_tensor(keyword_num):
        index = keyword_num.long() + 1
        text = text.scatter(1, index.unsqueeze(1), eot_token)
    else:
        index = torch.full((B,), int(keyword_num) + 1, dtype=torch.long)
        text[:, keyword_num + 1] = eot_token
    x = table[text]
    rows = []
    for i in range(B):  # out-of-place form of the slice assignment so that autograd reaches `keywords`
        n = int(index[i]) - 1
        rows.append(torch.cat([x[i, :1], keywords[i, :n].to(x.dtype), x[i, n + 1:]], dim=0))
    x = torch.stack(rows) + pos_emb
    return x, index


def keypadding_mask(max_length: int, data_lens: torch.Tensor) -> torch.Tensor:
    """True marks padding (data_utils.py:17-20)."""
    return torch.arange(max_length)[None, :] >= data_lens.long()[:, None]


# ----------------------------------------------------------------------------------------
# S3  MaskedContrastiveLoss.forward          avssl/module/losses.py:185-245
# ----------------------------------------------------------------------------------------
def nce_mask(index: Optional[torch.Tensor], n: int, dcl: bool = False) -> torch.Tensor:
    """Boolean (n, n) mask of logits that enter the denominators (losses.py:202-216):
    different-id pairs, plus the diagonal unless ``dcl``.  With ``index=None`` only the
    off-diagonal (+ diagonal unless dcl)."""
    eye = torch.eye(n, dtype=torch.bool)
    if index is not None:
        idx = index.reshape(-1, 1)
        neg = idx != idx.t()
    else:
        neg = ~eye
    if not dcl:  # :213-214 -- the positive pair also sits in its own denominator
        neg = neg | eye
    return neg


def nce_forward(feat_a: torch.Tensor, feat_b: torch.Tensor, index: Optional[torch.Tensor], scale,
                margin: float = 0.0, dcl: bool = False, a2b: bool = True, b2a: bool = True) -> torch.Tensor:
    """Masked two-way InfoNCE.  ``scale`` is the multiplicative logit scale the reference calls
    ``temperature`` after :219-222, i.e. exp(param) when trainable or 1/temperature when fixed.

    logits = A B^T * scale (:224); diagonal -= margin (:227-228); exp WITHOUT max-subtraction,
    masked (:232); loss = mean_i(-S_ii + log sum_j) [a2b, :234-237] + the column version
    [b2a, :238-241], halved when both are on (:242-243).
    """
    n = feat_a.shape[0]
    assert feat_a.shape == feat_b.shape
    mask = nce_mask(index, n, dcl).to(feat_a.dtype)
    logits = feat_a @ feat_b.t() * scale
    if margin > 0.0:
        logits = logits - margin * torch.eye(n, dtype=logits.dtype)
    pos = torch.diagonal(logits)
    e = logits.exp() * mask
    loss = 0
    if a2b:
        loss = loss + (-pos + torch.log(e.sum(1))).mean()
    if b2a:
        loss = loss + (-pos + torch.log(e.sum(0))).mean()
    if a2b and b2a:
        loss = loss / 2
    return loss


def nce_grads(feat_a: torch.Tensor, feat_b: torch.Tensor, index: Optional[torch.Tensor], scale,
              margin: float = 0.0, dcl: bool = False, a2b: bool = True, b2a: bool = True
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Closed-form gradients of :func:`nce_forward`: (dA, dB, d log(scale)).

    G = c_r * mask*e^S / Z^r_i + c_c * mask*e^S / Z^c_j - (c_r + c_c) I   with c = 1/(n * #directions);
    dA = scale * G B;  dB = scale * G^T A;  d(log scale) = sum(G * S)  (S incl. margin shift on the
    value path only: the margin is a constant, so d/dlog(scale) uses scale * A B^T).
    """
    n = feat_a.shape[0]
    mask = nce_mask(index, n, dcl).to(feat_a.dtype)
    raw = feat_a @ feat_b.t() * scale
    logits = raw - margin * torch.eye(n, dtype=raw.dtype) if margin > 0.0 else raw
    e = logits.exp() * mask
    ndir = int(a2b) + int(b2a)
    c = 1.0 / (n * ndir)
    G = torch.zeros_like(e)
    if a2b:
        G = G + c * e / e.sum(1, keepdim=True)
    if b2a:
        G = G + c * e / e.sum(0, keepdim=True)
    G = G - (ndir * c) * torch.eye(n, dtype=e.dtype)
    dA = scale * (G @ feat_b)
    dB = scale * (G.t() @ feat_a)
    dlog = (G * raw).sum()
    return dA, dB, dlog


# ----------------------------------------------------------------------------------------
# C0  KWClip_GeneralTransformer.compute_loss   avssl/model/kwClip.py:999-1040
# ----------------------------------------------------------------------------------------
def hybrid_loss(loss_feats: Dict[str, torch.Tensor], scale, cascaded_weight: float, parallel_weight: float,
                margin: float = 0.0, dcl: bool = False, a2b: bool = True, b2a: bool = True,
                quantity_loss_weight: float = 0.0) -> Dict[str, torch.Tensor]:
    """loss = sum_br w_br * criterion(audio_br, image, id)  (+ w_q * L1(cif_quantity_out, cif_target_len)).

    Branch order and key names as kwClip.py:1015-1028; features are up-cast to fp32 (:1012, :1024).
    """
    assert {"id", "image_feat"}.issubset(loss_feats.keys())
    out: Dict[str, torch.Tensor] = {"loss": 0}
    image = loss_feats["image_feat"].float() if loss_feats["image_feat"].dtype != torch.float64 else loss_feats["image_feat"]
    for name, w in (("cascaded", cascaded_weight), ("parallel", parallel_weight)):
        if w > 0.0:
            a = loss_feats[f"{name}_audio_feat"]
            a = a.float() if a.dtype != torch.float64 else a
            out[f"{name[0]}_cl_loss"] = nce_forward(a, image, loss_feats["id"], scale, margin, dcl, a2b, b2a)
            out["loss"] = out["loss"] + w * out[f"{name[0]}_cl_loss"]
    if quantity_loss_weight > 0.0 and "cif_quantity_out" in loss_feats and "cif_target_len" in loss_feats:
        out["quantity_loss"] = F.l1_loss(loss_feats["cif_quantity_out"], loss_feats["cif_target_len"])
        out["loss"] = out["loss"] + quantity_loss_weight * out["quantity_loss"]
    return out
