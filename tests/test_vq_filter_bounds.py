"""CPU: soundness of the two numerical filters of the exact arg-max (csrc/scp_vq.cu), checked by simulation with the
constants parsed from the CUDA source, so that a change of either constant that breaks the guarantee fails here.

  1. chunk / group candidates: a column whose tensor-core logit (fp16 operands, fp32 accumulation) lies within
     kRescueMargin of the row's largest tensor-core logit is re-scored exactly.  The true arg-max must always be one of
     them: |c16 - c| <= 2^-10 for unit vectors, hence c16(argmax) >= max(c16) - 2 * 2^-10 > max(c16) - kRescueMargin.
  2. group filter from the fp16 e^c scratch: every column with c16 >= thr must satisfy
     fp16(ex2.approx(c16 * log2 e)) >= float32(exp(thr)) * (1 - slack), slack parsed from the source.
"""
import os
import re

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "speechclip_plus_b200", "csrc", "scp_vq.cu")).read()


def _const(pattern: str) -> float:
    m = re.search(pattern, SRC)
    assert m, pattern
    return float(m.group(1).rstrip("f"))


MARGIN = _const(r"constexpr float kRescueMargin = ([0-9.e+-]+f?);")
SLACK = _const(r"const float thr_e = expf\(thr\) \* \(1\.0f - ([0-9.e+-]+f?)\);")


def test_constants_parsed():
    assert 2 * 2.0 ** -10 < MARGIN < 1e-2
    assert 2.0 ** -11 < SLACK < 5e-3


def test_fp16_product_bound_keeps_the_true_argmax_among_the_candidates():
    """Unit vectors rounded to fp16 (what the tensor cores see), fp32 accumulation: the logit error stays below 2^-10,
    so the exact arg-max column is always within MARGIN of the largest approximate logit -- across dimensions, including
    adversarial near-ties (the runner-up within 1e-6 of the best)."""
    g = torch.Generator().manual_seed(7122)
    worst = 0.0
    for D in (64, 128, 512, 768, 1024):
        V, M = 4096, 64
        table = torch.nn.functional.normalize(torch.randn(V, D, generator=g, dtype=torch.float64) + 0.3, dim=-1)
        kw = torch.nn.functional.normalize(torch.randn(M, D, generator=g, dtype=torch.float64) + 0.3, dim=-1)
        kw[: M // 2] = torch.nn.functional.normalize(table[torch.randint(0, V, (M // 2,), generator=g)]
                                                     + 1e-3 * torch.randn(M // 2, D, generator=g, dtype=torch.float64), dim=-1)
        table[1] = table[0] + 1e-7 * torch.randn(D, generator=g, dtype=torch.float64)      # a near-duplicate pair
        kw[0] = table[0]
        exact = kw @ table.t()
        approx = (kw.half().float() @ table.half().float().t()).double()                     # fp16 operands, fp32 accumulate
        err = (approx - exact).abs().max().item()
        worst = max(worst, err)
        assert err <= 2.0 ** -10, (D, err)
        best = exact.argmax(dim=1)
        gap = approx.max(dim=1).values - approx[torch.arange(M), best]
        assert (gap < MARGIN).all(), (D, gap.max().item())
    assert 2 * worst < MARGIN


def test_e16_group_filter_never_drops_a_column_at_or_above_the_threshold():
    """fp16(e^c) against e^thr (1 - SLACK): exhaustive over a fine grid of row maxima and of columns at / just above the
    threshold, with the worst-case ex2.approx error (2^-22 relative, both signs) applied before the fp16 rounding."""
    mx = np.linspace(-1.0, 1.0, 4001, dtype=np.float64)[:, None]
    thr = (mx.astype(np.float32) - np.float32(MARGIN)).astype(np.float32)                 # float32, as in the kernel
    thr_e = (np.exp(thr.astype(np.float64)).astype(np.float32) * np.float32(1.0 - SLACK)).astype(np.float32)
    # columns from exactly the threshold up to the row maximum (dense just above the threshold, where the filter is tight)
    offs = np.concatenate([np.linspace(0.0, 1e-5, 101), np.linspace(1e-5, MARGIN, 400)])[None, :]
    c = thr.astype(np.float64) + offs
    for approx_err in (-2.0 ** -22, 0.0, 2.0 ** -22):
        e = np.exp(c) * (1.0 + approx_err)
        e16 = e.astype(np.float32).astype(np.float16).astype(np.float32)                 # cvt.rn.f16x2.f32
        assert (e16 >= thr_e).all(), approx_err
    # and it is a filter, not a pass-through: a column 2 * (MARGIN + slack) below the maximum is rejected
    far = np.exp(thr.astype(np.float64) - MARGIN - 2 * SLACK).astype(np.float32).astype(np.float16).astype(np.float32)
    assert (far < thr_e).all()


SLACK_P = _const(r"const float thr_p = expf\(\(thr - 1\.0f\) / \*tau_ptr \+ 10\.0f\) \* \(1\.0f - ([0-9.e+-]+f?)\);")
SUBNORMAL_GUARD = _const(r"const bool p_all = e16_is_p && thr_p < ([0-9.e+-]+f?);")


@torch.no_grad()
def test_saved_numerator_group_filter_never_drops_a_column_at_or_above_the_threshold():
    """scp_vq_fwd_save stores P'' = exp((c - 1)/tau + 10) instead of e^c; the arg-max group filter tests
    fp16(P'') >= float32(exp((thr - 1)/tau + 10)) * (1 - SLACK_P) and admits every group once that threshold falls below
    SUBNORMAL_GUARD (fp16 subnormals round absolutely, not relatively).  Simulated like the kernel computes it: tau = 0.1
    through the repeated-squaring chain ((e^c)^10 = ((x^2)^2 x)^2 in float32, every product rounded), any other tau
    through one ex2.approx; worst-case ex2.approx error (2^-22 relative, both signs) applied to every exponential."""
    assert 2.0 ** -11 < SLACK_P < 5e-3 and 6.2e-5 < SUBNORMAL_GUARD < 1e-3
    mx = np.linspace(-1.0, 1.0, 2001, dtype=np.float64)[:, None]
    thr = (mx.astype(np.float32) - np.float32(MARGIN)).astype(np.float32)
    offs = np.concatenate([np.linspace(0.0, 1e-5, 101), np.linspace(1e-5, MARGIN, 300)])[None, :]
    c = (thr.astype(np.float64) + offs).astype(np.float32)                                 # the tensor-core logit, float32
    f32 = np.float32
    for tau in (0.1, 0.07, 0.25, 1.0):
        thr_p = (np.exp((thr.astype(np.float64) - 1.0) / tau + 10.0).astype(f32) * f32(1.0 - SLACK_P)).astype(f32)
        checked = thr_p >= f32(SUBNORMAL_GUARD)                                            # below: every group is admitted
        assert checked.any()
        for approx_err in (-2.0 ** -22, 0.0, 2.0 ** -22):
            if tau == 0.1:
                e1 = (np.exp(c.astype(np.float64)) * (1.0 + approx_err)).astype(f32)        # ex2.approx(c * log2 e)
                x2 = (e1 * e1).astype(f32)
                x4 = (x2 * x2).astype(f32)
                x5 = (x4 * e1).astype(f32)
                p = (x5 * x5).astype(f32)                                                   # ipow2<10>
            else:
                p = (np.exp((c.astype(np.float64) - 1.0) / tau + 10.0) * (1.0 + approx_err)).astype(f32)
            p16 = p.astype(np.float16).astype(f32)                                          # cvt.rn.f16x2.f32
            ok = (p16 >= thr_p) | ~checked
            assert ok.all(), (tau, approx_err)
        # and it filters: two margins below the row maximum is rejected wherever the relative bound applies
        far = np.exp((thr.astype(np.float64) - MARGIN - 2 * SLACK_P * tau - 1.0) / tau + 10.0).astype(f32)
        far16 = far.astype(np.float16).astype(f32)
        assert ((far16 < thr_p) | ~checked).all(), tau
