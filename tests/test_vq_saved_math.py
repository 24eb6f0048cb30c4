"""CPU: the algebra of the saved-numerator VQ backward (csrc/scp_vq.cu: Sweep1EpiT<true>, SweepTEpi, gemm_out,
vq_bwd_finalize) restated in torch and checked against the oracle's closed form of the reference gradient
(kw_branches.py:181-197 + my_vector_quantizer.py:130-136):

    P''[m,v] = exp((c[m,v] - 1)/tau + 10)              kept by the forward as fp16 (0 for masked columns)
    T'[m,v]  = (ghat_m . ehat_v) * |e_v| / norm_ref - s0_m,   s0_m = <ghat_m, mean(E)> / norm_ref,  norm_ref = max_v |e_v|
    Q''      = P'' * T'                                  written by sweep T as fp16
    U = Q'' Ehat,  W = P'' Ehat,  s = sum Q'' / sum P''
    g_khat = (U - s W) * |g| * norm_ref / (tau * sum P''),   g_kw = (g_khat - <g_khat, khat> khat) / |kw|

The row's own sum of P'' divides out the arbitrary scale of the numerators (no row normaliser is needed in the forward),
and the fp16 storage of P'' and Q'' stays within the 1e-3 gradient tolerance -- both are asserted here, in fp64 and with
the two fp16 roundings applied.
"""
import pytest
import torch

from conftest import norm_err
from oracle import speechclip_oracle as oracle


def _saved_numerator_grad(kw, table, g_out, tau, prob_msk, fp16_storage: bool):
    B, K, D = kw.shape
    k = kw.reshape(-1, D).double()
    g = g_out.reshape(-1, D).double()
    E = table.double()
    e_norm = E.norm(dim=-1).clamp_min(1e-8)
    e_hat = E / e_norm[:, None]
    k_norm = k.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    k_hat = k / k_norm
    c = k_hat @ e_hat.t()
    p = torch.exp((c - 1.0) / tau + 10.0)
    p[:, list(prob_msk)] = 0.0
    if fp16_storage:
        p = p.float().half().double()                       # cvt.rn.f16x2.f32 of the forward
    g_norm = g.norm(dim=-1, keepdim=True)
    g_hat = g / g_norm.clamp_min(1e-30)
    norm_ref = e_norm.max()
    s0 = (g_hat @ E.mean(dim=0)) / norm_ref
    t = (g_hat @ e_hat.t()) * (e_norm / norm_ref)[None, :] - s0[:, None]
    q = p * t
    if fp16_storage:
        q = q.float().half().double()                       # sweep T stores Q'' as fp16
    sum_p, sum_q = p.sum(-1, keepdim=True), (p * t).sum(-1, keepdim=True)   # the row sums are taken before the rounding
    u, w = q @ e_hat, p @ e_hat
    g_khat = (u - (sum_q / sum_p) * w) * g_norm * norm_ref / (tau * sum_p)
    g_kw = (g_khat - (g_khat * k_hat).sum(-1, keepdim=True) * k_hat) / k_norm
    return g_kw.reshape(B, K, D)


@pytest.mark.parametrize("tau", [0.1, 0.07, 0.5])
@pytest.mark.parametrize("planted", [False, True])
def test_saved_numerator_backward_equals_the_reference_gradient(tau, planted):
    gen = torch.Generator().manual_seed(int(tau * 1000) + planted)
    B, K, V, D = 5, 4, 700, 64
    table = torch.randn(V, D, generator=gen) * 0.02 + 0.003 * torch.randn(1, D, generator=gen)
    kw = torch.randn(B, K, D, generator=gen) * table.std(0) + table.mean(0)
    if planted:  # rows whose best cosine is ~1: P'' near the top of its range, strong cancellation in T' - s
        kw[:, 0] = table[torch.randint(4, V, (B,), generator=gen)] * 2.0 + 1e-3 * torch.randn(B, D, generator=gen)
    g_out = torch.randn(B, K, D, generator=gen)
    ref, _ = oracle.vq_keyword_grad(kw.double(), table.double(), torch.tensor(tau, dtype=torch.float64), g_out.double())
    exact = _saved_numerator_grad(kw, table, g_out, tau, (0, 2, 3), fp16_storage=False)
    assert norm_err(exact, ref) < 1e-9                      # the algebra: identical in fp64
    stored = _saved_numerator_grad(kw, table, g_out, tau, (0, 2, 3), fp16_storage=True)
    assert norm_err(stored, ref) < 1e-3                     # the two fp16 roundings stay inside the stated tolerance
    # the numerators fit fp16 for every cosine: exp(+-1/tau - 1/tau + 10) in [e^(10 - 2/tau), e^10]
    assert torch.exp(torch.tensor(10.0)).item() < 65504.0


@pytest.mark.parametrize("tau", [0.1, 0.25, 1.0, 0.07])
def test_column_sums_from_the_saved_numerators(tau):
    """vq_colsum_kernel<PMODE>: avg_probs[v] = 1/M sum_m 2^(tau lg2 P''[m,v] + (1 - 10 tau) log2 e - log2 Z_m) from the
    fp16 numerators equals mean_m softmax(c[m,:])[v] at temperature ONE (my_vector_quantizer.py:102) -- relative error of
    e^c = tau x the fp16 rounding of P''."""
    gen = torch.Generator().manual_seed(5 + int(tau * 100))
    M, V, D = 96, 900, 64
    table = torch.randn(V, D, generator=gen) * 0.02 + 0.003 * torch.randn(1, D, generator=gen)
    kw = torch.randn(M, D, generator=gen) * table.std(0) + table.mean(0)
    kw[:8] = table[10:18] * 1.5                                             # cosines ~1: numerators near e^10
    kw[8:16] = -table[20:28]                                                # cosines ~-1: numerators near e^(10 - 2/tau)
    c = oracle.cosine_scores(kw[None].double(), table.double())[0]
    msk = [0, 2, 3]
    cm = c.clone()
    cm[:, msk] = float("-inf")
    ref = torch.softmax(cm, dim=-1).mean(dim=0)
    p = torch.exp((c - 1.0) / tau + 10.0)
    p[:, msk] = 0.0
    p16 = p.float().half().double()
    log2e = 1.4426950408889634
    lz = torch.log2(torch.exp(cm).sum(-1, keepdim=True))                    # log2 Z_m from the unrounded statistics of sweep 1
    avg = torch.exp2(tau * torch.log2(p16) + (1.0 - 10.0 * tau) * log2e - lz).mean(dim=0)
    avg[msk] = 0.0
    rel = ((avg - ref).abs().max() / ref.max()).item()
    if tau >= 0.1:
        assert rel < 1e-3, rel
    else:
        # the reason the wrapper requires tau >= 0.1 for this path: cosines below 1 - 26.6 tau (here the planted rows at ~-1)
        # underflow fp16 and their columns drop out of avg_probs
        assert rel > 1e-3, rel
