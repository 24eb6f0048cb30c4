"""Bring-up diagnostics for the CUDA kernels (run on the GPU box):  python tools/gpu_bringup.py [stage ...]

Every stage runs in its own subprocess with a timeout so that a faulting kernel (sticky CUDA error) cannot hide
the results of the others.  Prints max / normwise errors against the oracle (tests' checker) for each piece.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["wsum", "table", "vq_small", "vq_mid", "vq_bwd_small", "vq_bwd_mid", "nce_small", "nce_mid", "vq_full"]


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    d = b.abs().max().item()
    return (a - b).abs().max().item() / (d if d > 0 else 1.0)


def nrm(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    d = b.norm().item()
    return (a - b).norm().item() / (d if d > 0 else 1.0)


def stage_wsum():
    import torch
    from oracle import speechclip_oracle as oracle
    from speechclip_plus_b200 import WeightedSumLayer
    torch.manual_seed(0)
    for (L, B, T, D, norm, dt) in [(13, 4, 50, 768, False, torch.float32), (13, 4, 50, 768, True, torch.float32),
                                   (25, 2, 33, 1024, True, torch.float32), (13, 4, 50, 768, False, torch.float16),
                                   (13, 3, 17, 768, True, torch.bfloat16), (4, 3, 9, 192, True, torch.float32)]:
        storage = [torch.randn(T, B, D) * (1 + 0.3 * l) + 0.1 * l for l in range(L)]
        storage = [s.to(dt) for s in storage]
        w = torch.randn(L) * 0.5
        gy = torch.randn(B, T, D)
        ref_layers = [s.float().transpose(0, 1).clone().requires_grad_(True) for s in storage]
        wr = w.clone().requires_grad_(True)
        y_ref = oracle.wsum_forward(ref_layers, wr, norm)
        g_ref = torch.autograd.grad(y_ref, [wr] + ref_layers, grad_outputs=gy)
        layer = WeightedSumLayer(L, norm).cuda()
        with torch.no_grad():
            layer.weights.copy_(w)
        dev_layers = [s.cuda().transpose(0, 1).requires_grad_(True) for s in storage]
        y = layer(dev_layers)
        g = torch.autograd.grad(y, [layer.weights] + dev_layers, grad_outputs=gy.cuda())
        print(f"wsum L={L} B={B} T={T} D={D} norm={norm} {dt}: y rel {rel(y, y_ref):.2e}  dW rel {rel(g[0], g_ref[0]):.2e}"
              f"  dX nrm {nrm(torch.stack([x.float() for x in g[1:]]), torch.stack(list(g_ref[1:]))):.2e}", flush=True)


def _mk_vq(B, K, V, D, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    table = torch.randn(V, D, generator=g) * 0.02 + 0.003 * torch.randn(1, D, generator=g)
    kw = torch.randn(B, K, D, generator=g) * table.std(0) + table.mean(0)
    gout = torch.randn(B, K, D, generator=g)
    return table, kw, gout


def stage_table():
    import torch
    from speechclip_plus_b200 import TokenTableCache
    for (V, D) in [(512, 64), (8112, 512), (19787, 768)]:
        table, _, _ = _mk_vq(1, 1, V, D)
        c = TokenTableCache().get(table.cuda())
        torch.cuda.synchronize()
        n = table.norm(dim=1).clamp_min(1e-8)
        hat = table / n[:, None]
        print(f"table V={V} D={D} Vp={c.Vp}: hat rel {rel(c.hat[:V].float(), hat):.2e} pad0 {c.hat[V:].abs().max().item() if c.Vp > V else 0:.1e}"
              f" hat_t rel {rel(c.hat_t[:, :V].float().t(), hat):.2e} norm rel {rel(c.norm[:V], n):.2e}"
              f" mean rel {rel(c.mean[:D], table.mean(0)):.2e} norm_ref {c.mean[D].item():.5f} vs {n.max().item():.5f}", flush=True)


def _vq_check(B, K, V, D, bwd, seed=0, tau=0.1):
    import torch
    from oracle import speechclip_oracle as oracle
    from speechclip_plus_b200 import SimpleVectorQuantizer
    table, kw, gout = _mk_vq(B, K, V, D, seed)
    vq = SimpleVectorQuantizer(f"fixed={tau}").cuda().train()
    kw_d = kw.cuda().requires_grad_(True)
    t0 = time.time()
    res, out = vq.quantize_keywords(kw_d, table.cuda())
    torch.cuda.synchronize()
    t1 = time.time()
    # oracle in fp64 (exact ordering) and fp32
    ref64, out64 = oracle.vq_audio_features(kw.double(), table.double(), torch.tensor([tau], dtype=torch.float64), training=True)
    idx = res["targets"].flatten().cpu()
    idx64 = ref64["targets"].flatten()
    mism = (idx != idx64).nonzero().flatten()
    cos64 = ref64["masked_scores"].reshape(B * K, V)
    gaps = [(cos64[m, idx64[m]] - cos64[m, idx[m]]).item() for m in mism.tolist()]
    print(f"vq B={B} K={K} V={V} D={D}: fwd {1e3 * (t1 - t0):.1f} ms  idx mismatches {len(mism)}/{B * K} gaps {gaps[:5]}", flush=True)
    print(f"   keywords rel {rel(out, out64):.2e}  code_ppl rel {rel(res['code_perplexity'], ref64['code_perplexity']):.2e}"
          f"  prob_ppl rel {rel(res['prob_perplexity'], ref64['prob_perplexity']):.2e}"
          f"  ent_per_t rel {rel(res['ent_per_t'], ref64['ent_per_t']):.2e}"
          f"  div rel {rel(res['diversity_loss'], ref64['diversity_loss']):.2e}"
          f"  avg_probs rel {rel(res['avg_probs'], ref64['avg_probs']):.2e}", flush=True)
    x = cos64
    lse1 = torch.logsumexp(x, -1)
    lset = torch.logsumexp(x / tau, -1)
    rs = res["row_stats"].cpu().double()
    print(f"   lse1 abs {(rs[:, 0] - lse1).abs().max().item():.2e} lse_tau abs {(rs[:, 1] - lset).abs().max().item():.2e}"
          f" inv_norm rel {rel(rs[:, 3], 1.0 / kw.reshape(B * K, D).double().norm(dim=1)):.2e}", flush=True)
    if bwd:
        t0 = time.time()
        (gk,) = torch.autograd.grad(out, [kw_d], grad_outputs=gout.cuda())
        torch.cuda.synchronize()
        t1 = time.time()
        g_ref, _ = oracle.vq_keyword_grad(kw.double(), table.double(), torch.tensor(tau, dtype=torch.float64), gout.double())
        print(f"   bwd {1e3 * (t1 - t0):.1f} ms  g_kw nrm {nrm(gk, g_ref):.2e} rel {rel(gk, g_ref):.2e}", flush=True)


def stage_vq_small():
    _vq_check(3, 4, 512, 64, False)
    _vq_check(2, 8, 1024, 128, False, seed=1)
    _vq_check(5, 7, 1000, 64, False, seed=2)


def stage_vq_mid():
    _vq_check(32, 8, 8112, 512, False, seed=3)


def stage_vq_bwd_small():
    _vq_check(3, 4, 512, 64, True)
    _vq_check(2, 8, 1024, 128, True, seed=1)


def stage_vq_bwd_mid():
    _vq_check(32, 8, 8112, 512, True, seed=3)
    _vq_check(16, 8, 19787, 768, True, seed=4)


def stage_vq_full():
    _vq_check(64, 8, 49408, 512, True, seed=5)


def _nce_check(N, D, seed=0, **kw):
    import math
    import torch
    from oracle import speechclip_oracle as oracle
    from speechclip_plus_b200 import MaskedContrastiveLoss
    g = torch.Generator().manual_seed(seed)
    a = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(N, D, generator=g) + 0.5 * a, dim=-1)
    ids = torch.randint(0, max(N // 5, 3), (N,), generator=g)
    crit = MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True, **kw).cuda()
    ad = a.cuda().requires_grad_(True)
    bd = b.cuda().requires_grad_(True)
    loss = crit(ad, bd, ids.cuda())
    gr = torch.autograd.grad(loss, [ad, bd, crit.temperature])
    torch.cuda.synchronize()
    scale = math.exp(math.log(1 / 0.07))
    okw = dict(margin=kw.get("margin", 0.0), dcl=kw.get("dcl", False), a2b=kw.get("a2b", True), b2a=kw.get("b2a", True))
    l64 = oracle.nce_forward(a.double(), b.double(), ids, scale, **okw)
    da, db, dl = oracle.nce_grads(a.double(), b.double(), ids, scale, **okw)
    print(f"nce N={N} D={D} {kw}: loss {loss.item():.6f} ref {l64.item():.6f} abs {abs(loss.item() - l64.item()):.2e}"
          f"  dA nrm {nrm(gr[0], da):.2e} dB nrm {nrm(gr[1], db):.2e} dT rel {abs(gr[2].item() - dl.item()) / abs(dl.item()):.2e}", flush=True)


def stage_nce_small():
    _nce_check(8, 64)
    _nce_check(48, 64, seed=1)
    _nce_check(40, 128, seed=2, margin=0.2)
    _nce_check(40, 64, seed=3, dcl=True)
    _nce_check(40, 64, seed=4, b2a=False)


def stage_nce_mid():
    _nce_check(256, 512, seed=5)
    _nce_check(300, 768, seed=6)
    _nce_check(1024, 512, seed=7)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        globals()["stage_" + sys.argv[2]]()
        sys.exit(0)
    stages = sys.argv[1:] or STAGES
    for s in stages:
        print(f"===== {s}", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", s], timeout=240, cwd=ROOT)
            print(f"===== {s}: exit {r.returncode} in {time.time() - t0:.1f}s", flush=True)
        except subprocess.TimeoutExpired:
            print(f"===== {s}: TIMEOUT", flush=True)
