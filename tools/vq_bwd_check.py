"""Bring-up check of the VQ forward/backward against the fp64 oracle evaluated on the GPU, plus device times.

    SCP_VQ_BWD_PIPE={0,1,2} python tools/vq_bwd_check.py [--shapes small|full|all] [--time]

Prints one JSON line per shape: relative errors of the keyword gradient / outputs and the device time of fwd and bwd.
Test infrastructure only (imports oracle/).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import speechclip_plus_b200 as scp  # noqa: E402
from oracle import speechclip_oracle as oracle  # noqa: E402

SMALL = [(7, 5, 1000, 128), (33, 1, 600, 128), (40, 8, 4096, 256), (32, 8, 8112, 512), (3, 7, 12000, 512),
         (64, 8, 19787, 512), (1, 1, 300, 128)]
FULL = [(256, 8, 49408, 512), (32, 8, 49408, 512), (128, 12, 49408, 512), (128, 8, 49408, 256)]


def norm_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run(shape, tau, do_time, learnable):
    B, K, V, D = shape
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(V + B)
    table = torch.randn(V, D, device=dev, generator=gen) * 0.02 + 0.003 * torch.randn(1, D, device=dev, generator=gen)
    kw = torch.randn(B, K, D, device=dev, generator=gen) * table.std(0) + table.mean(0)
    # a few peaked rows: keyword = a table row (softmax at tau concentrates on it)
    kw.view(-1, D)[::5] = table[torch.randint(4, V, (len(kw.view(-1, D)[::5]),), device=dev, generator=gen)] * 2.0
    gout = torch.randn(B, K, D, device=dev, generator=gen)
    vq = scp.SimpleVectorQuantizer(f"learnable={tau}" if learnable else f"fixed={tau}").to(dev).train()
    kwd = kw.clone().requires_grad_(True)
    res, out = vq.quantize_keywords(kwd, table)
    grads = torch.autograd.grad(out, [kwd] + ([vq.curr_temp] if learnable else []), grad_outputs=gout)
    gk = grads[0]
    torch.cuda.synchronize()
    t64 = torch.tensor(tau, dtype=torch.float64, device=dev)
    g_ref, g_tau_ref = oracle.vq_keyword_grad(kw.double(), table.double(), t64, gout.double())
    rec = dict(shape=shape, tau=tau, mode=os.environ.get("SCP_VQ_BWD_PIPE", "1"), g_kw_err=norm_err(gk, g_ref))
    if learnable:
        rec["g_tau"] = float(grads[1].item())
        rec["g_tau_ref"] = float(g_tau_ref.item())
    # per-row worst error (peaked rows have tiny gradients: check them relative to the row's own reference norm + floor)
    gr = g_ref.view(-1, D)
    row_err = ((gk.view(-1, D).double() - gr).norm(dim=1) / (gr.norm(dim=1) + 1e-3 * gr.norm(dim=1).max())).max()
    rec["g_kw_row_err_max"] = float(row_err)
    if do_time:
        def timed(fn, iters=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
            for a, b in evs:
                a.record(); fn(); b.record()
            torch.cuda.synchronize()
            return statistics.median(a.elapsed_time(b) for a, b in evs)
        state = {}

        def f():
            state["o"] = vq.quantize_keywords(kwd, table)[1]

        def fb():
            f()
            torch.autograd.grad(state["o"], [kwd], grad_outputs=gout)
        tf = timed(f)
        tfb = timed(fb)
        rec["fwd_ms"] = tf
        rec["bwd_ms"] = tfb - tf
        M = B * K
        rec["bwd_frac_of_1657"] = 6.0 * M * V * D / ((tfb - tf) * 1e-3) / 1657.6e12
        rec["fwd_frac_of_1657"] = 2.0 * M * V * D / (tf * 1e-3) / 1657.6e12
    print(json.dumps(rec), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="small")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--learnable", action="store_true")
    args = ap.parse_args()
    shapes = dict(small=SMALL, full=FULL, all=SMALL + FULL)[args.shapes]
    worst = 0.0
    for shp in shapes:
        for tau in (0.1,):
            try:
                rec = run(shp, tau, args.time, args.learnable)
                worst = max(worst, rec["g_kw_err"])
            except Exception as exc:  # keep going: the point is to see every shape
                print(json.dumps(dict(shape=shp, error=repr(exc)[:400])), flush=True)
                worst = float("inf")
                torch.cuda.synchronize()
    print(json.dumps(dict(worst_g_kw_err=worst)))
    sys.exit(0 if worst < 1e-3 else 1)


if __name__ == "__main__":
    main()
