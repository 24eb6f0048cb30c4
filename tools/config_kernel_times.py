"""Device times of the three hot-path kernel groups at the per-GPU sizes of EVERY BASELINE.json config (the bench line is
config 3 only): S1 weighted sum, S2 keyword VQ, S3 masked InfoNCE -- forward and backward, CUDA events on the launching
stream, median of 20, against the rooflines of SURVEY section 8(d) (HBM bytes for S1, 2*M*V*D / 6*M*V*D FLOP for S2,
2*N^2*D / 4*N^2*D FLOP per criterion call for S3).

    python tools/config_kernel_times.py > gpurun_out/config_kernels.json
"""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import speechclip_plus_b200 as scp

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(peaks.get("hbm_gbs", 6650.0))
TC = float(peaks.get("bf16_tflops_burst", peaks.get("bf16_tflops", 1657.6)))
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7122)


def timed(fn, n=20, graph=False):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    if graph:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        fn = gr.replay
        fn()
        torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e-3


rows = []


def s1(cfg, L, B, T, D, norm):
    storage = [torch.randn(T, B, D, device=dev, generator=g) for _ in range(L)]
    layers = [s.transpose(0, 1) for s in storage]
    gy = torch.randn(B, T, D, device=dev, generator=g)
    layer = scp.WeightedSumLayer(L, normalize_features=norm).to(dev)
    t_f = timed(lambda: layer(layers))
    def fb():
        torch.autograd.grad(layer(layers), [layer.weights], grad_outputs=gy)
    t_b = timed(fb) - t_f
    by = (L + 1) * B * T * D * 4
    for name, t in (("S1 fwd", t_f), ("S1 bwd (weights)", t_b)):
        rows.append(dict(config=cfg, kernel=name + (" +LayerNorm" if norm else ""), shape=f"L={L} B={B} T={T} D={D}", ms=t * 1e3,
                         bound="hbm", achieved=by / t / 1e9, unit="GB/s", frac=by / t / 1e9 / HBM))
    del storage, layers


def s2(cfg, M, K, V, D):
    table = torch.randn(V, D, device=dev, generator=g) * 0.02
    kw = (torch.randn(M // K, K, D, device=dev, generator=g) * 0.02).requires_grad_(True)
    gk = torch.randn(M // K, K, D, device=dev, generator=g)
    vq = scp.SimpleVectorQuantizer("fixed=0.1").to(dev).train()
    t_f = timed(lambda: vq.quantize_keywords(kw, table))
    def fb():
        _, out = vq.quantize_keywords(kw, table)
        torch.autograd.grad(out, [kw], grad_outputs=gk)
    t_b = timed(fb) - t_f
    for name, t, fl in (("S2 VQ fwd", t_f, 2.0 * M * V * D), ("S2 VQ bwd", t_b, 6.0 * M * V * D)):
        rows.append(dict(config=cfg, kernel=name, shape=f"M={M} V={V} D={D}", ms=t * 1e3, bound="tensor",
                         achieved=fl / t / 1e12, unit="TFLOP/s", frac=fl / t / 1e12 / TC))


def s3(cfg, N, D, n_local, calls):
    img = torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=g), dim=-1)
    auds = [torch.nn.functional.normalize(torch.randn(N, D, device=dev, generator=g) + img, dim=-1).requires_grad_(True)
            for _ in range(calls)]
    ids = torch.randint(0, N // 5, (N,), device=dev, generator=g)
    crit = scp.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True).to(dev)
    def fb():
        loss = sum(crit(a, img, ids, local_rows=(0, n_local)) for a in auds)
        torch.autograd.grad(loss, auds + [crit.temperature])
    t = timed(fb, graph=True)
    fl = calls * 6.0 * N * N * D
    rows.append(dict(config=cfg, kernel=f"S3 InfoNCE fwd+bwd x{calls} (graph replay)", shape=f"N={N} D={D} local rows={n_local}",
                     ms=t * 1e3, bound="tensor (latency-bound in practice)", achieved=fl / t / 1e12, unit="TFLOP/s",
                     frac=fl / t / 1e12 / TC))


V = 49408
# config 2: parallel base, one GPU, batch 256 (S1 + S3 only)
s1("c2 parallel base, 1 GPU", 13, 256, 249, 768, False); s3("c2 parallel base, 1 GPU", 256, 512, 256, 1)
# config 3: cascaded+ base, batch 256 over G GPUs: M = 2048 / G
for G in (1, 8):
    s2(f"c3 cascaded+ base, {G} GPU(s)", 2048 // G, 8, V, 512)
s3("c3 cascaded+ base", 256, 512, 256, 1)
# config 4: hybrid+ base, global batch 1024 over 8 GPUs (per GPU 128 pairs), both criterion calls
s1("c4 hybrid+ base, 8 GPUs", 13, 128, 249, 768, False); s2("c4 hybrid+ base, 8 GPUs", 1024, 8, V, 512)
s3("c4 hybrid+ base, 8 GPUs", 1024, 512, 128, 2)
# config 5: hybrid+ large (25 layers x 1024, CLIP 768), global batch 512 over 4 / 8 GPUs; the large non-plus recipes add LayerNorm
for G in (4, 8):
    B = 512 // G
    s1(f"c5 hybrid+ large, {G} GPUs", 25, B, 249, 1024, False)
    s2(f"c5 hybrid+ large, {G} GPUs", B * 8, 8, V, 768)
    s3(f"c5 hybrid+ large, {G} GPUs", 512, 768, B, 2)
s1("c5 large + normalize_hiddenstates", 25, 64, 249, 1024, True)
print(json.dumps(dict(hbm_peak_gbs=HBM, tensor_peak_tflops=TC, kernels=rows), indent=1))
