"""Device times and achieved HBM bandwidth of the kernels outside the bench step (S1' variants and the "next" rows
N1 / N3 / N4 of SURVEY section 8(f)) at BASELINE config-3 sizes, CUDA events on the launching stream, median of 20.

    python tools/aux_kernel_times.py > gpurun_out/aux_kernels.json
"""
import json, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speechclip_plus_b200 as scp
from speechclip_plus_b200.module.speech_encoder_plus import fuse_upstream_features
from speechclip_plus_b200.module.clip_glue import splice_keywords
from speechclip_plus_b200.module.cif import integrate_and_fire
from speechclip_plus_b200.module.kw_bn import Kw_BatchNorm, Kw_BatchNorm_dynamic

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7122)


def timed(fn, n=20, graph=False):
    """median device time; graph=True replays the call as a CUDA graph so that the CPU launch rate of the small
    Python-wrapped kernels (~30 us per call) does not hide the device time"""
    for _ in range(5):
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


out = []


def report(name, seconds, bytes_, note):
    out.append(dict(kernel=name, ms=seconds * 1e3, algorithmic_bytes=bytes_, achieved_gbs=bytes_ / seconds / 1e9,
                    frac_of_hbm_peak=bytes_ / seconds / 1e9 / PEAK, note=note))


# ---- S1 / S1' forward + weight-gradient backward: L=13, B=256, T=249, D=768 fp32
L, B, T, D = 13, 256, 249, 768
storage = [torch.randn(T, B, D, device=dev, generator=g) for _ in range(L)]
layers = [s.transpose(0, 1) for s in storage]
gy = torch.randn(B, T, D, device=dev, generator=g)
by = (L + 1) * B * T * D * 4
for ntype, norm in ((None, False), ("s3prl", True), ("method1", True), ("method2", True)):
    layer = scp.WeightedSumLayer(L, normalize_features=(ntype == "s3prl")).to(dev)
    def fwd():
        return fuse_upstream_features(layers, layer, norm, ntype or "s3prl")[0]
    t_f = timed(fwd)
    def fb():
        y = fwd()
        torch.autograd.grad(y, [layer.weights], grad_outputs=gy)
    t_fb = timed(fb)
    extra = L * B * T * D * 4 if ntype == "method2" else 0   # the statistics pre-pass reads the layers once more
    report(f"S1 fwd [{ntype or 'plain'}]", t_f, by + extra, "L*B*T*D*4 read (+ the same again for method2's statistics) + B*T*D*4 written")
    report(f"S1 bwd weights [{ntype or 'plain'}]", t_fb - t_f, by, "L*B*T*D*4 + B*T*D*4 read")

# ---- N1 keyword batch-norm, M = 2048 x 512 (latency-bound: 4 MB)
K, Dk = 8, 512
x = torch.randn(B, K, Dk, device=dev, generator=g).requires_grad_(True)
gk = torch.randn(B, K, Dk, device=dev, generator=g)
for name, layer in (("eachKw parallel", Kw_BatchNorm(K, Dk, "eachKw", torch.zeros(Dk), torch.ones(Dk), parallel=True)),
                    ("dynamic", Kw_BatchNorm_dynamic(Dk, torch.zeros(Dk), torch.ones(Dk)))):
    layer = layer.to(dev).train()
    layer.bn_layer.track_running_stats = True
    t_f = timed(lambda: layer(x), graph=True)
    def fb():
        y = layer(x)
        torch.autograd.grad(y, [x, layer.bn_layer.weight, layer.bn_layer.bias], grad_outputs=gk)
    t_fb = timed(fb, graph=True)
    report(f"N1 kw batch-norm fwd [{name}]", t_f, 3 * B * K * Dk * 4, "3 launches (graph replay); latency-bound: 4 MB tensor")
    report(f"N1 kw batch-norm bwd [{name}]", t_fb - t_f, 5 * B * K * Dk * 4, "3 launches (graph replay); latency-bound")

# ---- N3 splice: B=256, 77 x 512, fp32 and fp16 text tower
for dtype in (torch.float32, torch.float16):
    V, Lt, Kmax = 49408, 77, 75
    table = (torch.randn(V, Dk, device=dev, generator=g) * 0.02).to(dtype)
    pos = (torch.randn(Lt, Dk, device=dev, generator=g) * 0.01).to(dtype)
    kw = (torch.randn(B, Kmax, Dk, device=dev, generator=g) * 0.02).requires_grad_(True)
    num = torch.randint(4, 20, (B,), device=dev, generator=g)
    t_f = timed(lambda: splice_keywords(kw, num, table, pos, V - 2, V - 1), graph=True)
    es = 4 if dtype == torch.float32 else 2
    report(f"N3 splice fwd [{str(dtype).split('.')[-1]}]", t_f, B * Lt * Dk * es * 2, "B*77*D*s written + ~as much read (L2-resident table rows)")

# ---- N4 CIF: B=256 utterances x 249 frames x 768
S, C = 249, 768
xs = torch.randn(B, S, C, device=dev, generator=g).requires_grad_(True)
raw = (torch.rand(B, S, device=dev, generator=g) * 0.1).requires_grad_(True)
target = torch.full((B,), 12, device=dev)
def cif_f():
    alpha = raw * ((target.float() + 1e-5) / raw.sum(1)).unsqueeze(1)
    return integrate_and_fire(xs, alpha, 1.0, target)["dsample_feats"]
t_f = timed(cif_f)
gcf = torch.randn(B, 12, C, device=dev, generator=g)
def cif_fb():
    o = cif_f()
    torch.autograd.grad(o, [xs, raw], grad_outputs=gcf)
t_fb = timed(cif_fb)
report("N4 CIF integrate-and-fire fwd", t_f, B * S * C * 4 + B * 13 * C * 4, "includes the host read of max(feat_len) (one sync, as in the reference)")
report("N4 CIF integrate-and-fire bwd", t_fb - t_f, 2 * B * S * C * 4 + B * 12 * C * 4, "d_input written + input read + g_out read")
print(json.dumps(dict(hbm_peak_gbs=PEAK, kernels=out), indent=1))
