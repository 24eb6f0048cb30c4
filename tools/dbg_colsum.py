import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speechclip_plus_b200 as scp
from oracle import speechclip_oracle as oracle
for shape in [(7, 5, 1000, 64), (33, 1, 600, 128), (1, 1, 300, 64), (40, 8, 4096, 256)]:
    B, K, V, D = shape
    gen = torch.Generator().manual_seed(V + B)
    table = torch.randn(V, D, generator=gen) * 0.02 + 0.003 * torch.randn(1, D, generator=gen)
    kw = torch.randn(B, K, D, generator=gen) * table.std(0) + table.mean(0)
    vq = scp.SimpleVectorQuantizer("fixed=0.1").cuda().train()
    res, out = vq.quantize_keywords(kw.cuda(), table.cuda())
    ref, _ = oracle.vq_audio_features(kw.double(), table.double(), torch.tensor([0.1], dtype=torch.float64))
    a, b = res["avg_probs"].double().cpu(), ref["avg_probs"].reshape(-1)
    d = (a - b).abs()
    print(shape, "avg max err", d.max().item(), "at", d.argmax().item(), "ref max", b.max().item(), "sum ours", a.sum().item(),
          "bad cols", (d > 1e-3 * b.max()).nonzero().flatten()[:12].tolist(), "n bad", int((d > 1e-3 * b.max()).sum()))
    for key in ("code_perplexity", "prob_perplexity", "ent_per_t"):
        print("   ", key, res[key].flatten()[:3].tolist(), ref[key].flatten()[:3].tolist())
