"""Runs the VQ forward + backward a few times at one shape (for ncu).   python tools/vq_bwd_profile.py [B K V D] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import speechclip_plus_b200 as scp  # noqa: E402

B, K, V, D = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (256, 8, 49408, 512)
iters = int(sys.argv[5]) if len(sys.argv) >= 6 else 3
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(1)
table = torch.randn(V, D, device=dev, generator=gen) * 0.02
kw = (torch.randn(B, K, D, device=dev, generator=gen) * 0.02).requires_grad_(True)
gout = torch.randn(B, K, D, device=dev, generator=gen)
vq = scp.SimpleVectorQuantizer("fixed=0.1").to(dev).train()
for _ in range(iters):
    res, out = vq.quantize_keywords(kw, table)
    torch.autograd.grad(out, [kw], grad_outputs=gout)
torch.cuda.synchronize()
print("ok")
