"""Device time of the masked InfoNCE forward+backward at the global batch sizes of 1/2/4/8 GPUs (local rows = 256),
through the public module, replayed as a CUDA graph (what bench.py does)."""
import sys, os, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speechclip_plus_b200 as scp
D = 512
crit = scp.MaskedContrastiveLoss(temperature=0.07, temperature_trainable=True).cuda()
for N in (256, 512, 1024, 2048):
    g = torch.Generator(device="cuda").manual_seed(N)
    a = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=-1).requires_grad_(True)
    b = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=-1).requires_grad_(True)
    ids = torch.randint(0, 6000, (N,), device="cuda", generator=g)
    def step():
        loss = crit(a, b, ids, local_rows=(0, 256))
        return torch.autograd.grad(loss, [a, b, crit.temperature])
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for _ in range(3): graph.replay()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
    for x, y in ev:
        x.record(); graph.replay(); y.record()
    torch.cuda.synchronize()
    print(f"N={N}: {statistics.median(x.elapsed_time(y) for x, y in ev) * 1e3:.1f} us per fwd+bwd (graph)")
