"""Time the VQ forward/backward groups under engine ablations (SCP_DEBUG_ABLATE is read once per process).

Needs a library built with the ablation switches compiled in: SCP_BUILD_ABLATION=1 python -m speechclip_plus_b200.build
(a regular build ignores the variable; rebuild without it afterwards)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch, statistics
sys.path.insert(0, %r)
import speechclip_plus_b200 as scp
B,K,V,D=256,8,49408,512
g=torch.Generator(device="cuda").manual_seed(1)
table=torch.randn(V,D,device="cuda",generator=g)*0.02
kw=(torch.randn(B,K,D,device="cuda",generator=g)*0.02).requires_grad_(True)
gout=torch.randn(B,K,D,device="cuda",generator=g)
vq=scp.SimpleVectorQuantizer("fixed=0.1").cuda().train()
def timed(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev=[(torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a,b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a,b in ev)
st={}
def f():
    st["r"],st["o"]=vq.quantize_keywords(kw,table)
tf=timed(f)
def fb():
    f(); torch.autograd.grad(st["o"],[kw],grad_outputs=gout)
tb=timed(fb)-tf
print("RESULT fwd_ms=%%.3f bwd_ms=%%.3f" %% (tf,tb))
''' % ROOT
for mode in [0, 1, 2, 3]:
    env = dict(os.environ, SCP_DEBUG_ABLATE=str(mode))
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=300)
    out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print(f"ablate={mode} (1=no epilogue maths, 2=no MMAs):", out[0] if out else (r.stdout[-300:], r.stderr[-600:]))
