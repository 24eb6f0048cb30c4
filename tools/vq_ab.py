"""A/B timing of the VQ forward / backward groups under environment switches (each setting in its own process).

    python tools/vq_ab.py SCP_VQ_RESIDENT=0 SCP_VQ_RESIDENT=1 [--shape B,K,V,D]
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
shape = "256,8,49408,512"
settings = []
args = sys.argv[1:]
while args:
    a = args.pop(0)
    if a == "--shape":
        shape = args.pop(0)
    else:
        settings.append(a)
CODE = r'''
import sys, torch, statistics
sys.path.insert(0, %r)
import speechclip_plus_b200 as scp
B,K,V,D=%s
g=torch.Generator(device="cuda").manual_seed(1)
table=torch.randn(V,D,device="cuda",generator=g)*0.02
kw=(torch.randn(B,K,D,device="cuda",generator=g)*0.02).requires_grad_(True)
gout=torch.randn(B,K,D,device="cuda",generator=g)
vq=scp.SimpleVectorQuantizer("fixed=0.1").cuda().train()
def timed(fn,n=20):
    for _ in range(5): fn()
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
    f(); st["g"]=torch.autograd.grad(st["o"],[kw],grad_outputs=gout)[0]
tb=timed(fb)-tf
r=st["r"]
print("RESULT fwd_ms=%%.4f bwd_ms=%%.4f idxsum=%%d pp=%%.6f cp=%%.6f gnorm=%%.8e" %% (tf,tb,int(r["targets"].sum()),float(r["prob_perplexity"]),float(r["code_perplexity"]),float(st["g"].double().norm())))
''' % (ROOT, shape)
for setting in settings or [""]:
    env = dict(os.environ)
    for kv in setting.split(","):
        if "=" in kv:
            k, v = kv.split("=", 1)
            env[k] = v
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=600)
    out = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print(f"[{setting}]", out[0] if out else (r.stdout[-300:], r.stderr[-800:]), flush=True)
