#!/bin/bash
# per-kernel device times (ncu, one pass) of the VQ forward+backward under engine ablations / A-B switches
# usage: tools/vq_ncu_times.sh "ENV1=a,ENV2=b" "ENV1=c" ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/vq_drv.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import speechclip_plus_b200 as scp
B,K,V,D=256,8,49408,512
g=torch.Generator(device="cuda").manual_seed(1)
table=torch.randn(V,D,device="cuda",generator=g)*0.02
kw=(torch.randn(B,K,D,device="cuda",generator=g)*0.02).requires_grad_(True)
gout=torch.randn(B,K,D,device="cuda",generator=g)
vq=scp.SimpleVectorQuantizer("fixed=0.1").cuda().train()
for _ in range(3):
    r,o=vq.quantize_keywords(kw,table)
    torch.autograd.grad(o,[kw],grad_outputs=gout)
torch.cuda.synchronize()
PY
for setting in "$@"; do
  tag=$(echo "$setting" | tr ',=' '__')
  env $(echo "$setting" | tr ',' ' ') ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/vqt_$tag.csv python /tmp/vq_drv.py > /dev/null 2>&1
  echo "== $setting"
  python - gpurun_out/vqt_$tag.csv <<'PY'
import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; d=collections.OrderedDict()
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        k=r[hdr.index('Kernel Name')]; v=float(r[hdr.index('Metric Value')])
        d.setdefault(k,[]).append(v)
for k,v in d.items():
    v=v[len(v)//3:]  # skip the first (cold) iteration
    print(f"  {k[:100]:100s} n={len(v):2d} avg={sum(v)/len(v)/1000:8.1f} us")
PY
done
