import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
print("done")
