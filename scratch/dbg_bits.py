import torch, sys
sys.path.insert(0, '.')
from reslic_tcm_b200 import ops, _cabi, synthetic
dev='cuda:0'
torch.manual_seed(0)
tab=synthetic.scale_table(dev)
def run(B,C,h,w,want,explicit_ws):
    y=torch.randn(B,C,h,w,device=dev); mu=torch.randn_like(y); sg=torch.rand_like(y)*4+0.05
    out={}
    if explicit_ws:
        out["workspace"]=torch.zeros(int(_cabi.load().reslic_workspace_bytes(B)),dtype=torch.uint8,device=dev)
    r=ops.gc_forward(y,sg,mu,want=want,out=out,scale_table=tab)
    torch.cuda.synchronize()
    own=-(torch.log2(r.lik.double()).reshape(B,-1).sum(1))
    print(B,C,want,explicit_ws, "ok" if torch.allclose(r.bits,own,rtol=1e-6) else "BAD", r.bits[:6].tolist())
run(1,64,16,16,("lik","bits"),False)
run(8,64,48,32,("lik","bits"),False)
run(24,320,48,32,("lik","bits"),False)
run(24,320,48,32,("lik","bits","idx"),False)
run(24,320,48,32,("ste","lik","sym","idx","bits"),False)
run(24,320,48,32,("ste","lik","sym","idx","bits"),True)
b=synthetic.make_batch(2, range(24))
y,mu,sg=(b[k].to(dev) for k in ("y","mu","sigma"))
r=ops.gc_forward(y,sg,mu,want=("ste","lik","sym","idx","bits"),scale_table=tab)
own=-(torch.log2(r.lik.double()).reshape(24,-1).sum(1))
print((r.bits-own).abs().max().item(), r.bits[:6].tolist(), own[:6].tolist())
print(_cabi._ws[list(_cabi._ws)[0]][:96].view(torch.int32).tolist())
