import sys, torch
sys.path.insert(0, ".")
from gemnet_pytorch_b200 import dense_ops as D, ops
dev = torch.device("cuda")
torch.manual_seed(0)
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a-b).abs().max()/b.abs().max()).item(), ((a-b).pow(2).mean().sqrt()/b.pow(2).mean().sqrt()).item(), (((a-b)/b.abs().clamp(min=1e-3*b.abs().max())).mean()).item()
ssilu = lambda t: torch.nn.functional.silu(t) / 0.6
for M,K,N in ((8186,128,128),(8186,128,64),(8186,64,128),(762,128,128),(8186,16,128)):
    x = torch.randn(M,K,device=dev)*1.5; W = torch.randn(N,K,device=dev)/K**0.5; g = torch.randn(M,N,device=dev); res = torch.randn(M,N,device=dev)
    xd, Wd = x.double(), W.double(); z = xd@Wd.t()
    y = D.dense_tc(x, D.packed_weight(W, False)); print(M,K,N,"plain   tc", "%.1e %.1e %+.1e"%rel(y, z), " cublas", "%.1e %.1e %+.1e"%rel(torch.mm(x,W.t()), z))
    y2, d2 = D.dense_tc(x, D.packed_weight(W, False), act=1, want_dact=True)
    zz = z.clone().requires_grad_(True); (dref,) = torch.autograd.grad(ssilu(zz).sum(), zz)
    z32 = torch.mm(x, W.t()); y32 = ops.scaled_silu(z32)
    print("        act     tc", "%.1e %.1e %+.1e"%rel(y2, ssilu(z)), " cublas+kernel", "%.1e %.1e %+.1e"%rel(y32, ssilu(z)))
    zc = z32.clone().requires_grad_(True); (d32,) = torch.autograd.grad(ssilu(zc).sum(), zc)
    print("        dact    tc", "%.1e %.1e %+.1e"%rel(d2, dref), " torch fp32", "%.1e %.1e %+.1e"%rel(d32, dref))
    gx = D.dense_tc(g, D.packed_weight(W, True), xmul=d2); gref = (g.double()*dref)@Wd
    gx32 = torch.mm(g*d32, W)
    print("        bwd     tc", "%.1e %.1e %+.1e"%rel(gx, gref), " cublas", "%.1e %.1e %+.1e"%rel(gx32, gref))
