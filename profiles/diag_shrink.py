"""Does the tensor-core product shrink toward zero?  slope = <(y - ref) * ref> / <ref^2> (0 for unbiased rounding)."""
import sys, torch
sys.path.insert(0, ".")
from gemnet_pytorch_b200 import dense_ops as D
dev = torch.device("cuda")
torch.manual_seed(1)
for K in (16, 64, 128):
    x = torch.randn(16384, K, device=dev); W = torch.randn(128, K, device=dev) / K ** 0.5
    ref = x.double() @ W.double().t()
    for name, y in (("dense_tc", D.dense_tc(x, D.packed_weight(W, False))), ("cuBLAS  ", torch.mm(x, W.t()))):
        e = y.double() - ref
        slope = (e * ref).sum() / (ref * ref).sum()
        big = ref.abs() > 1.0
        slope_big = (e[big] * ref[big]).sum() / (ref[big] ** 2).sum()
        print(f"K={K:4d} {name}: slope {slope.item():+.3e} (|ref|>1: {slope_big.item():+.3e}), rms rel {(e.pow(2).mean().sqrt()/ref.pow(2).mean().sqrt()).item():.2e}")
