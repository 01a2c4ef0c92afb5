"""Micro-benchmark of the fused tcgen05 dense layer vs cuBLAS fp32 (+ pointwise): fixed vs per-tile cost."""
import sys

import torch

sys.path.insert(0, ".")
from gemnet_pytorch_b200 import dense_ops as D, ops  # noqa: E402

dev = torch.device("cuda")


def bench(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / n * 1e3


for M in (128, 1024, 8186, 18944, 180442):
    x = torch.randn(M, 128, device=dev)
    W = torch.randn(128, 128, device=dev) / 11
    pk = D.packed_weight(W, False)
    small = torch.randn(1024, device=dev)
    t_tc = bench(lambda: D.dense_tc(x, pk, act=1))
    t_tc_alt = bench(lambda: (D.dense_tc(x, pk, act=1), ops.scaled_silu(small)))
    t_mm = bench(lambda: torch.mm(x, W.t()))
    t_mm_act = bench(lambda: ops.scaled_silu(torch.mm(x, W.t())))
    t_small = bench(lambda: ops.scaled_silu(small))
    print(f"M={M:7d}: dense_tc {t_tc:7.1f} us | dense_tc + tiny kernel {t_tc_alt:7.1f} us | cuBLAS mm {t_mm:7.1f} us | "
          f"mm + silu kernel {t_mm_act:7.1f} us | tiny kernel alone {t_small:5.1f} us")


# accumulation bias of the tensor-core path: signed mean relative error vs fp64 on well-conditioned entries
torch.manual_seed(0)
for K in (32, 128):
    x = torch.rand(4096, K, device=dev) + 0.5          # positive operands: no cancellation, errors are pure rounding
    W = torch.rand(128, K, device=dev) + 0.5
    ref = x.double() @ W.double().t()
    for name, y in (("dense_tc bf16x3x3", D.dense_tc(x, D.packed_weight(W, False))), ("cuBLAS fp32", torch.mm(x, W.t()))):
        rel = (y.double() - ref) / ref
        print(f"K={K:4d} {name:16s}: mean signed rel err {rel.mean().item():+.3e}, rms {rel.pow(2).mean().sqrt().item():.3e}, max {rel.abs().max().item():.3e}")
