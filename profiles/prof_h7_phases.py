"""clock64() stamps of the H7 forward tcgen05 kernel (thread 0 of CTA 0).  python profiles/prof_h7_phases.py [n_mol]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gemnet_pytorch_b200 import _lib, build_batch, graph, ops  # noqa: E402

dev = torch.device("cuda:0")
n_mol = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Rl, Zl = bench.synthetic_molecules(n_mol, 1234)
inputs = {k: v.to(dev) for k, v in build_batch(Rl, Zl, 5.0, 10.0, True).items()}
plan = graph.GraphPlan(inputs, True)
E, T = plan.n_edges, plan.n_triplets
g = torch.Generator().manual_seed(0)
x, sph = torch.randn(E, 64, generator=g).to(dev), torch.randn(T, 7, generator=g).to(dev)
rw, W = torch.randn(E, 16, 7, generator=g).to(dev), (torch.randn(64, 16, 64, generator=g) / 32).to(dev)
with torch.no_grad():
    for _ in range(3):
        ops.segmented_bilinear(x, sph, rw, W, plan.id3_expand, plan.rowptr3, 1.0, 1)
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    _lib.load().gemnet_chain_set_profile_buffer(buf.data_ptr())
    ops.segmented_bilinear(x, sph, rw, W, plan.id3_expand, plan.rowptr3, 1.0, 1)
    torch.cuda.synchronize()
    _lib.load().gemnet_chain_set_profile_buffer(None)
t = [v for v in buf.cpu().tolist() if v]
names = ["tile start", "phase 1 done (this warp)", "As complete (all warps)", "previous epilogue done", "phase 2 done"]
print(f"n_mol={n_mol} E={E} T={T}: {len(t)} stamps, total {t[-1] - t[0]} cycles")
for i in range(1, min(len(t), 40)):
    print(f"{t[i] - t[i - 1]:8d}  {t[i] - t[0]:8d}  {names[(i - 1) % 5] if i < len(t) - 1 else 'last epilogue done'}")
