"""Launch each hot-path kernel a few times on a large COLL-shaped batch so that `ncu -k regex:...` can capture it."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gemnet_pytorch_b200 import build_batch, dense_ops, graph, layers, ops  # noqa: E402

dev = torch.device("cuda")
Rb, Zb = bench.synthetic_molecules(2048, 4321)
inp = {k: v.to(dev) for k, v in build_batch(Rb, Zb, 5.0, 10.0, True).items()}
p = graph.GraphPlan(inp, True)
E, T = p.n_edges, p.n_triplets
g = torch.Generator().manual_seed(0)
with torch.no_grad():
    for _ in range(3):
        D, V = ops.edge_geometry(inp["R"], p.id_c, p.id_a)
        ops.sph_bessel_radial(D, 7, 6, 5.0, 5)
        ops.triplet_sph(inp["R"], p.id_c, p.id_a, p.id3_reduce, p.id3_expand, 7)
        m = torch.randn(E, 128, generator=g).to(dev)
        r = torch.randn(E, 128, generator=g).to(dev)
        ops.edge_to_atom(m, r, p.id_a, p.atom_perm, p.atom_rowptr, 1.0)
        W = (torch.randn(128, 128, generator=g) / 11).to(dev)
        dense_ops.dense_tc(m, dense_ops.packed_weight(W, False), act=1)
torch.cuda.synchronize()
print("done", E, T)
