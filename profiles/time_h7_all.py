"""Time the fp32 H7 kernels (forward and backward) for the triplet and quadruplet shapes on COLL-shaped synthetic
batches, isolated launches with an L2 flush in between.  Usage: python profiles/time_h7_all.py [n_molecules]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gemnet_pytorch_b200 import build_batch, graph, ops  # noqa: E402

dev = torch.device("cuda")
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush_buf.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return round(float(np.median(ts)), 1)


def case(name, n_x, expand, rowptr, S, C, I, O, algos):
    E, U = rowptr.numel() - 1, expand.numel()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n_x, C, generator=g).to(dev).requires_grad_()
    sph = torch.randn(U, S, generator=g).to(dev).requires_grad_()
    rw = torch.randn(E, I, S, generator=g).to(dev).requires_grad_()
    W = (torch.randn(C, I, O, generator=g) / 32).to(dev)
    go = torch.randn(E, O, generator=g).to(dev)
    res = {"case": name, "edges": E, "units": U}
    for algo in algos:
        with torch.no_grad():
            res[f"fwd_algo{algo}_us"] = timeit(lambda: ops.segmented_bilinear(x, sph, rw, W, expand, rowptr, 1.0, algo))
    out = ops.segmented_bilinear(x, sph, rw, W, expand, rowptr, 1.0, 0)
    res["bwd_us (incl. g_x memset)"] = timeit(lambda: torch.autograd.grad(out, (x, sph, rw), go, retain_graph=True))
    print(json.dumps(res), flush=True)


nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Rb, Zb = bench.synthetic_molecules(nmol, 4321)
inp = {k: v.to(dev) for k, v in build_batch(Rb, Zb, 5.0, 10.0, False).items()}
p = graph.GraphPlan(inp, False)
case("triplet", p.n_edges, p.id3_expand, p.rowptr3, 7, 64, 16, 64, (1, 0))
case("quad", int(p.abd_center.numel()), p.id4_expand_abd, p.rowptr4, 49, 32, 32, 32, (0,))
