"""Time the H7 forward kernels (tcgen05 algo 1 vs fp32 algo 0) on COLL-shaped synthetic batches."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gemnet_pytorch_b200 import build_batch, graph  # noqa: E402

dev = torch.device("cuda")
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
hbm, tcp, _ = bench.peaks()
SEED = 4321
if "--seed" in sys.argv:  # --seed 1234 = the batch bench.py times (its `roofline` entry)
    i = sys.argv.index("--seed")
    SEED = int(sys.argv[i + 1])
    del sys.argv[i:i + 2]
for nmol in [int(a) for a in sys.argv[1:]] or [64, 2048]:
    Rb, Zb = bench.synthetic_molecules(nmol, SEED)
    big = {k: v.to(dev) for k, v in build_batch(Rb, Zb, 5.0, 10.0, True).items()}
    plan = graph.GraphPlan(big, True)
    res = {"molecules": nmol, "edges": plan.n_edges, "triplets": plan.n_triplets}
    for algo in (1, 0):
        t, b, f = bench.time_bilinear(plan, dev, flush_buf.zero_, 10, algo=algo)
        res[f"algo{algo}_us"] = round(t * 1e6, 1)
        res[f"algo{algo}_GBs"] = round(b / t / 1e9, 1)
        res[f"algo{algo}_frac_hbm"] = round(b / t / 1e9 / hbm, 4)
    print(json.dumps(res))
