"""BASELINE configs[4] ("large-cutoff stress"): segmented-GEMM (H7) throughput sweep at cutoff 8 A.

SURVEY §8d input 3: COLL molecules give only ~1.3 k triplets/molecule at 8 A, so the ~5 k triplets/molecule of the config
are reached with synthetic 18-atom clusters (18*17*16 = 4896 triplets): atoms uniform in a 3.5 A ball, rejection-sampled
minimum distance 0.9 A, seed 1234.  Both inputs are reported.  Per size: isolated launches, L2 flushed, CUDA events;
algorithmic bytes of §8d over the time = GB/s, fraction of the measured HBM peak.
Usage: python profiles/sweep_h7_config5.py [--gpus-note N]   (one process per GPU replicates the sweep: no collective)"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gemnet_pytorch_b200 import build_batch, graph, ops  # noqa: E402

dev = torch.device("cuda")
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
HBM, TC, KIND = bench.peaks()


def clusters(n_mol, n_atoms=18, radius=3.5, dmin=0.9, seed=1234):
    rng = np.random.default_rng(seed)
    Rl, Zl = [], []
    for _ in range(n_mol):
        pts = []
        while len(pts) < n_atoms:
            p = rng.uniform(-radius, radius, size=3)
            if p @ p <= radius * radius and all(np.linalg.norm(p - q) >= dmin for q in pts):
                pts.append(p)
        Rl.append(np.array(pts))
        Zl.append(rng.choice([1, 6, 8], size=n_atoms, p=[0.543, 0.330, 0.127]))
    return Rl, Zl


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
        ts.append(s.elapsed_time(e) * 1e-3)
    return float(np.median(ts))


def row(kind, shape, n_mol, E, U, S, C, I, O, t, algo):
    b, f = bench.bilinear_bytes(E, U, S, C, I, O), bench.bilinear_flops(E, U, S, C, I, O)
    print(json.dumps({"input": kind, "path": shape, "kernel": algo, "molecules": n_mol, "edges": E, "units": U,
                      "units_per_molecule": round(U / n_mol), "us": round(t * 1e6, 1), "GBs": round(b / t / 1e9, 1),
                      "frac_hbm": round(b / t / 1e9 / HBM, 4), "TFLOPs": round(f / t / 1e12, 2)}), flush=True)


def sweep(kind, make, sizes_T, sizes_Q):
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for n in sizes_T:
            Rl, Zl = make(n)
            p = graph.GraphPlan({k: v.to(dev) for k, v in build_batch(Rl, Zl, 8.0, 8.0, True).items()}, True)
            E, T = p.n_edges, p.n_triplets
            x = torch.randn(E, 64, generator=g).to(dev)
            sph = torch.randn(T, 7, generator=g).to(dev)
            rw = torch.randn(E, 16, 7, generator=g).to(dev)
            W = (torch.randn(64, 16, 64, generator=g) / 32).to(dev)
            for algo, name in ((1, "tcgen05 bf16x3"), (0, "fp32 CUDA cores")):
                t = timeit(lambda: ops.segmented_bilinear(x, sph, rw, W, p.id3_expand, p.rowptr3, 1.0, algo))
                row(kind, "triplet (7,64,16,64)", n, E, T, 7, 64, 16, 64, t, name)
            del p, x, sph, rw
        for n in sizes_Q:
            Rl, Zl = make(n)
            p = graph.GraphPlan({k: v.to(dev) for k, v in build_batch(Rl, Zl, 8.0, 8.0, False).items()}, False)
            E, M, Q = p.n_edges, int(p.abd_center.numel()), p.n_quads
            x = torch.randn(M, 32, generator=g).to(dev)
            ylm = torch.randn(Q, 49, generator=g).to(dev)
            rw = torch.randn(E, 32, 49, generator=g).to(dev)
            W = (torch.randn(32, 32, 32, generator=g) / 32).to(dev)
            t = timeit(lambda: ops.segmented_bilinear(x, ylm, rw, W, p.id4_expand_abd, p.rowptr4, 1.0, 0), 5)
            row(kind, "quadruplet (49,32,32,32)", n, E, Q, 49, 32, 32, 32, t, "fp32 CUDA cores (split pipeline)")
            del p, x, ylm, rw
            torch.cuda.empty_cache()


print(json.dumps({"hbm_peak_GBs": HBM, "peak_kind": KIND, "device": torch.cuda.get_device_name(0)}))
sweep("18-atom clusters, cutoff 8 A", clusters, (64, 256, 1024, 4096), (16, 64))
sweep("COLL-shaped, cutoff 8 A", lambda n: bench.synthetic_molecules(n, 1234), (64, 1024, 4096), (64,))
