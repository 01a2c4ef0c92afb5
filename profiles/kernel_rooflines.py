"""Per-row kernel roofline table: algorithmic bytes / CUDA-event time of isolated launches
vs the measured HBM peak, on COLL-shaped synthetic batches.  Usage: python profiles/kernel_rooflines.py [nT] [nQ]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gemnet_pytorch_b200 import basis_consts, build_batch, graph, layers, ops  # noqa: E402

dev = torch.device("cuda")
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
HBM, _, KIND = bench.peaks()


def timeit(fn, reps=10):
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


rows = []


def report(row, kernel, nbytes, t, note=""):
    rows.append({"row": row, "kernel": kernel, "MB": round(nbytes / 1e6, 2), "us": round(t * 1e6, 1),
                 "GBs": round(nbytes / t / 1e9, 1), "frac_hbm": round(nbytes / t / 1e9 / HBM, 4), "note": note})
    print(json.dumps(rows[-1]), flush=True)


def main():
    nT = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    nQ = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        # ---------------- triplet-sized graph
        Rb, Zb = bench.synthetic_molecules(nT, 4321)
        inp = {k: v.to(dev) for k, v in build_batch(Rb, Zb, 5.0, 10.0, True).items()}
        p = graph.GraphPlan(inp, True)
        A, E, T = p.n_atoms, p.n_edges, p.n_triplets
        print(json.dumps({"graph_T": {"molecules": nT, "atoms": A, "edges": E, "triplets": T}, "hbm_peak": HBM,
                          "peak_kind": KIND}))
        R = inp["R"]
        D, V = ops.edge_geometry(R, p.id_c, p.id_a)
        report("H1", "edge_geometry_fwd", E * (8 + 16) + A * 12, timeit(lambda: ops.edge_geometry(R, p.id_c, p.id_a)))
        freq = torch.arange(1, 7, device=dev, dtype=torch.float32) * 3.14159
        report("H3", "bessel_rbf_fwd", E * (4 + 24), timeit(lambda: ops.bessel_rbf(D, freq, 5.0, 5)))
        report("H4 radial", "sph_bessel_radial_fwd (fp64 inside)", E * (4 + 168),
               timeit(lambda: ops.sph_bessel_radial(D, 7, 6, 5.0, 5)))
        report("H2+H4 angular", "triplet_sph_fwd", T * (8 + 28) + E * 8 + A * 12,
               timeit(lambda: ops.triplet_sph(R, p.id_c, p.id_a, p.id3_reduce, p.id3_expand, 7)))
        rad = ops.sph_bessel_radial(D, 7, 6, 5.0, 5)
        Wd = torch.randn(7, 6, 16, generator=g).to(dev)
        los = layers.l_of_s_tensor(7, False, dev)
        report("H6", "downproj_fwd (S=7,I=16)", E * (168 + 448), timeit(lambda: ops.down_projection(rad, Wd, los)))
        x = torch.randn(E, 64, generator=g).to(dev)
        sph = torch.randn(T, 7, generator=g).to(dev)
        rw = torch.randn(E, 16, 7, generator=g).to(dev)
        W = (torch.randn(64, 16, 64, generator=g) / 32).to(dev)
        bts = bench.bilinear_bytes(E, T, 7, 64, 16, 64)
        report("H7 fwd", "segbil_fwd_tc_kernel (tcgen05 bf16x3)", bts,
               timeit(lambda: ops.segmented_bilinear(x, sph, rw, W, p.id3_expand, p.rowptr3, 1.0, 1)),
               f"{bench.bilinear_flops(E, T, 7, 64, 16, 64) / 1e9:.1f} GFLOP")
        report("H7 fwd", "segbil_fwd_kernel (fp32 CUDA cores)", bts,
               timeit(lambda: ops.segmented_bilinear(x, sph, rw, W, p.id3_expand, p.rowptr3, 1.0, 0)))
        m = torch.randn(E, 128, generator=g).to(dev)
        r = torch.randn(E, 128, generator=g).to(dev)
        report("H10", "edge_to_atom_fwd (F=128)", 4 * (2 * E * 128 + A * 128) + 4 * E,
               timeit(lambda: ops.edge_to_atom(m, r, p.id_a, p.atom_perm, p.atom_rowptr, 1.0)))
        Fca = torch.randn(E, 1, generator=g).to(dev)
        report("H12", "force_readout_fwd", E * (4 + 12 + 4) + A * 12,
               timeit(lambda: ops.force_readout(Fca, V, p.id_a, p.atom_perm, p.atom_rowptr)))
        report("N1", "silu_fwd_kernel (E x 128)", 2 * 4 * E * 128, timeit(lambda: ops.scaled_silu(m)))
        report("N1", "gather_rows (E x 128 by id_swap)", 2 * 4 * E * 128 + 4 * E,
               timeit(lambda: ops.gather_rows(m, p.id_swap, p.swap_perm, p.swap_rowptr)))
        del inp, p, x, sph, rw, m, r, rad
        torch.cuda.empty_cache()

        # ---------------- quadruplet-sized graph
        Rb, Zb = bench.synthetic_molecules(nQ, 99)
        inp = {k: v.to(dev) for k, v in build_batch(Rb, Zb, 5.0, 10.0, False).items()}
        p = graph.GraphPlan(inp, False)
        A, E, M, Q = p.n_atoms, p.n_edges, int(p.abd_center.numel()), p.n_quads
        print(json.dumps({"graph_Q": {"molecules": nQ, "atoms": A, "edges": E, "intm_triplets": M, "quadruplets": Q}}))
        R = inp["R"]
        report("H2 (quad intm)", "quad_intm_fwd (a-b<-d side: Y_l0 + rejection)", M * (12 + 28 + 12),
               timeit(lambda: ops.quad_intermediate(R, p.abd_center, p.abd_first, p.abd_second, 7, True)))
        sph_abd, bd = ops.quad_intermediate(R, p.abd_center, p.abd_first, p.abd_second, 7, True)
        cs, ac = ops.quad_intermediate(R, p.cab_center, p.cab_first, p.cab_second, 7, False)
        report("H2+H5", "quad_ylm_fwd (49 Y_lm / quadruplet)", Q * (8 + 196) + M * (8 + 24),
               timeit(lambda: ops.quad_ylm(cs, ac, bd, p.id4_reduce_cab, p.id4_expand_abd, 7)))
        ylm = ops.quad_ylm(cs, ac, bd, p.id4_reduce_cab, p.id4_expand_abd, 7)
        xq = torch.randn(M, 32, generator=g).to(dev)
        rwq = torch.randn(E, 32, 49, generator=g).to(dev)
        Wq = (torch.randn(32, 32, 32, generator=g) / 32).to(dev)
        bts = bench.bilinear_bytes(E, Q, 49, 32, 32, 32) + 4 * (M - E) * 32
        report("H7 fwd (quad)", "segbil_fwd_kernel<49,32,32,32> (fp32 CUDA cores)", bts,
               timeit(lambda: ops.segmented_bilinear(xq, ylm, rwq, Wq, p.id4_expand_abd, p.rowptr4, 1.0, 0), 5),
               f"{bench.bilinear_flops(E, Q, 49, 32, 32, 32) / 1e9:.1f} GFLOP")
    print(json.dumps({"rows": rows}))


if __name__ == "__main__":
    main()
