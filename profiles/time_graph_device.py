"""Row N2: time the device index builder (csrc/graph_device.cu) against the native host builder (row H0) on the bench's
COLL-shaped synthetic batches.  Device times are CUDA-event times of the kernel stages only (inputs resident);
`e2e_ms` is build_batch_device from host lists to a complete device dict (uploads, allocations and the header reads
included); the host time is graph.build_batch (int64 tensors, as the reference's DataContainer returns them)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gemnet_pytorch_b200 import build_batch, graph  # noqa: E402

dev = torch.device("cuda")


def cuda_ms(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def host_ms(fn, n=5):
    fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t.append(time.perf_counter() - t0)
    return float(np.median(t)) * 1e3


for model, nmol in [("T", 64), ("Q", 64), ("T", 2048)] if len(sys.argv) < 2 else [(sys.argv[1], int(sys.argv[2]))]:
    trip = model == "T"
    Rl, Zl = bench.synthetic_molecules(nmol, 4321)
    res = {"model": model, "molecules": nmol}
    res["host_builder_ms"] = host_ms(lambda: build_batch(Rl, Zl, 5.0, 10.0, trip))
    res["device_builder_e2e_ms"] = host_ms(lambda: graph.build_batch_device(Rl, Zl, 5.0, 10.0, trip, dev))
    # stage times on resident inputs: rebuild with exact capacities through the same descriptor
    be = graph._CudaBackend(dev)
    N = np.array([len(z) for z in Zl], dtype=np.int32)
    R64 = np.ascontiguousarray(np.concatenate(Rl).astype(np.float64))
    a, hdr = graph._device_graph_build(be, R64, N, 5.0, 10.0, trip)
    res.update(edges=int(hdr[graph.HDR_EDGES]), triplets=int(hdr[graph.HDR_TRIPLETS]), quadruplets=int(hdr[graph.HDR_QUADS]))
    d = graph.DeviceGraphDesc()
    d.R, d.n_per_mol, d.hdr = be.ptr(a["R64"]), be.ptr(a["n_per_mol"]), be.ptr(a["hdr"])
    d.cutoff, d.int_cutoff, d.triplets_only = 5.0, 10.0, int(trip)
    d.cap_atoms, d.cap_molecules = int(R64.shape[0]), len(N)
    d.cap_edges, d.cap_triplets = int(hdr[graph.HDR_EDGES]), int(hdr[graph.HDR_TRIPLETS])
    d.cap_int_edges, d.cap_intm, d.cap_quads = int(hdr[graph.HDR_INT_EDGES]), int(hdr[graph.HDR_INTM]), int(hdr[graph.HDR_QUADS])
    for k in a:
        if k not in ("R64", "n_per_mol", "hdr"):
            setattr(d, k, be.ptr(a[k]))
    res["count_stage_us"] = cuda_ms(lambda: be.run("count", d)) * 1e3
    res["edges_stage_us"] = cuda_ms(lambda: be.run("edges", d)) * 1e3
    if not trip:
        res["quads_stage_us"] = cuda_ms(lambda: be.run("quads", d)) * 1e3
    print(json.dumps(res))
