"""CPU tests of the device index builder's algorithm (row N2): the per-item bodies that the kernels of
csrc/graph_device.cu run are compiled here as host C++ (tests/emu/graph_device_emu.cpp) and driven through the SAME
Python orchestration as the CUDA path (graph._device_graph_build with a numpy backend).  The emulated builder must
reproduce the native host builder element for element — and with it the reference-generated golden graphs — and the
pad stage must reproduce graph.pad_batch.  The GPU test (tests/test_gpu_graph_device.py) then only has to show that the
kernels equal the emulation."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from gemnet_pytorch_b200 import _lib, build_batch, graph
from tests import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_SRC = os.path.join(HERE, "emu", "graph_device_emu.cpp")
EMU_LIB = os.path.join(HERE, "emu", "build", "libgraph_device_emu.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    deps = [EMU_SRC, os.path.join(HERE, "..", "gemnet_pytorch_b200", "csrc", "graph_device_bodies.h"),
            os.path.join(HERE, "..", "include", "gemnet_b200.h")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", EMU_SRC, "-o", EMU_LIB],
                       check=True)
    return ctypes.CDLL(EMU_LIB)


class NumpyBackend:
    """The backend interface of graph._CudaBackend on host memory + the emulation library."""

    def __init__(self, lib):
        self.lib = lib
        self.stages = []

    _DT = {torch.int32: np.int32, torch.float32: np.float32, torch.float64: np.float64}

    def empty(self, n, dtype=torch.int32):
        return np.full(max(int(n), 1), -7, dtype=self._DT[dtype])[:int(n)]  # poison: every element must be written

    def host_buffer(self, n, dtype=torch.int32):
        return np.zeros(max(int(n), 1), dtype=self._DT[dtype])[:int(n)]

    @staticmethod
    def np_view(t):
        return t

    @staticmethod
    def copy_(dst, src):
        dst[...] = src

    def upload(self, arr):
        return np.ascontiguousarray(arr).copy()

    @staticmethod
    def ptr(t):
        return t.ctypes.data or None

    @staticmethod
    def host(t):
        return t.copy()

    @staticmethod
    def widen(t):
        return t.astype(np.int64)

    def run(self, stage, desc, *extra):
        self.stages.append(stage)
        rc = getattr(self.lib, "emu_graph_device_" + stage)(ctypes.byref(desc), *extra)
        assert rc == 0


def emulated_batch(emu, Rl, Zl, cutoff, int_cutoff, trip_only):
    return graph.build_batch_device(Rl, Zl, cutoff, int_cutoff, trip_only, _backend=NumpyBackend(emu))


CASES = [(True, list(range(100, 164)), 5.0, 10.0), (False, list(range(30, 42)), 5.0, 10.0),
         (True, list(range(20)), 8.0, 10.0), (False, [5, 9, 17], 3.0, 4.0), (False, [255], 5.0, 10.0),
         (False, list(range(8)), 8.0, 10.0)]


@pytest.mark.parametrize("trip_only,ids,cutoff,int_cutoff", CASES)
def test_emulated_device_builder_equals_host_builder(emu, trip_only, ids, cutoff, int_cutoff):
    Rl, Zl = H.molecules(ids)
    host = build_batch(Rl, Zl, cutoff, int_cutoff, trip_only, as_torch=False)
    dev = emulated_batch(emu, Rl, Zl, cutoff, int_cutoff, trip_only)
    assert set(host) <= set(dev)
    for k, v in host.items():
        assert np.array_equal(np.asarray(dev[k]).reshape(v.shape), v), k
    # the CSR row pointers the builder computes on the way (= what GraphPlan derives from the sorted reduce ids)
    N = np.array([len(z) for z in Zl], dtype=np.int32)
    a, hdr = graph._device_graph_build(NumpyBackend(emu), np.concatenate(Rl).astype(np.float64), N, cutoff, int_cutoff,
                                       trip_only)
    n_e = host["id_c"].shape[0]
    assert np.array_equal(a["rowptr3"], np.searchsorted(host["id3_reduce_ca"], np.arange(n_e + 1)))
    if not trip_only:
        assert np.array_equal(a["quad_ptr"], np.searchsorted(host["id4_reduce_ca"], np.arange(n_e + 1)))
        assert hdr[graph.HDR_QUADS] == host["id4_reduce_ca"].shape[0]


@pytest.mark.parametrize("name,trip_only", [("T", True), ("Q", False)])
def test_emulated_device_builder_matches_reference_fixture(emu, name, trip_only):
    """against the arrays the UNMODIFIED reference DataContainer produced (tests/golden/graph_{T,Q}.npz)"""
    fx = H.load_npz(f"graph_{name}.npz")
    Rl, Zl = H.molecules(fx["mol_ids"].tolist())
    mine = emulated_batch(emu, Rl, Zl, 5.0, 10.0, trip_only)
    exact = ["batch_seg", "id_c", "id_a", "id_undir", "id_swap", "Kidx3"]
    if not trip_only:
        exact += ["id4_int_a", "id4_int_b", "id4_reduce_intm_ca", "id4_expand_intm_db", "id4_reduce_intm_ab",
                  "id4_expand_intm_ab", "Kidx4"]
    for k in exact:
        assert np.array_equal(fx[k], mine[k]), k
    a = H.canonical_segments(fx["id3_reduce_ca"], fx["id3_expand_ba"])
    b = H.canonical_segments(mine["id3_reduce_ca"], mine["id3_expand_ba"])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    if not trip_only:
        ks = ["id4_reduce_ca", "id4_expand_db", "id4_reduce_cab", "id4_expand_abd"]
        a = H.canonical_segments(*[fx[k] for k in ks])
        b = H.canonical_segments(*[mine[k] for k in ks])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_emulated_device_builder_edge_cases(emu):
    # no molecules at all; single atoms (no edges); an empty molecule between two real ones; a far-apart pair
    for Rl, Zl in [([], []),
                   ([np.zeros((1, 3))], [np.array([1])]),
                   ([np.zeros((1, 3)), np.ones((1, 3))], [np.array([1]), np.array([6])]),
                   ([np.array([[0., 0, 0], [1, 0, 0]]), np.zeros((0, 3)), np.array([[0., 0, 0], [0, 1, 0], [0, 0, 1.5]])],
                    [np.array([1, 1]), np.zeros(0, int), np.array([8, 1, 1])]),
                   ([np.array([[0., 0, 0], [50, 0, 0]])], [np.array([1, 1])])]:
        for trip_only in (True, False):
            if len(Zl) == 0:
                host = None
            else:
                host = build_batch(Rl, Zl, 5.0, 10.0, trip_only, as_torch=False)
            dev = emulated_batch(emu, Rl, Zl, 5.0, 10.0, trip_only)
            if host is None:
                assert all(np.asarray(dev[k]).size == 0 for k in graph.TRIPLET_KEYS)
                continue
            for k, v in host.items():
                assert np.array_equal(np.asarray(dev[k]).reshape(v.shape), v), (k, trip_only)


def test_emulated_builder_flags_inconsistent_sizes(emu):
    be = NumpyBackend(emu)
    R = np.zeros((4, 3))
    with pytest.raises(_lib.GemnetB200Error):
        graph._device_graph_build(be, R, np.array([2, 1], dtype=np.int32), 5.0, 10.0, True)  # 3 != 4 atoms


def test_count_host_matches_builder():
    lib = _lib.load()
    for ids, cutoff in [(list(range(100, 164)), 5.0), (list(range(20)), 8.0), ([255], 5.0)]:
        Rl, Zl = H.molecules(ids)
        host = build_batch(Rl, Zl, cutoff, 10.0, True, as_torch=False)
        n_e, n_t = graph.count_batch(Rl, Zl, cutoff)
        assert (n_e, n_t) == (host["id_c"].shape[0], host["id3_reduce_ca"].shape[0])
    assert lib is not None


@pytest.mark.parametrize("ids,margin", [(list(range(100, 164)), 1.05), (list(range(3)), 1.5), ([255], 1.0)])
def test_emulated_static_build_and_pad_equals_pad_batch(emu, ids, margin):
    """static-shape use: count, edges, pad on capacity-sized arrays == graph.pad_batch of the host-built batch"""
    Rl, Zl = H.molecules(ids)
    host = build_batch(Rl, Zl, 5.0, 10.0, True)
    cap = graph.BatchCapacity.for_batches([host], margin=margin)
    cap = graph.BatchCapacity(cap.atoms, cap.edges, cap.triplets, cap.molecules + 2)  # a short batch
    want, (A, E, T, M) = graph.pad_batch(host, cap, check=False)
    be = NumpyBackend(emu)
    sb = graph.DeviceStaticBuilder(cap, 5.0, backend=be)
    sb.stage(Rl, Zl)
    sb.enqueue()
    got = sb.inputs()
    hdr = sb.header()
    assert hdr[graph.HDR_OVERFLOW] == 0 and (hdr[graph.HDR_EDGES], hdr[graph.HDR_TRIPLETS]) == (E, T)
    for k in graph.SERVING_KEYS:
        w = want[k].numpy()
        assert np.array_equal(np.asarray(got[k]).reshape(w.shape), w), k
    assert be.stages == ["count", "edges", "pad"]
    # a second, smaller batch through the same buffers (what a graph replay does)
    Rl2, Zl2 = H.molecules(ids[:1])
    host2 = build_batch(Rl2, Zl2, 5.0, 10.0, True)
    want2, _ = graph.pad_batch(host2, cap, check=False)
    sb.stage(Rl2, Zl2)
    sb.enqueue()
    got2 = sb.inputs()
    for k in graph.SERVING_KEYS:
        w = want2[k].numpy()
        assert np.array_equal(np.asarray(got2[k]).reshape(w.shape), w), k


def test_emulated_static_build_reports_overflow(emu):
    Rl, Zl = H.molecules(list(range(100, 164)))
    host = build_batch(Rl, Zl, 5.0, 10.0, True)
    small = graph.BatchCapacity(int(host["Z"].shape[0]) + 8, 1024, 4096, 64)
    sb = graph.DeviceStaticBuilder(small, 5.0, backend=NumpyBackend(emu))
    sb.stage(Rl, Zl)
    sb.enqueue()
    assert sb.header()[graph.HDR_OVERFLOW] & 6
    # the evaluation that still follows in the captured graph must see valid atomic numbers everywhere (Z - 1 indexes
    # the embedding table): the pad kernel did not run, the staging default covers the rows behind the real atoms
    assert (np.asarray(sb.inputs()["Z"]) >= 1).all()


@pytest.mark.parametrize("seed", range(6))
def test_emulated_device_builder_random_molecules(emu, seed):
    """random point clouds (1-45 atoms per molecule, incl. molecules larger than a warp, empty molecules, atoms exactly
    at the cutoff distance, coincident atoms) — the emulated device builder equals the host builder array for array"""
    rng = np.random.default_rng(100 + seed)
    Rl, Zl = [], []
    for _ in range(int(rng.integers(1, 9))):
        n = int(rng.choice([0, 1, 2, 3, 7, 20, 33, 45]))
        box = float(rng.choice([2.0, 4.0, 9.0]))
        R = rng.uniform(0.0, box, size=(n, 3))
        if n >= 3:
            R[1] = R[0] + np.array([3.0, 0.0, 0.0])   # a pair at exactly the cutoff (3.0, representable): <= keeps it
            R[2] = R[0]                               # coincident atoms: distance 0 is an edge as in the reference
        Rl.append(R)
        Zl.append(rng.integers(1, 9, size=n))
    if sum(len(z) for z in Zl) == 0:
        Rl.append(np.zeros((1, 3)))
        Zl.append(np.array([1]))
    for trip_only in (True, False):
        host = build_batch(Rl, Zl, 3.0, 4.5, trip_only, as_torch=False)
        dev = emulated_batch(emu, Rl, Zl, 3.0, 4.5, trip_only)
        for k, v in host.items():
            assert np.array_equal(np.asarray(dev[k]).reshape(v.shape), v), (k, trip_only, seed)
