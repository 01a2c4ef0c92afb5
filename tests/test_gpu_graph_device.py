"""GPU tests of the device index builder (row N2, csrc/graph_device.cu): the kernels reproduce the native host builder
element for element (which tests/test_host_cpu.py pins against the reference-generated golden graphs and
tests/test_graph_device_cpu.py shows for the same per-item code run on the host), the static-shape use reproduces
graph.pad_batch, and the model / serving paths give the same energies and forces from device-built indices."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

from oracle import gemnet_oracle as go  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


CASES = [(True, list(range(100, 164)), 5.0, 10.0), (False, list(range(30, 42)), 5.0, 10.0),
         (True, list(range(20)), 8.0, 10.0), (False, [5, 9, 17], 3.0, 4.0), (False, [255], 5.0, 10.0),
         (False, list(range(64)), 5.0, 10.0), (True, list(range(256)) * 8, 5.0, 10.0)]


@pytest.mark.parametrize("trip_only,ids,cutoff,int_cutoff", CASES)
def test_device_builder_equals_host_builder(trip_only, ids, cutoff, int_cutoff, dev):
    from gemnet_pytorch_b200 import build_batch, graph
    Rl, Zl = H.molecules(ids)
    host = build_batch(Rl, Zl, cutoff, int_cutoff, trip_only, as_torch=False)
    got = graph.build_batch_device(Rl, Zl, cutoff, int_cutoff, trip_only, dev)
    assert set(host) <= set(got)
    for k, v in host.items():
        g = got[k].cpu().numpy()
        assert g.dtype == (np.float32 if k == "R" else np.int64 if k in ("N", "Z") else np.int32), k
        assert np.array_equal(g.reshape(v.shape), v), k
    got64 = graph.build_batch_device(Rl, Zl, cutoff, int_cutoff, trip_only, dev, index_dtype=torch.int64)
    ref = build_batch(Rl, Zl, cutoff, int_cutoff, trip_only)  # the reference's format: int64 tensors
    for k, v in ref.items():
        assert got64[k].dtype == v.dtype and torch.equal(got64[k].cpu().reshape(v.shape), v), k


def test_device_builder_edge_cases(dev):
    from gemnet_pytorch_b200 import _lib, build_batch, graph
    for Rl, Zl in [([np.zeros((1, 3))], [np.array([1])]),
                   ([np.zeros((1, 3)), np.ones((1, 3))], [np.array([1]), np.array([6])]),
                   ([np.array([[0., 0, 0], [1, 0, 0]]), np.zeros((0, 3)), np.array([[0., 0, 0], [0, 1, 0], [0, 0, 1.5]])],
                    [np.array([1, 1]), np.zeros(0, int), np.array([8, 1, 1])]),
                   ([np.array([[0., 0, 0], [50, 0, 0]])], [np.array([1, 1])])]:
        for trip_only in (True, False):
            host = build_batch(Rl, Zl, 5.0, 10.0, trip_only, as_torch=False)
            got = graph.build_batch_device(Rl, Zl, 5.0, 10.0, trip_only, dev)
            for k, v in host.items():
                assert np.array_equal(got[k].cpu().numpy().reshape(v.shape), v), (k, trip_only)
    empty = graph.build_batch_device([], [], 5.0, 10.0, False, dev)
    assert all(empty[k].numel() == 0 for k in graph.TRIPLET_KEYS + graph.QUAD_KEYS)
    with pytest.raises(_lib.GemnetB200Error):  # 3 != 4 atoms
        graph._device_graph_build(graph._CudaBackend(dev), np.zeros((4, 3)), np.array([2, 1], dtype=np.int32), 5.0, 10.0, True)
    with pytest.raises(_lib.GemnetB200Error):  # no CPU fallback
        graph.build_batch_device([np.zeros((1, 3))], [np.array([1])], device="cpu")


@pytest.mark.parametrize("ids,margin", [(list(range(100, 164)), 1.05), (list(range(3)), 1.5), ([255], 1.0)])
def test_static_device_build_and_pad_equals_pad_batch(ids, margin, dev):
    from gemnet_pytorch_b200 import build_batch, graph
    Rl, Zl = H.molecules(ids)
    host = build_batch(Rl, Zl, 5.0, 10.0, True)
    cap = graph.BatchCapacity.for_batches([host], margin=margin)
    cap = graph.BatchCapacity(cap.atoms, cap.edges, cap.triplets, cap.molecules + 2)
    want, (A, E, T, M) = graph.pad_batch(host, cap, check=False)
    sb = graph.DeviceStaticBuilder(cap, 5.0, device=dev)
    for Rb, Zb, w in [(Rl, Zl, want), (Rl[:1], Zl[:1], graph.pad_batch(build_batch(Rl[:1], Zl[:1], 5.0, 10.0, True), cap,
                                                                          check=False)[0]), (Rl, Zl, want)]:
        w = {k: v.clone() for k, v in w.items()}
        sb.stage(Rb, Zb)
        sb.enqueue()
        hdr = sb.header()
        assert hdr[graph.HDR_OVERFLOW] == 0
        got = sb.inputs()
        for k in graph.SERVING_KEYS:
            assert np.array_equal(got[k].cpu().numpy().reshape(w[k].shape), w[k].numpy()), k
    small = graph.BatchCapacity(cap.atoms, 1024 if E + 6 > 1024 else 8, 4096 if T > 4096 else 8, cap.molecules)
    sb = graph.DeviceStaticBuilder(small, 5.0, device=dev)
    sb.stage(Rl, Zl)
    sb.enqueue()
    assert sb.header()[graph.HDR_OVERFLOW] != 0


@pytest.mark.parametrize("name,n", [("T", 64), ("Q", 12), ("dT", 32)])
def test_model_on_device_built_indices(name, n, dev):
    """GemNet.forward on the device-built dict == on the host-built dict (same index arrays: only the float atomics of
    the force accumulation may differ), and MoleculeServer(index_build="device") == index_build="host"."""
    from gemnet_pytorch_b200 import GemNet, build_batch, graph
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    model = GemNet(**kw)
    model.load_state_dict(go.synthetic_state_dict(kw, seed=int(fx["seed"])), strict=True)
    model.to(dev).eval()
    Rl, Zl = H.molecules(list(range(100, 100 + n)))
    host = build_batch(Rl, Zl, kw["cutoff"], kw["int_cutoff"], kw["triplets_only"])
    E0, F0 = model.predict({k: v.to(dev) for k, v in host.items()})
    E1, F1 = model.predict(graph.build_batch_device(Rl, Zl, kw["cutoff"], kw["int_cutoff"], kw["triplets_only"], dev))
    assert float((E1 - E0).abs().max()) <= 2e-6 * float(E0.abs().max())
    assert float((F1 - F0).abs().max()) <= 2e-6 * float(F0.abs().max())
    batches = [H.molecules(list(range(s, s + m))) for s, m in ((0, 12), (30, n), (60, 5), (90, n), (200, 1))]
    a = model.make_molecule_server(n, n * 30, workers=1, index_build="host")
    b = model.make_molecule_server(n, n * 30, index_build="device")
    for rep in range(2):
        got = list(b.stream(batches)) if rep else [b(*m) for m in batches]
        for (Rb, Zb), (Eb, Fb) in zip(batches, got):
            Ea, Fa = a(Rb, Zb)
            assert Eb.shape == Ea.shape and Fb.shape == Fa.shape
            assert float((Eb - Ea).abs().max()) <= 5e-6 * float(Ea.abs().max())
            assert float((Fb - Fa).abs().max()) <= 5e-6 * float(Fa.abs().max())


def test_device_build_runner_reports_a_batch_that_does_not_fit(dev):
    from gemnet_pytorch_b200 import GemNet, _lib, graph
    fx = H.load_npz("model_T.npz")
    kw = H.model_kwargs(fx)
    model = GemNet(**kw)
    model.load_state_dict(go.synthetic_state_dict(kw, seed=int(fx["seed"])), strict=True)
    model.to(dev).eval()
    Rl, Zl = H.molecules(list(range(100, 132)))
    n_e, n_t = graph.count_batch(Rl, Zl, kw["cutoff"])
    runner = model.make_serving(graph.BatchCapacity(sum(len(z) for z in Zl) + 8, (n_e // 2) // 8 * 8, n_t, 32), device_build=True)
    with pytest.raises(_lib.GemnetB200Error, match="does not fit"):
        runner.call_molecules(Rl, Zl)
    E, F = runner.call_molecules(Rl[:4], Zl[:4])  # and it keeps serving
    E0, F0 = model.make_molecule_server(4, 200, workers=1)(Rl[:4], Zl[:4])
    assert float((E - E0).abs().max()) <= 5e-6 * float(E0.abs().max())
    assert float((F - F0).abs().max()) <= 5e-6 * float(F0.abs().max())
