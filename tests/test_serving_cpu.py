"""Static-shape serving batches (gemnet_pytorch_b200.graph.pad_batch): padding must be inert.

The checker is the oracle model (CPU): energies of the real molecules and forces on the real atoms of a padded batch
are those of the unpadded batch, for the autograd-force model (T) and the direct-force model (dT)."""
import numpy as np
import pytest
import torch

from tests import helpers as H
from gemnet_pytorch_b200 import graph
from oracle import gemnet_oracle as go
from oracle import graph_oracle as gr


def _kidx(reduce_ids):
    """position of each entry inside its run of equal reduce ids (the reference's Kidx3)"""
    r = np.asarray(reduce_ids)
    start = np.r_[0, np.flatnonzero(np.diff(r)) + 1]
    run = np.repeat(start, np.diff(np.r_[start, r.size]))
    return np.arange(r.size) - run


def test_capacity_and_layout():
    b = gr.to_torch_inputs(H.oracle_batch([0, 3, 5], True))
    cap = graph.BatchCapacity.for_batches([b], margin=1.1)
    A, E, T = b["Z"].shape[0], b["id_c"].shape[0], b["id3_reduce_ca"].shape[0]
    assert cap.fits(A, E, T, 3) and cap.edges % 2 == 0
    out, sizes = graph.pad_batch(b, cap)
    assert sizes == (A, E, T, 3)
    o = {k: v.numpy() for k, v in out.items()}
    for k in ("id_c", "id_a", "id_swap", "id3_expand_ba", "id3_reduce_ca", "Z", "batch_seg"):
        n = b[k].shape[0]
        assert np.array_equal(o[k][:n], b[k].numpy()), k          # real entries keep their place
    # pads only reference pads
    assert o["id_c"][E:].min() >= A and o["id_a"][E:].min() >= A and max(o["id_c"][E:].max(), o["id_a"][E:].max()) < cap.atoms
    assert np.bincount(o["id_a"][E:]).max() <= 2 * (-(-(cap.edges - E) // 6) // max(1, (cap.atoms - A) // 3) + 1)  # degrees spread
    assert o["id_swap"][E:].min() >= E and np.array_equal(o["id_swap"][o["id_swap"][E:]], np.arange(E, cap.edges))
    assert np.array_equal(o["id_c"][o["id_swap"][E:]], o["id_a"][E:])        # the swap edge is the reverse edge
    assert o["id3_reduce_ca"][T:].min() >= E and o["id3_expand_ba"][T:].min() >= E
    assert np.all(np.diff(o["id3_reduce_ca"]) >= 0)                                 # CSR order kept
    red, exp = o["id3_reduce_ca"][T:], o["id3_expand_ba"][T:]
    assert np.array_equal(o["id_a"][red], o["id_a"][exp]) and np.all(o["id_c"][red] != o["id_c"][exp])
    assert np.all(np.diff(o["batch_seg"]) >= 0) and o["batch_seg"][A:].min() == cap.molecules
    assert o["N"].sum() == cap.atoms
    with pytest.raises(Exception):
        graph.pad_batch(b, graph.BatchCapacity(A + 2, E + 8, T, 3))


@pytest.mark.parametrize("name", ["T", "dT"])
def test_padding_is_inert(name):
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    ids = fx["mol_ids"].tolist()[:5]
    b = gr.to_torch_inputs(H.oracle_batch(ids, True))
    A, M = b["Z"].shape[0], len(ids)
    cap = graph.BatchCapacity.for_batches([b], margin=1.07)
    padded, _ = graph.pad_batch(b, cap)
    padded = dict(padded)
    padded["Kidx3"] = torch.from_numpy(_kidx(padded["id3_reduce_ca"].numpy()))
    model = go.OracleGemNet(kw, go.synthetic_state_dict(kw, seed=int(fx["seed"])), dtype=torch.float64)
    E0, F0 = model.predict(b)
    E1, F1 = model.predict(padded)
    assert E1.shape[0] == cap.molecules + 1 and F1.shape[0] == cap.atoms
    assert torch.isfinite(E1).all() and torch.isfinite(F1).all()
    assert (E1[:M] - E0).abs().max() <= 1e-12 * E0.abs().max()
    assert (F1[:A] - F0).abs().max() <= 1e-12 * F0.abs().max()


def test_serving_pool_capacity_ladder():
    """ServingPool buckets: monotone geometric ladder, every batch fits its bucket, padding overhead bounded."""
    from gemnet_pytorch_b200.model import ServingPool
    pool = ServingPool(model=None, molecules=64, max_atoms=1000, growth=1.12)
    last = None
    for e, t in [(10, 0), (1000, 5000), (5046, 52360), (6254, 61204), (8186, 95444), (20000, 400000)]:
        cap = pool.capacity_for(e, t)
        assert cap.fits(1000, e, t, 64) and cap.edges % 8 == 0 and cap.triplets % 8 == 0
        if e >= 1024:
            assert cap.edges <= 1.14 * (e + 6) + 16 and cap.triplets <= 1.14 * t + 16
        if last is not None:
            assert cap.edges >= last.edges and cap.triplets >= last.triplets
        last = cap
    assert pool.capacity_for(8186, 95444).edges == pool.capacity_for(8100, 95000).edges  # neighbours share a bucket


def test_native_padding_equals_python_padding():
    """gemnet_pad_batch_host (C ABI, one pass per array) is bit-identical to the Python specification graph.pad_batch,
    including over stale buffer contents, for tiny, empty and full-size batches."""
    from gemnet_pytorch_b200 import build_batch
    id_sets = [[0, 3, 5], list(range(10, 74)), [7], []]
    batches = [build_batch(*H.molecules(ids), 5.0, 10.0, True) if ids else None for ids in id_sets]
    cap = graph.BatchCapacity.for_batches([b for b in batches if b is not None], margin=1.05)
    bufs = graph.serving_buffers(cap)
    empty = {k: v[:0].to(torch.float32 if k == "R" else torch.int64) for k, v in bufs.items()}
    for b in batches:
        b = empty if b is None else b
        o1, o2 = graph.serving_buffers(cap), graph.serving_buffers(cap)
        for o in (o1, o2):
            for v in o.values():
                v.fill_(-7)
        assert graph._pad_batch_native(b, cap, o1) is not None                      # the native path took it
        _, s2 = graph.pad_batch({k: np.asarray(v) for k, v in b.items()}, cap, out=o2, check=False)  # numpy -> Python path
        _, s1 = graph.pad_batch(b, cap, out=o1, check=False)
        assert s1 == s2
        for k in o1:
            assert torch.equal(o1[k], o2[k]), k
    with pytest.raises(Exception, match="exceeds"):
        graph.pad_batch(batches[1], graph.BatchCapacity(8, 16, 8, 64), out=graph.serving_buffers(graph.BatchCapacity(8, 16, 8, 64)),
                        check=False)
