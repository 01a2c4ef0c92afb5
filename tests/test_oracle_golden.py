"""CPU: pins oracle/ against the fixtures generated from the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import basis_oracle as bo
from oracle import gemnet_oracle as go
from oracle import graph_oracle as gr
from tests import helpers as H


def test_numba_docstring_known_answers():
    # data_container.py:527-533 and :553-556
    assert gr.repeat_blocks([1, 3, 2], [3, 2, 3]).tolist() == [0, 0, 0, 1, 2, 3, 1, 2, 3, 4, 5, 4, 5, 4, 5]
    assert gr.repeat_blocks([0, 3, 2], [3, 2, 3]).tolist() == [0, 1, 2, 0, 1, 2, 3, 4, 3, 4, 3, 4]
    assert gr.repeat_blocks([2, 3, 2], [2, 0, 2]).tolist() == [0, 1, 0, 1, 5, 6, 5, 6]
    assert gr.ragged_range([1, 3, 2]).tolist() == [0, 0, 1, 2, 0, 1]


@pytest.mark.parametrize("name,trip_only", [("T", True), ("Q", False)])
def test_graph_oracle_matches_reference_fixture(name, trip_only):
    fx = H.load_npz(f"graph_{name}.npz")
    mine = H.oracle_batch(fx["mol_ids"].tolist(), trip_only)
    exact = ["id_c", "id_a", "id_undir", "id_swap", "batch_seg"]
    if not trip_only:
        exact += ["id4_int_a", "id4_int_b", "id4_reduce_intm_ca", "id4_expand_intm_db", "id4_reduce_intm_ab",
                  "id4_expand_intm_ab"]
    for k in exact:
        assert np.array_equal(fx[k], mine[k].astype(np.int32)), k
    # sorted arrays: identical as multisets per segment (and here, with the same numpy argsort, element-wise)
    a = H.canonical_segments(fx["id3_reduce_ca"], fx["id3_expand_ba"])
    b = H.canonical_segments(mine["id3_reduce_ca"], mine["id3_expand_ba"])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(fx["Kidx3"], mine["Kidx3"])
    if not trip_only:
        ks = ["id4_reduce_ca", "id4_expand_db", "id4_reduce_cab", "id4_expand_abd"]
        a = H.canonical_segments(*[fx[k] for k in ks])
        b = H.canonical_segments(*[mine[k] for k in ks])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert np.array_equal(fx["Kidx4"], mine["Kidx4"])


def test_index_invariants():
    # the commented-out node identities of data_container.py:341-344,393-405
    d = H.oracle_batch(list(range(6)), False)
    s, t = d["id_c"], d["id_a"]
    assert np.all(t[d["id3_reduce_ca"]] == t[d["id3_expand_ba"]])
    assert np.all(s[d["id3_reduce_ca"]] != s[d["id3_expand_ba"]])
    assert np.all(s[d["id_swap"]] == t) and np.all(t[d["id_swap"]] == s)
    n = len(s) // 2
    assert np.all(t[:n] < s[:n])
    assert np.all(np.diff(d["id3_reduce_ca"]) >= 0) and np.all(np.diff(d["id4_reduce_ca"]) >= 0)
    c, a = s[d["id4_reduce_ca"]], t[d["id4_reduce_ca"]]
    b, dd = t[d["id4_expand_db"]], s[d["id4_expand_db"]]
    it, is_ = d["id4_int_a"], d["id4_int_b"]
    assert np.all(c == s[d["id4_reduce_intm_ca"]][d["id4_reduce_cab"]])
    assert np.all(a == t[d["id4_reduce_intm_ca"]][d["id4_reduce_cab"]])
    assert np.all(a == it[d["id4_reduce_intm_ab"]][d["id4_reduce_cab"]])
    assert np.all(a == it[d["id4_expand_intm_ab"]][d["id4_expand_abd"]])
    assert np.all(b == is_[d["id4_reduce_intm_ab"]][d["id4_reduce_cab"]])
    assert np.all(b == is_[d["id4_expand_intm_ab"]][d["id4_expand_abd"]])
    assert np.all(b == t[d["id4_expand_intm_db"]][d["id4_expand_abd"]])
    assert np.all(dd == s[d["id4_expand_intm_db"]][d["id4_expand_abd"]])
    assert np.all(c != b) and np.all(a != dd) and np.all(c != dd)


def test_empty_and_tiny_graphs():
    far = [np.array([[0.0, 0, 0], [50.0, 0, 0]])]
    d = gr.build_batch(far, [np.array([1, 1])], 5.0, 10.0, False)
    assert len(d["id_c"]) == 0 and len(d["id4_reduce_ca"]) == 0
    pair = [np.array([[0.0, 0, 0], [1.0, 0, 0]])]
    d = gr.build_batch(pair, [np.array([1, 8])], 5.0, 10.0, False)
    assert len(d["id_c"]) == 2 and len(d["id3_reduce_ca"]) == 0 and len(d["id4_reduce_ca"]) == 0


def test_basis_ops_match_reference_fixture():
    fx = H.load_npz("ops.npz")
    D = torch.tensor(fx["D_ca"])
    rbf = bo.bessel_rbf(D, torch.tensor(fx["frequencies"]).double(), 5.0, 5)
    np.testing.assert_allclose(rbf.numpy(), fx["rbf64"], rtol=1e-10, atol=1e-12)
    rad = bo.radial_sph_bessel(D, 7, 6, 5.0, 5)  # (E,7,6)
    np.testing.assert_allclose(rad.permute(1, 0, 2).numpy(), fx["cbf3_rbf_env64"], rtol=0, atol=1e-11)
    sph = bo.legendre_y_l0(torch.tensor(fx["Angles3"]), 7)
    np.testing.assert_allclose(sph.numpy(), fx["cbf3_sph64"], rtol=0, atol=1e-12)
    ylm = bo.real_sph_harm_all(torch.tensor(fx["Phi_cab"]), torch.tensor(fx["Theta_cabd"]), 7)
    np.testing.assert_allclose(ylm.numpy(), fx["sbf_sph64"], rtol=0, atol=1e-11)
    deg = torch.arange(7) * 2 + 1
    rad49 = torch.repeat_interleave(rad, deg, dim=1).permute(1, 0, 2)
    np.testing.assert_allclose(rad49.numpy(), fx["sbf_rbf_env64"], rtol=0, atol=1e-11)
    # the reference's own float32 evaluation deviates from its float64 one by cancellation noise:
    assert np.abs(fx["cbf3_rbf_env32"] - fx["cbf3_rbf_env64"]).max() < 2e-3


def test_geometry_and_bilinear_match_reference_fixture():
    fx = H.load_npz("ops.npz")
    ids = fx["mol_ids"].tolist()
    b = gr.to_torch_inputs(H.oracle_batch(ids, False))
    R = b["R"].double()
    D, V = go.edge_vectors(R, b["id_c"], b["id_a"])
    np.testing.assert_allclose(D.numpy(), fx["D_ca"], atol=1e-13)
    np.testing.assert_allclose(V.numpy(), fx["V_ca"], atol=1e-13)
    A3 = go.triplet_angles(R, b["id_c"], b["id_a"], b["id3_reduce_ca"], b["id3_expand_ba"])
    np.testing.assert_allclose(A3.numpy(), fx["Angles3"], atol=1e-12)
    pc, pa, th = go.quad_angles(R, b["id_c"], b["id_a"], b["id4_int_b"], b["id4_int_a"], b["id4_expand_abd"],
                                b["id4_reduce_cab"], b["id4_expand_intm_db"], b["id4_reduce_intm_ca"],
                                b["id4_expand_intm_ab"], b["id4_reduce_intm_ab"])
    np.testing.assert_allclose(pc.numpy(), fx["Phi_cab"], atol=1e-12)
    np.testing.assert_allclose(pa.numpy(), fx["Phi_abd"], atol=1e-12)
    np.testing.assert_allclose(th.numpy(), fx["Theta_cabd"], atol=1e-10)
    for tag, idr, kidx, sphk in (("trip", b["id3_reduce_ca"], b["Kidx3"], "cbf3_sph64"),
                                 ("quad", b["id4_reduce_ca"], b["Kidx4"], "sbf_sph64")):
        nE = len(b["id_c"])
        sph = go.pad_by_segment(torch.tensor(fx[sphk]), idr, kidx, nE)
        basis = (torch.tensor(fx[f"{tag}_rbf_W1"]), sph.transpose(1, 2))
        out = go.efficient_bilinear(basis, torch.tensor(fx[f"{tag}_m"]).double(), idr, kidx,
                                    torch.tensor(fx[f"{tag}_Wbil"]))
        np.testing.assert_allclose(out.numpy(), fx[f"{tag}_out64"], rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name", ["T", "Q", "dT", "dQ"])
def test_model_oracle_matches_reference_fixture(name):
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    ids = fx["mol_ids"].tolist()
    if name in ("Q", "dQ"):
        ids = ids[:6]  # keep the CPU suite short; energies/forces of a molecule do not depend on the rest of the batch
    inputs = gr.to_torch_inputs(H.oracle_batch(ids, kw["triplets_only"]))
    nat = inputs["R"].shape[0]
    sd = go.synthetic_state_dict(kw, seed=int(fx["seed"]))
    # float64 oracle vs float64 reference: algebraic identity
    E64, F64 = go.OracleGemNet(kw, sd, dtype=torch.float64).predict(inputs)
    assert np.abs(E64.numpy() - fx["E64"][:len(ids)]).max() <= 1e-10 * np.abs(fx["E64"]).max()
    assert np.abs(F64.numpy() - fx["F64"][:nat]).max() <= 1e-10 * np.abs(fx["F64"]).max()
    # float32 oracle vs float32 reference: the north_star tolerance (1e-4 relative) with a wide margin
    E32, F32 = go.OracleGemNet(kw, sd, dtype=torch.float32).predict(inputs)
    refE, refF = fx["E32"][:len(ids)], fx["F32"][:nat]
    assert np.abs(E32.numpy() - refE).max() <= 1e-5 * np.abs(refE).max() + 1e-5
    assert np.abs(F32.numpy() - refF).max() <= 1e-5 * np.abs(refF).max()
