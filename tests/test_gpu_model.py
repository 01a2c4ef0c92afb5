"""GPU parity of the whole hot path behind the reference boundary: GemNet.forward energies and forces against
  (a) the committed fixtures produced by the unmodified reference (tests/golden/model_*.npz, synthetic weights),
  (b) the CPU oracle on other batches (incl. COLL molecules 64-255, larger cutoff, edge cases),
  (c) the reference's outputs with the shipped pretrained checkpoints (tracked fixtures; the checkpoints themselves come
      with the staged reference, oracle/_ref).
Gate (north_star): |dE| <= 1e-4 |E|max, |dF| <= 1e-4 |F|max per batch (SURVEY §6 last row); we assert 2e-5.
"""
import os

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

from oracle import gemnet_oracle as go  # noqa: E402
from oracle import graph_oracle as gr  # noqa: E402

TOL = 2e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def make_model(kw, sd, dev):
    from gemnet_pytorch_b200 import GemNet
    model = GemNet(**kw)
    model.load_state_dict(sd, strict=True)
    return model.to(dev).eval()


def to_dev(inputs, dev):
    return {k: v.to(dev) for k, v in inputs.items()}


def check(E, F, E_ref, F_ref, tol=TOL):
    E, F = E.detach().double().cpu().numpy(), F.detach().double().cpu().numpy()
    E_ref, F_ref = np.asarray(E_ref, dtype=np.float64), np.asarray(F_ref, dtype=np.float64)
    assert E.shape == E_ref.shape and F.shape == F_ref.shape
    dE = np.abs(E - E_ref).max() / max(np.abs(E_ref).max(), 1e-30)
    dF = np.abs(F - F_ref).max() / max(np.abs(F_ref).max(), 1e-30)
    assert dE <= tol, f"energy rel err {dE:.3e}"
    assert dF <= tol, f"force rel err {dF:.3e}"
    return dE, dF


@pytest.mark.parametrize("name", ["T", "Q", "dT", "dQ"])
def test_forward_matches_reference_fixture(name, dev):
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    sd = go.synthetic_state_dict(kw, seed=int(fx["seed"]))
    model = make_model(kw, sd, dev)
    from gemnet_pytorch_b200 import build_batch
    inputs = build_batch(*H.molecules(fx["mol_ids"].tolist()), kw["cutoff"], kw["int_cutoff"], kw["triplets_only"])
    E, F = model(to_dev(inputs, dev))
    # float64 reference = truth; the reference's own float32 run is the yardstick
    dE, dF = check(E, F, fx["E64"], fx["F64"])
    refE = np.abs(fx["E32"] - fx["E64"]).max() / np.abs(fx["E64"]).max()
    refF = np.abs(fx["F32"] - fx["F64"]).max() / np.abs(fx["F64"]).max()
    print(f"[{name}] ours vs ref64: dE {dE:.2e} dF {dF:.2e} | ref32 vs ref64: dE {refE:.2e} dF {refF:.2e}")
    check(E, F, fx["E32"], fx["F32"])


@pytest.mark.parametrize("fixture", ["model_Q_b64", "model_Q_c8", "model_T_c8_clusters"])
def test_forward_matches_reference_fixture_other_configs(fixture, dev):
    """BASELINE configs[2] (GemNet-Q, COLL batch 64) and configs[4] (cutoff 8 A: GemNet-Q on COLL molecules, GemNet-T on
    the synthetic 18-atom clusters with 4 896 triplets each) against outputs of the UNMODIFIED reference
    (oracle/make_golden_configs.py): float64 = truth, the reference's own float32 run = yardstick."""
    fx = H.load_npz(f"{fixture}.npz")
    kw = H.model_kwargs(fx)
    sd = go.synthetic_state_dict(kw, seed=int(fx["seed"]))
    model = make_model(kw, sd, dev)
    from gemnet_pytorch_b200 import build_batch
    if "mol_ids" in fx:
        Rl, Zl = H.molecules(fx["mol_ids"].tolist())
    else:
        off = np.concatenate([[0], np.cumsum(fx["N"])])
        Rl = [fx["R"][off[i]:off[i + 1]] for i in range(len(fx["N"]))]
        Zl = [fx["Z"][off[i]:off[i + 1]] for i in range(len(fx["N"]))]
    inputs = build_batch(Rl, Zl, kw["cutoff"], kw["int_cutoff"], kw["triplets_only"])
    E, F = model(to_dev(inputs, dev))
    dE, dF = check(E, F, fx["E64"], fx["F64"])
    refE = np.abs(fx["E32"] - fx["E64"]).max() / np.abs(fx["E64"]).max()
    refF = np.abs(fx["F32"] - fx["F64"]).max() / np.abs(fx["F64"]).max()
    print(f"[{fixture}] ours vs ref64: dE {dE:.2e} dF {dF:.2e} | ref32 vs ref64: dE {refE:.2e} dF {refF:.2e}")


@pytest.mark.parametrize("name", ["T", "Q"])
def test_forward_accepts_reference_datacontainer_order(name, dev):
    """Inputs in the reference DataContainer's own (argsort-scrambled) within-segment order give the same result."""
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    model = make_model(kw, go.synthetic_state_dict(kw, seed=1234), dev)
    ids = fx["mol_ids"].tolist()[:10]
    inputs = gr.to_torch_inputs(H.oracle_batch(ids, kw["triplets_only"]))
    E, F = model(to_dev(inputs, dev))
    nat = inputs["R"].shape[0]
    check(E, F, fx["E64"][:len(ids)], fx["F64"][:nat])
    inputs.pop("N")  # the reference forward does not need N (gemnet.py:578 uses batch_seg.max())
    inputs = {k: v for k, v in inputs.items() if not k.startswith("_")}
    E2, F2 = model(to_dev(inputs, dev))
    check(E2, F2, fx["E64"][:len(ids)], fx["F64"][:nat])


@pytest.mark.parametrize("name,ids,cutoff", [("T", list(range(100, 228)), 5.0), ("T", list(range(200, 230)), 8.0),
                                             ("Q", list(range(230, 238)), 5.0), ("dT", list(range(150, 200)), 6.0)])
def test_forward_matches_oracle_other_batches(name, ids, cutoff, dev):
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    kw["cutoff"] = cutoff
    sd = go.synthetic_state_dict(kw, seed=99)
    model = make_model(kw, sd, dev)
    inputs = gr.to_torch_inputs(H.oracle_batch(ids, kw["triplets_only"], cutoff, kw["int_cutoff"]))
    E, F = model(to_dev(inputs, dev))
    E0, F0 = go.OracleGemNet(kw, sd, dtype=torch.float64).predict(inputs)
    check(E, F, E0.numpy(), F0.numpy())


def test_edge_cases(dev):
    fx = H.load_npz("model_T.npz")
    kw = H.model_kwargs(fx)
    sd = go.synthetic_state_dict(kw, seed=5)
    model = make_model(kw, sd, dev)
    oracle = go.OracleGemNet(kw, sd, dtype=torch.float64)
    from gemnet_pytorch_b200 import build_batch
    # single molecule; a diatomic (edges but no triplets) inside a batch; molecule 7 (4 atoms)
    di_R, di_Z = np.array([[0.0, 0, 0], [1.1, 0.1, 0]]), np.array([8, 1])
    Rl, Zl = H.molecules([7])
    for R_list, Z_list in (([Rl[0]], [Zl[0]]), ([di_R, Rl[0], di_R + 0.3], [di_Z, Zl[0], di_Z])):
        inputs = build_batch(R_list, Z_list, 5.0, 10.0, True)
        E, F = model(to_dev(inputs, dev))
        E0, F0 = oracle.predict(gr.to_torch_inputs(gr.build_batch(R_list, Z_list, 5.0, 10.0, True)))
        check(E, F, E0.numpy(), F0.numpy())
    # requires_grad side effect (gemnet.py:493-494,613) and predict() returning CPU tensors
    inputs = to_dev(build_batch(Rl, Zl, 5.0, 10.0, True), dev)
    model(inputs)
    assert inputs["R"].requires_grad is False
    E, F = model.predict(inputs)
    assert E.device.type == "cpu" and F.device.type == "cpu" and F.shape == (4, 3)
    # multi-target head
    kw2 = dict(kw, num_targets=2)
    sd2 = go.synthetic_state_dict(kw2, seed=5)
    m2 = make_model(kw2, sd2, dev)
    E, F = m2(to_dev(build_batch(Rl, Zl, 5.0, 10.0, True), dev))
    E0, F0 = go.OracleGemNet(kw2, sd2, dtype=torch.float64).predict(gr.to_torch_inputs(H.oracle_batch([7], True)))
    assert F.shape == (4, 2, 3)
    check(E, F, E0.numpy(), F0.numpy())
    # non-extensive (mean) read-out
    kw3 = dict(kw, extensive=False)
    m3 = make_model(kw3, sd, dev)
    ids = [3, 5, 9]
    E, F = m3(to_dev(build_batch(*H.molecules(ids), 5.0, 10.0, True), dev))
    E0, F0 = go.OracleGemNet(kw3, sd, dtype=torch.float64).predict(gr.to_torch_inputs(H.oracle_batch(ids, True)))
    check(E, F, E0.numpy(), F0.numpy())
    # unsorted reduce indices are rejected loudly
    bad = to_dev(build_batch(*H.molecules(ids), 5.0, 10.0, True), dev)
    bad["id3_reduce_ca"] = bad["id3_reduce_ca"].flip(0)
    from gemnet_pytorch_b200._lib import GemnetB200Error
    with pytest.raises(GemnetB200Error):
        model(bad)


@pytest.mark.parametrize("M", ["T", "Q"])
def test_pretrained_checkpoint_parity(M, dev):
    """The SHIPPED checkpoints (pretrained/GemNet-{T,Q}/model.pth, strict-loaded from the staged reference,
    oracle/build_ref.sh) against the unmodified reference's own outputs with them (tests/golden/model_*_pretrained.npz,
    oracle/make_golden_pretrained.py).  Fails — does not skip — when the staged reference did not travel with the tree."""
    from oracle import ref_shims
    ck = os.path.join(ref_shims.REFERENCE_ROOT, "pretrained", f"GemNet-{M}", "model.pth")
    assert os.path.exists(ck), (
        f"{ck} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists "
        "(it stages the reference and its checkpoints into the git-ignored oracle/_ref, which travels to the GPU box)")
    fx = H.load_npz(f"model_{M}_pretrained.npz")
    kw = H.model_kwargs(fx)
    model = make_model(kw, torch.load(ck, map_location="cpu"), dev)
    from gemnet_pytorch_b200 import build_batch
    inputs = build_batch(*H.molecules(fx["mol_ids"].tolist()), kw["cutoff"], kw["int_cutoff"], kw["triplets_only"])
    E, F = model(to_dev(inputs, dev))
    dE, dF = check(E, F, fx["E64"], fx["F64"])
    check(E, F, fx["E32"], fx["F32"])
    print(f"[pretrained {M}] dE {dE:.2e} dF {dF:.2e}")


def test_training_step_direct_forces(dev):
    """config 4 plumbing: weight gradients of a direct-force model against the oracle's autograd."""
    fx = H.load_npz("model_dT.npz")
    kw = H.model_kwargs(fx)
    sd = go.synthetic_state_dict(kw, seed=11)
    model = make_model(kw, sd, dev).train()
    ids = list(range(12))
    inputs = gr.to_torch_inputs(H.oracle_batch(ids, True))
    E, F = model(to_dev(inputs, dev))
    loss = E.sum() + 0.3 * (F ** 2).sum()
    loss.backward()
    orc = go.OracleGemNet(kw, sd, dtype=torch.float64)
    for k in orc.p:
        orc.p[k].requires_grad_(orc.p[k].dtype.is_floating_point and "scale_factor" not in k)
    E0, F0 = orc.forward(inputs)
    (E0.sum() + 0.3 * (F0 ** 2).sum()).backward()
    checked = 0
    for name, p in model.named_parameters():
        if not p.requires_grad or p.grad is None:
            continue
        ref = orc.p[name].grad
        if ref is None:
            ref = orc.p[name.replace(".linear.weight", ".weight")].grad
        if ref is None:
            continue
        err = (p.grad.double().cpu() - ref).abs().max() / ref.abs().max().clamp(min=1e-12)
        assert err < 2e-4, f"{name}: {err:.2e}"
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("name", ["T", "Q", "dT", "dQ"])
def test_engine_matches_autograd_path(name, dev):
    """eval() runs on the tape engine (engine.py); it must agree with the torch.autograd evaluation of the same
    kernels to float32 round-off (summation order of gradient accumulation differs)."""
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    sd = go.synthetic_state_dict(kw, seed=21)
    model = make_model(kw, sd, dev)
    ids = fx["mol_ids"].tolist()[:8]
    from gemnet_pytorch_b200 import build_batch
    inputs = to_dev(build_batch(*H.molecules(ids), kw["cutoff"], kw["int_cutoff"], kw["triplets_only"]), dev)
    assert model.use_engine
    E1, F1 = model(inputs)
    E2, F2 = model.forward_autograd(inputs)
    assert inputs["R"].requires_grad is False
    # engine = chain-fused tcgen05 dense stacks (exact operand split, compensated accumulation; ScaledSiLU through
    # ex2.approx / rcp.approx, ~3e-7 per activation), autograd path = cuBLAS fp32 GEMMs + expf kernels
    dE, dF = (E1 - E2).abs().max() / E2.abs().max(), (F1 - F2).abs().max() / F2.abs().max()
    print(f"[{name}] engine vs autograd path: dE {dE:.2e} dF {dF:.2e}")
    assert dE <= 1e-5 and dF <= 1.5e-5
    model.dense_algo = 0  # same engine on the cuBLAS GEMMs: only the gradient-accumulation order differs
    E3, F3 = model(inputs)
    model.dense_algo = 1
    assert (E3 - E2).abs().max() <= 1e-6 * E2.abs().max()
    assert (F3 - F2).abs().max() <= 2e-6 * F2.abs().max()
    # multi-target through the engine
    if name == "T":
        kw2 = dict(kw, num_targets=2)
        m2 = make_model(kw2, go.synthetic_state_dict(kw2, seed=21), dev)
        Ea, Fa = m2(inputs)
        Eb, Fb = m2.forward_autograd(inputs)
        assert Fa.shape == Fb.shape == (inputs["R"].shape[0], 2, 3)
        assert (Fa - Fb).abs().max() <= 5e-6 * Fb.abs().max() and (Ea - Eb).abs().max() <= 3e-6 * Eb.abs().max()


@pytest.mark.parametrize("name", ["T", "dT"])
def test_static_shape_serving_matches_fixture_and_eager(name, dev):
    """ServingGemNet: one captured graph (H2D, index plan, E+F, D2H) serves batches of different topology padded
    to one capacity; results = the reference fixture (first batch) and the eager path (all batches)."""
    from gemnet_pytorch_b200 import build_batch, graph
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    model = make_model(kw, go.synthetic_state_dict(kw, seed=int(fx["seed"])), dev)
    id_sets = [fx["mol_ids"].tolist(), list(range(40, 70)), list(range(100, 164)), [7]]
    batches = [build_batch(*H.molecules(ids), kw["cutoff"], kw["int_cutoff"], True) for ids in id_sets]
    cap = graph.BatchCapacity.for_batches(batches, margin=1.03)
    serve = model.make_serving(cap)
    for rep in range(2):  # the second pass runs over stale pads of other batches
        for j, b in enumerate(batches):
            E, F = serve(b)
            assert E.shape[0] == len(id_sets[j]) and F.shape[0] == b["Z"].shape[0]
            assert torch.isfinite(E).all() and torch.isfinite(F).all()
            E0, F0 = model.predict(to_dev(b, dev))
            check(E, F, E0.numpy(), F0.numpy(), tol=5e-6)
            if j == 0:
                check(E, F, fx["E64"], fx["F64"])
    with pytest.raises(Exception):
        serve(build_batch(*H.molecules(list(range(0, 200))), kw["cutoff"], kw["int_cutoff"], True))
    bad = dict(batches[1])
    bad["id3_reduce_ca"] = batches[1]["id3_reduce_ca"].flip(0)  # not grouped as the DataContainer produces it
    with pytest.raises(Exception, match="sorted"):
        serve(bad)
    E, F = serve(batches[1])  # and the runner keeps working afterwards
    E0, F0 = model.predict(to_dev(batches[1], dev))
    check(E, F, E0.numpy(), F0.numpy(), tol=5e-6)


@pytest.mark.parametrize("name", ["T", "dT", "Q"])
def test_two_lane_evaluation_matches_single_stream(name, dev, monkeypatch):
    """engine lanes: output blocks on a second stream (forward and backward) — same numbers as the one-stream run,
    eagerly (GEMNET_B200_LANES=1) and inside a captured graph (the default there)."""
    from gemnet_pytorch_b200 import build_batch
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    model = make_model(kw, go.synthetic_state_dict(kw, seed=int(fx["seed"])), dev)
    ids = fx["mol_ids"].tolist()[:(6 if name == "Q" else 32)]
    inputs = to_dev(build_batch(*H.molecules(ids), kw["cutoff"], kw["int_cutoff"], kw["triplets_only"]), dev)
    monkeypatch.setenv("GEMNET_B200_LANES", "0")
    E0, F0 = model.predict(inputs)
    monkeypatch.setenv("GEMNET_B200_LANES", "1")
    for _ in range(3):
        E1, F1 = model.predict(inputs)
        check(E1, F1, E0.numpy(), F0.numpy(), tol=5e-6)
    monkeypatch.setenv("GEMNET_B200_LANES", "capture")
    graphed = model.make_graphed(inputs)
    for _ in range(3):
        E2, F2 = graphed()
        torch.cuda.synchronize()
        check(E2, F2, E0.numpy(), F0.numpy(), tol=5e-6)


def test_serving_pool_buckets(dev):
    """ServingPool: batches of different size run in different capacity buckets (one captured graph each)."""
    from gemnet_pytorch_b200 import build_batch
    fx = H.load_npz("model_T.npz")
    kw = H.model_kwargs(fx)
    model = make_model(kw, go.synthetic_state_dict(kw, seed=int(fx["seed"])), dev)
    id_sets = [list(range(0, 8)), list(range(8, 40)), list(range(40, 48)), list(range(100, 164))]
    batches = [build_batch(*H.molecules(ids), kw["cutoff"], kw["int_cutoff"], True) for ids in id_sets]
    pool = model.make_serving_pool(64, max(int(b["Z"].shape[0]) for b in batches), growth=1.25)
    for rep in range(2):
        for ids, b in zip(id_sets, batches):
            E, F = pool(b)
            assert E.shape[0] == len(ids) and F.shape[0] == b["Z"].shape[0]
            E0, F0 = model.predict(to_dev(b, dev))
            check(E, F, E0.numpy(), F0.numpy(), tol=5e-6)
    assert 2 <= len(pool.runners) <= 4
