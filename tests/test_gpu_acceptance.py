"""Acceptance tests of the drop-in boundary with the UNMODIFIED reference's own training code on top of
gemnet_pytorch_b200.GemNet.  The reference is staged into the git-ignored oracle/_ref by oracle/build_ref.sh
(__graft_entry__.build()) and travels to the GPU box with the tree: these tests FAIL (not skip) when it is missing."""
import itertools

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _reference():
    from oracle import ref_shims
    assert ref_shims.reference_available(), (
        "the staged reference (oracle/_ref) is missing: run `bash oracle/build_ref.sh` where /root/reference exists "
        "(python -c 'import __graft_entry__ as g; g.build()' does it) before sending the tree to the GPU box")
    ref_shims.install()
    return ref_shims


def test_reference_trainer_trains_this_gemnet(dev):
    """trainer.py:325-409 (train_on_batch: predict, MAE losses, backward, scale_shared_grads, clipping, AdamW, EMA) and
    :411-470 (test_on_batch with the EMA weights: load_averaged_variables / restore_variable_backups) run unchanged on a
    direct-force GemNet-dT of this package; the loss goes down on a repeated batch."""
    _reference()
    from gemnet.training.metrics import Metrics
    from gemnet.training.trainer import Trainer

    from gemnet_pytorch_b200 import GemNet, build_batch
    from oracle import gemnet_oracle as go
    fx = H.load_npz("model_dT.npz")
    kw = H.model_kwargs(fx)
    model = GemNet(**kw)
    model.load_state_dict(go.synthetic_state_dict(kw, seed=3), strict=True)
    model.to(dev)
    trainer = Trainer(model, learning_rate=2e-4, warmup_steps=0, decay_steps=1000, ema_decay=0.9, rho_force=0.9, loss="mae")
    ids = list(range(20, 36))
    inputs = build_batch(*H.molecules(ids), kw["cutoff"], kw["int_cutoff"], True)
    g = torch.Generator().manual_seed(0)
    targets = {"E": torch.randn(len(ids), 1, generator=g), "F": torch.randn(inputs["R"].shape[0], 3, generator=g) * 0.3}
    it = itertools.repeat((inputs, targets))
    metrics = Metrics("train", trainer.tracked_metrics)
    losses = [float(trainer.train_on_batch(it, metrics)) for _ in range(12)]
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], f"loss did not decrease: {losses[0]:.4f} -> {losses[-1]:.4f}"
    # evaluation with the averaged weights and back, in the order train_seml.py:347-381 / train.ipynb use: the
    # packed-weight caches must follow the `.data` swaps of ema_decay.py (train()/eval() mode changes invalidate them)
    val = Metrics("val", trainer.tracked_metrics)
    trainer.load_averaged_variables()
    l_avg = float(trainer.test_on_batch(it, val))
    trainer.restore_variable_backups()
    l_next = float(trainer.train_on_batch(it, metrics))
    assert np.isfinite(l_avg) and np.isfinite(l_next)
    # the raw weights' evaluation equals an evaluation on a freshly built copy of the model (no stale images)
    fresh = GemNet(**kw)
    fresh.load_state_dict(model.state_dict(), strict=True)
    fresh.to(dev).eval()
    model.eval()
    dinp = {k: v.to(dev) for k, v in inputs.items()}
    E1, F1 = model(dict(dinp))
    E2, F2 = fresh(dict(dinp))
    assert torch.allclose(E1, E2, rtol=0, atol=1e-5 * float(E2.abs().max()))
    assert torch.allclose(F1, F2, rtol=0, atol=1e-5 * float(F2.abs().max()))


def test_training_the_autograd_force_model_is_refused(dev):
    """direct_forces=False training needs a double backward the kernels do not have: it must raise, not train on the
    energy alone."""
    from gemnet_pytorch_b200 import GemNet, build_batch
    fx = H.load_npz("model_T.npz")
    kw = H.model_kwargs(fx)
    model = GemNet(**kw).to(dev).train()
    inputs = {k: v.to(dev) for k, v in build_batch(*H.molecules([1, 4]), 5.0, 10.0, True).items()}
    with pytest.raises(NotImplementedError, match="double backward"):
        model(inputs)
    model.eval()
    model(inputs)


def test_molecule_matches_build_batch_and_predict(dev):
    """graph.Molecule (ase_calculator.py:23-99 / predict.ipynb cell 2 interface): get() equals build_batch of the same
    molecule, update() moves the atoms, and GemNet.predict on it equals the oracle."""
    from gemnet_pytorch_b200 import GemNet, build_batch
    from gemnet_pytorch_b200.graph import Molecule
    from oracle import gemnet_oracle as go
    from oracle import graph_oracle as gr
    fx = H.load_npz("model_T.npz")
    kw = H.model_kwargs(fx)
    sd = go.synthetic_state_dict(kw, seed=8)
    model = GemNet(**kw)
    model.load_state_dict(sd, strict=True)
    model.to(dev).eval()
    (R,), (Z,) = H.molecules([12])
    mol = Molecule(R, Z, kw["cutoff"], kw["int_cutoff"], triplets_only=True)
    a, b = mol.get(), build_batch([R], [Z], kw["cutoff"], kw["int_cutoff"], True)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    R2 = R + 0.02 * np.random.default_rng(0).normal(size=R.shape)
    mol.update(R2)
    E, F = model.predict({k: v.to(dev) for k, v in mol.get().items()})
    E0, F0 = go.OracleGemNet(kw, sd, dtype=torch.float64).predict(gr.to_torch_inputs(gr.build_batch([R2], [Z], 5.0, 10.0, True)))
    assert float((E.double() - E0).abs().max() / E0.abs().max()) < 2e-5
    assert float((F.double() - F0).abs().max() / F0.abs().max()) < 2e-5


def test_molecule_server_streams_in_order(dev):
    """GemNet.make_molecule_server: stream() (index construction on worker threads, pipelined) returns the same
    energies / forces as one-at-a-time calls, in order."""
    from gemnet_pytorch_b200 import GemNet
    from oracle import gemnet_oracle as go
    fx = H.load_npz("model_T.npz")
    kw = H.model_kwargs(fx)
    model = GemNet(**kw)
    model.load_state_dict(go.synthetic_state_dict(kw, seed=8), strict=True)
    model.to(dev).eval()
    batches = [H.molecules(list(range(s, s + 12))) for s in (0, 30, 60, 90, 120)]
    server = model.make_molecule_server(12, 12 * 30, workers=2)
    single = [server(Rl, Zl) for Rl, Zl in batches]
    streamed = list(server.stream(batches))
    assert len(streamed) == len(single)
    for (E1, F1), (E2, F2) in zip(single, streamed):
        assert torch.equal(E1, E2) and F1.shape == F2.shape
        assert float((F1 - F2).abs().max()) <= 1e-6 * float(F1.abs().max())   # force atomics: not bitwise


@pytest.mark.parametrize("name,blocks,n_mol,n_batches", [("dT", 2, 6, 2), ("dQ", 1, 3, 1)])
def test_scale_fitting_matches_reference(name, blocks, n_mol, n_batches, dev, tmp_path):
    """the fit_scaling.py protocol (fit_scaling.py:93-160 — set2fitmode, build the model, per factor: run
    batches, fit(), next) on this package's GemNet (GPU kernels + ScalingFactor.observe hooks) gives the factors the
    UNMODIFIED reference gives on its own CPU forward, factor by factor (every later factor depends on the earlier ones)."""
    import importlib
    import json
    _reference()
    from gemnet.model.gemnet import GemNet as RefGemNet
    ref_scaling = importlib.import_module("gemnet.model.layers.scaling")
    from gemnet_pytorch_b200 import GemNet, build_batch
    from gemnet_pytorch_b200 import scale_fit as scaling
    kw = dict(H.model_kwargs(H.load_npz(f"model_{name}.npz")), num_blocks=blocks)
    id_sets = [list(range(40, 40 + n_mol)), list(range(120, 120 + n_mol))][:n_batches]
    batches = [build_batch(*H.molecules(ids), kw["cutoff"], kw["int_cutoff"], kw["triplets_only"]) for ids in id_sets]

    def fit(cls, mod, path, device):
        with open(path, "w") as f:
            json.dump({"comment": "acceptance test"}, f)
        mod.AutomaticFit.set2fitmode()
        try:
            torch.manual_seed(11)
            model = cls(**dict(kw, scale_file=path)).to(device).eval()
            rounds = 0
            while not mod.AutomaticFit.fitting_completed():
                with torch.no_grad():
                    for b in batches:
                        model({k: v.to(device) for k, v in b.items()})
                mod.AutomaticFit.activeVar.fit()
                rounds += 1
                assert rounds < 100
        finally:
            mod.AutomaticFit.reset()
            mod.AutomaticFit.queue = None
            mod.AutomaticFit.fitting_mode = False
        with open(path) as f:
            return json.load(f), rounds, model

    want, r_ref, _ = fit(RefGemNet, ref_scaling, str(tmp_path / "ref.json"), torch.device("cpu"))
    got, r_mine, model = fit(GemNet, scaling, str(tmp_path / "mine.json"), dev)
    assert r_mine == r_ref == len(want) - 1 and list(got) == list(want)   # same factors in the same (creation) order
    worst = max(abs(got[k] - want[k]) / abs(want[k]) for k in want if k != "comment")
    print(f"[{name}] {r_mine} factors fitted, worst relative difference {worst:.2e}")
    assert worst <= 2e-4, worst
    # the fitted factors are what the fused engine path uses afterwards, and a model built from the file loads them
    assert not scaling.fitting_active()
    again = GemNet(**dict(kw, scale_file=str(tmp_path / "mine.json"))).to(dev).eval()
    again.load_state_dict(model.state_dict(), strict=True)
    b = {k: v.to(dev) for k, v in batches[0].items()}
    E1, F1 = model.predict(dict(b))
    E2, F2 = again.predict(dict(b))
    assert torch.allclose(E1, E2, rtol=0, atol=1e-5 * float(E2.abs().max()))
