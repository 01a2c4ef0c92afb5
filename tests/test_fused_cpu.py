"""Host logic of the chain-fused engine (gemnet_pytorch_b200/fused.py + engine.Tape) checked WITHOUT a GPU: every kernel
launch goes to a torch stand-in (tests/chain_emulator.py, float64), and the fused InteractionBlock / OutputBlock
evaluation — forward AND the hand-derived backward programs with their lazy gradient summands — must reproduce the
oracle's blocks (oracle/gemnet_oracle.py, torch autograd) to 1e-10.  The kernel itself is checked against the same
interpreter in tests/test_gpu_chain.py."""
import types

import numpy as np
import pytest
import torch

from oracle import gemnet_oracle as go
from oracle import graph_oracle as gr
from tests import chain_emulator as ce
from tests import helpers as H


def csr_by_key(ids, n):
    ids = np.asarray(ids)
    perm = np.argsort(ids, kind="stable")
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(ids, minlength=n))])
    return torch.tensor(rowptr, dtype=torch.int64), torch.tensor(perm, dtype=torch.int64)


def csr_sorted(ids, n):
    return torch.tensor(np.concatenate([[0], np.cumsum(np.bincount(np.asarray(ids), minlength=n))]), dtype=torch.int64)


def cpu_plan(inp, triplets_only):
    p = types.SimpleNamespace()
    p.id_c, p.id_a, p.id_swap = inp["id_c"], inp["id_a"], inp["id_swap"]
    p.n_edges, p.n_atoms = int(p.id_c.numel()), int(inp["Z"].numel())
    p.atom_rowptr, p.atom_perm = csr_by_key(p.id_a, p.n_atoms)
    p.src_rowptr, p.src_perm = csr_by_key(p.id_c, p.n_atoms)
    p.swap_rowptr, p.swap_perm = csr_by_key(p.id_swap, p.n_edges)
    p.id3_expand, p.id3_reduce = inp["id3_expand_ba"], inp["id3_reduce_ca"]
    p.rowptr3 = csr_sorted(p.id3_reduce, p.n_edges)
    if not triplets_only:
        p.id4_expand_intm_db, p.id4_expand_abd = inp["id4_expand_intm_db"], inp["id4_expand_abd"]
        p.intm_db_rowptr, p.intm_db_perm = csr_by_key(p.id4_expand_intm_db, p.n_edges)
        p.rowptr4 = csr_sorted(inp["id4_reduce_ca"], p.n_edges)
    return p


def setup(name, mol_ids, seed):
    from gemnet_pytorch_b200 import GemNet
    fx = H.load_npz(f"model_{name}.npz")
    kw = H.model_kwargs(fx)
    sd = go.synthetic_state_dict(kw, seed=seed, dtype=torch.float64)
    model = GemNet(**kw).double()
    model.load_state_dict(sd, strict=True)
    model.eval()
    oracle = go.OracleGemNet(kw, sd, dtype=torch.float64)
    inp = gr.to_torch_inputs(H.oracle_batch(mol_ids, kw["triplets_only"]))
    return kw, model, oracle, inp


@pytest.mark.parametrize("name", ["T", "Q", "dT"])
def test_fused_blocks_match_oracle_blocks(name):
    from gemnet_pytorch_b200 import engine
    kw, model, oracle, inp = setup(name, [1, 4, 7], seed=11)
    tri = kw["triplets_only"]
    plan = cpu_plan(inp, tri)
    E, A, T = plan.n_edges, plan.n_atoms, int(plan.id3_reduce.numel())
    g = torch.Generator().manual_seed(3)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)  # noqa: E731
    S = kw["num_spherical"]
    leaves = {"h": rnd(A, 128), "m": rnd(E, 128), "rbf3": rnd(E, 16), "rbf_h": rnd(E, 16), "rbf_out": rnd(E, 16),
              "sph3": rnd(T, S), "rbf_W1_3": rnd(E, 16, S) * 0.3}
    if not tri:
        M, Q = int(inp["id4_expand_intm_db"].numel()), int(inp["id4_expand_abd"].numel())
        leaves.update({"rbf4": rnd(E, 16), "cbf4": rnd(M, 16), "ylm": rnd(Q, S * S) * 0.3,
                       "rbf_W1_4": rnd(E, 32, S * S) * 0.1})
    Gh, Gm, GE = rnd(A, 128), rnd(E, 128), rnd(A, 1)
    GF = rnd(E, 1)
    direct = kw["direct_forces"]

    # ---- oracle: blocks through torch autograd
    L = {k: v.clone().requires_grad_(True) for k, v in leaves.items()}
    gd = {"id_swap": inp["id_swap"], "id3_expand_ba": inp["id3_expand_ba"], "id3_reduce_ca": inp["id3_reduce_ca"],
          "Kidx3": inp["Kidx3"], "id_c": inp["id_c"], "id_a": inp["id_a"], "rbf3": L["rbf3"], "rbf_h": L["rbf_h"],
          "cbf3": (L["rbf_W1_3"], go.pad_by_segment(L["sph3"], inp["id3_reduce_ca"], inp["Kidx3"], E).transpose(1, 2))}
    if not tri:
        gd.update(rbf4=L["rbf4"], cbf4=L["cbf4"], Kidx4=inp["Kidx4"], id4_reduce_ca=inp["id4_reduce_ca"],
                  id4_expand_intm_db=inp["id4_expand_intm_db"], id4_expand_abd=inp["id4_expand_abd"],
                  sbf4=(L["rbf_W1_4"], go.pad_by_segment(L["ylm"], inp["id4_reduce_ca"], inp["Kidx4"], E).transpose(1, 2)))
    h1, m1 = oracle.interaction_block("int_blocks.0", L["h"], L["m"], gd)
    h2, m2 = oracle.interaction_block("int_blocks.1", h1, m1, gd)
    E0, F0 = oracle.output_block("out_blocks.1", h1, m1, L["rbf_out"], inp["id_a"])
    E1, F1 = oracle.output_block("out_blocks.2", h2, m2, L["rbf_out"], inp["id_a"])
    Esum = E0 + E1
    loss = (h2 * Gh).sum() + (m2 * Gm).sum() + (Esum * GE).sum()
    if direct:
        loss = loss + ((F0 + F1) * GF).sum()
    ref_grads = dict(zip(L, torch.autograd.grad(loss, list(L.values()))))

    # ---- fused engine on the emulated kernels
    with ce.patched(), torch.no_grad():
        tape = engine.Tape(enabled=True, streams=None)
        X = {k: v.clone() for k, v in leaves.items()}
        for v in X.values():
            tape.require(v)
        run = engine._Run(model, tape, plan)
        assert run.fuse and run.fuse_out, "the published configuration must take the fused path"
        basis = {k: X[k] for k in X if k not in ("h", "m", "rbf_out")}
        a1, b1 = run.interaction_block(model.int_blocks[0], X["h"], X["m"], basis, 0)
        Ea, Fa = run.output_block_fused(model.out_blocks[1], b1, X["rbf_out"], None, None)
        a2, b2 = run.interaction_block(model.int_blocks[1], a1, b1, basis, 0)
        Eb, Fb = run.output_block_fused(model.out_blocks[2], b2, X["rbf_out"], Ea, Fa)
        for ours, ref, nm in ((a2, h2, "h"), (b2, m2, "m"), (Eb, Esum, "E")) + (((Fb, F0 + F1, "F"),) if direct else ()):
            err = (ours - ref).abs().max() / ref.abs().max()
            assert err < 1e-10, f"forward {nm}: {err:.2e}"
        tape.add_grad(a2, Gh)
        tape.add_grad(b2, Gm)
        tape.add_grad(Eb, GE)
        if direct:
            tape.add_grad(Fb, GF)
        tape.backward()
        for k in X:
            ours = tape._take(id(X[k]))
            ref = ref_grads[k]
            assert ours is not None, f"no gradient for {k}"
            err = (ours - ref).abs().max() / ref.abs().max()
            assert err < 1e-9, f"gradient {k}: {err:.2e}"


@pytest.mark.parametrize("name", ["dT", "dQ"])
def test_training_weight_gradients_match_oracle(name):
    """Tape(train=True): the fused blocks' backward also yields dL/dW for every weight they hold (operands saved by the
    chain programs -> library GEMMs), equal to torch autograd on the oracle's blocks (float64)."""
    from gemnet_pytorch_b200 import engine, fused
    kw, model, oracle, inp = setup(name, [1, 4, 7], seed=13)
    tri = kw["triplets_only"]
    plan = cpu_plan(inp, tri)
    E, A, T = plan.n_edges, plan.n_atoms, int(plan.id3_reduce.numel())
    g = torch.Generator().manual_seed(4)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)  # noqa: E731
    S = kw["num_spherical"]
    rbf = rnd(E, kw["num_radial"])
    leaves = {"h": rnd(A, 128), "m": rnd(E, 128), "sph3": rnd(T, S), "rbf_W1_3": rnd(E, 16, S) * 0.3}
    if not tri:
        M, Q = int(inp["id4_expand_intm_db"].numel()), int(inp["id4_expand_abd"].numel())
        leaves.update({"cbf4": rnd(M, 16), "ylm": rnd(Q, S * S) * 0.3, "rbf_W1_4": rnd(E, 32, S * S) * 0.1})
    GE, GF, Gm, Gh = rnd(A, 1), rnd(E, 1), rnd(E, 128), rnd(A, 128)

    # ---- oracle blocks, parameters as autograd leaves
    for v in oracle.p.values():
        v.requires_grad_(True)
    gd = {"id_swap": inp["id_swap"], "id3_expand_ba": inp["id3_expand_ba"], "id3_reduce_ca": inp["id3_reduce_ca"],
          "Kidx3": inp["Kidx3"], "id_c": inp["id_c"], "id_a": inp["id_a"],
          "rbf3": oracle.dense("mlp_rbf3", rbf, False), "rbf_h": oracle.dense("mlp_rbf_h", rbf, False),
          "cbf3": (leaves["rbf_W1_3"], go.pad_by_segment(leaves["sph3"], inp["id3_reduce_ca"], inp["Kidx3"], E).transpose(1, 2))}
    if not tri:
        gd.update(rbf4=oracle.dense("mlp_rbf4", rbf, False), cbf4=leaves["cbf4"], Kidx4=inp["Kidx4"],
                  id4_reduce_ca=inp["id4_reduce_ca"], id4_expand_intm_db=inp["id4_expand_intm_db"],
                  id4_expand_abd=inp["id4_expand_abd"],
                  sbf4=(leaves["rbf_W1_4"], go.pad_by_segment(leaves["ylm"], inp["id4_reduce_ca"], inp["Kidx4"], E).transpose(1, 2)))
    rbf_out = oracle.dense("mlp_rbf_out", rbf, False)
    h1, m1 = oracle.interaction_block("int_blocks.0", leaves["h"], leaves["m"], gd)
    E0, F0 = oracle.output_block("out_blocks.1", h1, m1, rbf_out, inp["id_a"])
    h2, m2 = oracle.interaction_block("int_blocks.1", h1, m1, gd)
    E1, F1 = oracle.output_block("out_blocks.2", h2, m2, rbf_out, inp["id_a"])
    loss = ((E0 + E1) * GE).sum() + ((F0 + F1) * GF).sum() + (m2 * Gm).sum() + (h2 * Gh).sum()
    loss.backward()

    # ---- fused engine in training mode
    with ce.patched(), torch.no_grad():
        tape = engine.Tape(enabled=True, train=True)
        for prm in model.parameters():
            tape.require(prm)
        run = engine._Run(model, tape, plan)
        assert run.fuse and run.fuse_out and fused.rbf_fan_ok(model)
        level1, uses = fused.fan_layout(model)
        outs = tape.apply(fused.RbfFan, rbf.clone(), level1, uses)
        run.fan_out = {id(d): o for (_, d), o in zip(uses, outs)}
        X = {k: v.clone() for k, v in leaves.items()}
        tape.require(X["h"])
        tape.require(X["m"])
        basis = dict(X, rbf3=None, rbf_h=None, rbf4=None)
        a1, b1 = run.interaction_block(model.int_blocks[0], X["h"], X["m"], basis, 0)
        Ea, Fa = run.output_block_fused(model.out_blocks[1], b1, None, None, None)
        a2, b2 = run.interaction_block(model.int_blocks[1], a1, b1, basis, 0)
        Eb, Fb = run.output_block_fused(model.out_blocks[2], b2, None, Ea, Fa)
        for ours, ref in ((Eb, E0 + E1), (Fb, F0 + F1)):
            assert (ours - ref).abs().max() / ref.abs().max() < 1e-10
        tape.add_grad(Eb, GE)
        tape.add_grad(Fb, GF)
        tape.add_grad(b2, Gm)
        tape.add_grad(a2, Gh)
        tape.backward()
        checked = 0
        for name_, prm in model.named_parameters():
            if name_.endswith(".weight") and ".linear." not in name_ and not name_.startswith(("mlp_", "atom_emb", "int_blocks", "out_blocks")):
                continue
            key = name_.replace(".linear.weight", ".weight")
            ref = oracle.p.get(key)
            if ref is None or ref.grad is None or "scale_factor" in name_ or ref.grad.abs().max() == 0:
                continue
            ours = tape._take(id(prm))
            if ours is None:
                continue
            err = (ours - ref.grad).abs().max() / ref.grad.abs().max()
            assert err < 1e-9, f"{name_}: {err:.2e}"
            checked += 1
        assert checked >= (45 if tri else 60), checked


@pytest.mark.parametrize("name", ["T", "Q"])
def test_rbf_fan_layout_and_gradient(name):
    """fused.RbfFan on the tape: every level-2 projection of the radial basis equals layer(level1(rbf)), and the
    summed gradient w.r.t. rbf equals torch autograd's (including consumers without a gradient)."""
    from gemnet_pytorch_b200 import engine, fused
    kw, model, oracle, inp = setup(name, [1, 4], seed=5)
    assert fused.rbf_fan_ok(model)
    E = int(inp["id_c"].numel())
    g = torch.Generator().manual_seed(1)
    rbf = torch.randn(E, kw["num_radial"], generator=g, dtype=torch.float64)
    level1, uses = fused.fan_layout(model)
    ref_in = rbf.clone().requires_grad_(True)
    ref = [torch.nn.functional.linear(torch.nn.functional.linear(ref_in, level1[gi].linear.weight), d.linear.weight)
           for gi, d in uses]
    Gs = [torch.randn(E, 128, generator=g, dtype=torch.float64) if j % 3 else None for j in range(len(uses))]
    loss = sum((r * G).sum() for r, G in zip(ref, Gs) if G is not None)
    g_ref = torch.autograd.grad(loss, ref_in)[0]
    with ce.patched(), torch.no_grad():
        tape = engine.Tape(enabled=True)
        x = rbf.clone()
        tape.require(x)
        outs = tape.apply(fused.RbfFan, x, level1, uses)
        for o, r in zip(outs, ref):
            assert (o - r).abs().max() < 1e-12
        for o, G in zip(outs, Gs):
            if G is not None:
                tape.add_grad(o, G, 0.5)
                tape.add_grad(o, G, 0.5)
        tape.backward()
        assert (tape._take(id(x)) - g_ref).abs().max() / g_ref.abs().max() < 1e-12


def test_fuse_ok_rejects_other_structures():
    from gemnet_pytorch_b200 import GemNet, engine
    kw = H.model_kwargs(H.load_npz("model_T.npz"))
    assert engine.fuse_ok(GemNet(**kw))
    assert not engine.fuse_ok(GemNet(**dict(kw, num_before_skip=0)))
    assert not engine.fuse_ok(GemNet(**dict(kw, num_concat=2)))
    # hyperparameters no kernel is built for are refused by the constructor, with the list of what is supported
    from gemnet_pytorch_b200 import _lib
    with pytest.raises(_lib.GemnetB200Error, match="built for"):
        GemNet(**dict(kw, emb_size_trip=20, emb_size_bil_trip=20))
    with pytest.raises(_lib.GemnetB200Error, match="num_spherical"):
        GemNet(**dict(kw, num_spherical=9))
