"""GPU parity of gemnet_chain_tc (csrc/chain_tc.cu) against the float64 interpreter of the same program
(tests/chain_emulator.py): every operand-fill / epilogue feature, the K / N paddings, ragged row counts, multi-tile CTAs.
Tolerance: 2e-6 of the output's max (fp32-grade: the kernel splits operands exactly and compensates the accumulator)."""
import numpy as np
import pytest
import torch

from tests import chain_emulator as ce

pytestmark = pytest.mark.gpu
TOL = 2e-6


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def run_both(build, dev):
    """build(out: dict-like allocator) -> Program; run the kernel — with automatic, 16-row and 64-row tiles — and the
    interpreter on separate outputs, compare."""
    from gemnet_pytorch_b200 import _lib
    try:
        for rows in (0, 16, 64):
            _lib.check(_lib.load().gemnet_chain_set_tile_rows(rows), "gemnet_chain_set_tile_rows")
            out = _run_both(build, dev)
    finally:
        _lib.load().gemnet_chain_set_tile_rows(0)
    return out


def _run_both(build, dev):
    from gemnet_pytorch_b200 import chain
    res = []
    for which in (0, 1):
        outs = {}

        def alloc(name, *shape, outs=outs):
            outs[name] = torch.full(shape, float("nan"), dtype=torch.float32, device=dev)
            return outs[name]

        prog = build(alloc)
        if which == 0:
            chain.run(prog)
            torch.cuda.synchronize()
        else:
            ce.emulate(prog, dtype=torch.float64)
        res.append(outs)
    for k in res[0]:
        a, b = res[0][k].double(), res[1][k].double()
        assert torch.isfinite(a).all(), f"{k}: non-finite kernel output"
        err = (a - b).abs().max() / b.abs().max().clamp(min=1e-30)
        assert err < TOL, f"{k}: rel err {err:.3e}"
    return res[0]


def W(n, k, dev, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, k, generator=g) / k ** 0.5).to(dev)


def X(m, k, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(m, k, generator=g) * scale).to(dev)


@pytest.mark.parametrize("M,K,N", [(300, 128, 128), (1, 128, 128), (129, 64, 128), (1000, 128, 64), (257, 16, 32),
                                   (77, 32, 128), (148 * 128 * 2 + 77, 128, 128)])
def test_single_layer(M, K, N, dev):
    from gemnet_pytorch_b200 import chain, dense_ops
    from gemnet_pytorch_b200.chain import ACT, Program, Step, Term
    x, w = X(M, K, dev, 1), W(N, K, dev, 2)

    def build(alloc):
        return Program(M, [Step(dense_ops.packed_weight(w, False), terms=[Term(x)], flags=ACT, s_acc=0.7,
                                dact=alloc("d", M, N), out=alloc("y", M, N))])

    run_both(build, dev)


def test_residual_chain_with_terms(dev):
    """operand fill with three terms (gather + Hadamard), carry from the load with amul, a residual layer, post terms"""
    from gemnet_pytorch_b200 import dense_ops
    from gemnet_pytorch_b200.chain import ACT, CARRY_ADD, SET_CARRY, TO_A, Program, Step, Term
    M, H = 1500, 128
    g = torch.Generator().manual_seed(5)
    src_rows = 400
    t1, t2, t3 = X(M, H, dev, 3), X(src_rows, H, dev, 4), X(M, H, dev, 6)
    idx = torch.randint(0, src_rows, (M,), generator=g, dtype=torch.int32).to(dev)
    mul, amul, post0, post1 = X(M, H, dev, 7), X(M, H, dev, 8), X(M, H, dev, 9), X(M, H, dev, 10)
    w1, w2, w3 = W(H, H, dev, 11), W(H, H, dev, 12), W(H, H, dev, 13)
    d1 = X(M, H, dev, 14)

    def build(alloc):
        pw = dense_ops.packed_weight
        return Program(M, [
            Step(pw(w1, True), terms=[Term(t1, 0.5), Term(t2, -1.25, idx=idx, mul=mul), Term(t3, 2.0)],
                 load_set_carry=True, amul=amul, flags=TO_A, next_amul=d1),
            Step(pw(w2, False), flags=ACT | CARRY_ADD | SET_CARRY | TO_A, s_acc=0.7, s_carry=0.3, post0=post0, s_post0=0.1,
                 post1=post1, s_post1=-0.2, dact=alloc("d2", M, H), out=alloc("y2", M, H), s_store=1.5, s_set_carry=0.9,
                 s_a=1.1),
            Step(pw(w3, False), flags=CARRY_ADD, s_carry=2.0, out=alloc("y3", M, H))])

    run_both(build, dev)


def test_prologue_shapes_and_aux(dev):
    """the TripPrologue program: mul_early + aux + dact, 128 -> 128 -> 64, and its backward with smul / next_amul"""
    from gemnet_pytorch_b200 import dense_ops
    from gemnet_pytorch_b200.chain import ACT, TO_A, Program, Step, Term
    M = 8186
    m, r = X(M, 128, dev, 1), X(M, 128, dev, 2)
    wb, wd = W(128, 128, dev, 3), W(64, 128, dev, 4)
    pw = dense_ops.packed_weight

    def fwd(alloc):
        return Program(M, [Step(pw(wb, False), terms=[Term(m)], flags=ACT | TO_A, s_acc=1.3, mul_early=r,
                                aux=alloc("aux", M, 128), dact=alloc("d1", M, 128)),
                           Step(pw(wd, False), flags=ACT, dact=alloc("d2", M, 64), out=alloc("x", M, 64))])

    o = run_both(fwd, dev)
    gx = X(M, 64, dev, 9)

    def bwd(alloc):
        return Program(M, [Step(pw(wd, True), terms=[Term(gx)], amul=o["d2"], flags=TO_A, out=alloc("g_r", M, 128),
                                smul=o["aux"], next_amul=o["d1"]),
                           Step(pw(wb, True), out=alloc("g_m", M, 128))])

    run_both(bwd, dev)


def test_from_carry_scatter_dot_outer(dev):
    """A_FROM_CARRY, scatter store through out_idx, mul_late, the dot epilogue (with dot_in) and the outer-product fill"""
    from gemnet_pytorch_b200 import dense_ops
    from gemnet_pytorch_b200.chain import A_FROM_CARRY, ACT, CARRY_ADD, SET_CARRY, TO_A, Program, Step, Term
    M, H, C = 700, 128, 64
    g = torch.Generator().manual_seed(1)
    gE = X(M, 1, dev, 1)
    wout = X(1, H, dev, 2)
    w1, w2, w3 = W(H, H, dev, 3), W(H, C, dev, 4), W(H, C, dev, 5)   # w2/w3 used transposed: N = C, K = H
    perm = torch.randperm(M, generator=g).to(torch.int32).to(dev)
    da, db, late = X(M, H, dev, 6), X(M, H, dev, 7), X(M, H, dev, 8)
    dot_in = X(M, 1, dev, 9)
    pw = dense_ops.packed_weight

    def build(alloc):
        return Program(M, [
            Step(pw(w1, False), terms=[Term(gE, 0.8, mul=wout, kind=1)], load_set_carry=True, flags=ACT | CARRY_ADD | SET_CARRY
                 | A_FROM_CARRY, s_carry=0.5, mul_late=late, s_a=0.6, next_amul=da, out=alloc("y1", M, H),
                 dot_w=wout, dot_out=alloc("dot", M, 1), dot_in=dot_in),
            Step(pw(w2, True), out=alloc("ga", M, C), flags=A_FROM_CARRY, s_a=0.6, next_amul=db),
            Step(pw(w3, True), out=alloc("gb", M, C), out_idx=perm)])

    run_both(build, dev)


def test_edge_embedding_epilogue(dev):
    """pre terms gathered by two index lists (EdgeEmb), unchanged-operand steps (AtomChain's p_c / p_a)"""
    from gemnet_pytorch_b200 import dense_ops
    from gemnet_pytorch_b200.chain import ACT, SET_CARRY, TO_A, Program, Step, Term
    M, A, H = 5000, 600, 128
    g = torch.Generator().manual_seed(2)
    m, pc, pa = X(M, H, dev, 1), X(A, H, dev, 2), X(A, H, dev, 3)
    ic = torch.randint(0, A, (M,), generator=g, dtype=torch.int32).to(dev)
    ia = torch.randint(0, A, (M,), generator=g, dtype=torch.int32).to(dev)
    w0, w1, w2 = W(H, H, dev, 4), W(H, H, dev, 5), W(H, H, dev, 6)
    pw = dense_ops.packed_weight

    def build(alloc):
        return Program(M, [
            Step(pw(w0, False), terms=[Term(m)], pre0=pc, pre0_idx=ic, pre1=pa, pre1_idx=ia, flags=ACT | SET_CARRY | TO_A,
                 dact=alloc("d0", M, H), out=alloc("y0", M, H)),
            Step(pw(w1, False), out=alloc("p1", M, H)),
            Step(pw(w2, False), out=alloc("p2", M, H))])

    run_both(build, dev)


def test_rejects_bad_programs(dev):
    from gemnet_pytorch_b200 import _lib, chain, dense_ops
    from gemnet_pytorch_b200.chain import Program, Step, Term
    x, w = X(10, 128, dev, 1), W(128, 128, dev, 2)
    with pytest.raises(_lib.GemnetB200Error):
        chain.run(Program(10, [Step(dense_ops.packed_weight(w, False), out=torch.empty(10, 128, device=dev))]))  # no operand
    with pytest.raises(_lib.GemnetB200Error):
        chain.run(Program(10, [Step(dense_ops.packed_weight(w, False), terms=[Term(x.cpu())])]))  # CPU tensor
