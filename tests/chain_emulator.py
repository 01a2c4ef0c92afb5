"""TEST INFRASTRUCTURE: a torch interpreter of gemnet_chain_tc programs (include/gemnet_b200.h) and torch stand-ins
for the few other kernels the fused engine calls, so that the ALGEBRA of the chain programs in
gemnet_pytorch_b200/fused.py (forward and hand-derived backward) can be checked against the CPU oracle without a GPU,
in float64.  Nothing in the product imports this; `patched()` swaps the launch functions for the duration of a test.

The same interpreter (on CUDA tensors, float64) is the yardstick for the kernel itself in tests/test_gpu_chain.py.
"""
import contextlib

import torch

from gemnet_pytorch_b200 import chain, dense_ops, ops

ACT, CARRY_ADD, SET_CARRY, TO_A, A_FROM_CARRY = chain.ACT, chain.CARRY_ADD, chain.SET_CARRY, chain.TO_A, chain.A_FROM_CARRY


def _ssilu(z):
    sg = torch.sigmoid(z)
    return z * sg / 0.6, sg * (1 + z * (1 - sg)) / 0.6


def weight_matrix(rec, dtype):
    """(N, K) operand of acc = A @ B^T from a packed-weight record [buf, N, K, W, flag, transpose, version]"""
    W = rec[3].detach()
    B = W.t() if rec[5] else W
    assert tuple(B.shape) == (rec[1], rec[2])
    return B.to(dtype)


def emulate(prog, settled=True, dtype=None):
    """Execute a chain.Program with whole-tensor torch ops; writes the program's output tensors in place."""
    M = prog.M
    carry = A = None
    for i, st in enumerate(prog.steps):
        K, N = st.K, st.N
        dt = dtype
        if st.terms:
            v = None
            for tm in st.terms:
                d = dt or tm.src.dtype
                if tm.kind == 0:
                    src = tm.src.reshape(-1, K).to(d)
                    src = src[tm.idx.long()] if tm.idx is not None else src[:M]
                    if tm.mul is not None:
                        src = src * tm.mul.reshape(M, K).to(d)
                    t = tm.scale * src
                else:
                    t = tm.scale * tm.src.reshape(-1)[:M].to(d)[:, None] * tm.mul.reshape(-1)[:K].to(d)[None, :]
                v = t if v is None else v + t
            if st.load_set_carry:
                carry = v.clone()
            A = v * st.amul.reshape(M, K).to(v.dtype) if st.amul is not None else v
        assert A is not None and A.shape == (M, K), f"step {i}: operand {None if A is None else A.shape} vs K={K}"
        if st.save_x is not None:
            st.save_x.copy_(A)
        z = A @ weight_matrix(st.weight, A.dtype).t()

        def rows(t, idx, width):
            t = t.reshape(-1, width).to(z.dtype)
            return t[idx.long()] if idx is not None else t[:M]

        if st.pre0 is not None:
            z = z + rows(st.pre0, st.pre0_idx, N)
        if st.pre1 is not None:
            z = z + rows(st.pre1, st.pre1_idx, N)
        if st.flags & ACT:
            v, d = _ssilu(z)
        else:
            v, d = z, torch.ones_like(z)
        if st.mul_early is not None:
            mu = rows(st.mul_early, None, N)
            if st.aux is not None:
                st.aux.copy_(v * st.s_acc)
            v, d = v * mu, d * mu
        v, d = v * st.s_acc, d * st.s_acc
        if st.flags & CARRY_ADD:
            assert carry is not None and carry.shape == v.shape, f"step {i}: carry shape"
            v = v + carry * st.s_carry
        if st.post0 is not None:
            v = v + rows(st.post0, None, N) * st.s_post0
        if st.post1 is not None:
            v = v + rows(st.post1, None, N) * st.s_post1
        if st.mul_late is not None:
            v = v * rows(st.mul_late, None, N)
        if st.dact is not None:
            st.dact.copy_(d)
        if st.out is not None:
            o = v * st.s_store
            if st.smul is not None:
                o = o * rows(st.smul, None, N)
            if st.out_idx is not None:
                st.out[st.out_idx.long()] = o.to(st.out.dtype)
            else:
                st.out.copy_(o)
        if st.dot_w is not None:
            dot = (v * st.dot_w.reshape(-1)[:N].to(v.dtype)[None, :]).sum(1, keepdim=True)
            if st.dot_in is not None:
                dot = dot + st.dot_in.reshape(M, 1).to(v.dtype)
            st.dot_out.copy_(dot.reshape(st.dot_out.shape))
        if st.flags & SET_CARRY:
            carry = v * st.s_set_carry
        last = i + 1 == len(prog.steps)
        if not last:
            Kn = prog.steps[i + 1].K
            if st.flags & TO_A:
                A = v * st.s_a
                assert A.shape[1] == Kn or prog.steps[i + 1].terms, f"step {i}: TO_A width {A.shape[1]} vs next K {Kn}"
                if st.next_amul is not None:
                    A = A * st.next_amul.reshape(M, Kn).to(A.dtype)
            elif st.flags & A_FROM_CARRY:
                A = carry * st.s_a
                if st.next_amul is not None:
                    A = A * st.next_amul.reshape(M, Kn).to(A.dtype)


# ------------------------------------------------------------------------------------------------ other stand-ins
def _dense_tc(x, pack, act=0, s_acc=1.0, res=None, s_res=0.0, xmul=None, want_dact=False):
    a = x * xmul if xmul is not None else x
    z = a @ weight_matrix(pack, x.dtype).t()
    if act:
        v, d = _ssilu(z)
    else:
        v, d = z, torch.ones_like(z)
    y = v * s_acc + (res * s_res if res is not None else 0)
    return (y, d * s_acc) if want_dact else y


def _segbil(x, sph, rbf_W1, W, expand, rowptr, scale):
    E = rbf_W1.shape[0]
    counts = (rowptr[1:] - rowptr[:-1]).long()
    red = torch.repeat_interleave(torch.arange(E), counts)
    A = torch.zeros(E, sph.shape[1], x.shape[1], dtype=x.dtype).index_add(0, red, sph[:, :, None] * x[expand.long()][:, None, :])
    B = torch.einsum("eis,esc->eci", rbf_W1, A)
    return torch.einsum("eci,cio->eo", B, W) * scale


class _SegBilinearCPU:
    @staticmethod
    def forward(ctx, x, sph, rbf_W1, W, expand, rowptr, scale, algo):
        ctx.save = (x, sph, rbf_W1, W, expand, rowptr, scale)
        return _segbil(x, sph, rbf_W1, W, expand, rowptr, scale)

    @staticmethod
    def backward(ctx, g):
        x, sph, rbf_W1, W, expand, rowptr, scale = ctx.save
        with torch.enable_grad():
            xs = [t.detach().requires_grad_(True) for t in (x, sph, rbf_W1, W)]
            out = _segbil(xs[0], xs[1], xs[2], xs[3], expand, rowptr, scale)
            gs = torch.autograd.grad(out, xs, g)
        return gs[0], gs[1], gs[2], (gs[3] if ctx.needs_input_grad[3] else None), None, None, None, None


def _e2a_forward(ctx, m, r, id_a, perm, rowptr, scale, had_scale):
    A = rowptr.numel() - 1
    x = m * r
    out = torch.zeros(A, m.shape[1], dtype=m.dtype).index_add(0, id_a.long(), x) * scale
    ctx.save_for_backward(m, r, id_a)
    ctx.cfg = (scale, had_scale)
    if had_scale is None:
        return out
    return out, x * had_scale


def _grad_r(m, id_a, g_out, s_out, g_had, s_had):
    g = 0
    if g_out is not None:
        g = g + s_out * g_out[id_a.long()]
    if g_had is not None:
        g = g + s_had * g_had
    return g * m


def _materialize(g, idx, mul, scale):
    t = g[idx.long()] if idx is not None else g
    if mul is not None:
        t = t * mul
    return t * scale


def _segment_sum_rows(src, perm, rowptr, n_rows):
    counts = (rowptr[1:] - rowptr[:-1]).long()
    seg = torch.repeat_interleave(torch.arange(n_rows), counts)
    return torch.zeros(n_rows, src.shape[1], dtype=src.dtype).index_add(0, seg, src[perm.long()])


def _lincomb(terms):
    tot = None
    for t, s in terms:
        tot = t * s if tot is None else tot + t * s
    return tot


class _GatherRowsCPU:
    @staticmethod
    def forward(ctx, x, idx, perm, rowptr):
        ctx.idx, ctx.n = idx.long(), x.shape[0]
        return x[ctx.idx]

    @staticmethod
    def backward(ctx, g):
        return torch.zeros(ctx.n, g.shape[1], dtype=g.dtype).index_add(0, ctx.idx, g), None, None, None


class _MulScaleCPU:
    @staticmethod
    def forward(ctx, a, b, s):
        ctx.save = (a, b, s)
        return a * b * s

    @staticmethod
    def backward(ctx, g):
        a, b, s = ctx.save
        return g * b * s, g * a * s, None


class _AddScaleCPU:
    @staticmethod
    def forward(ctx, s, a, b, c):
        out = a
        if b is not None:
            out = out + b
        if c is not None:
            out = out + c
        return out * s


def _rbf_fan_fwd(rbf, W1s, W2s, groups):
    return torch.stack([(rbf @ W1s[g].t()) @ W2.t() for g, W2 in zip(groups, W2s)])


def _rbf_fan_bwd(E, R, W1s, W2s, groups, g_outs):
    tot = torch.zeros(E, R, dtype=W2s[0].dtype)
    for g, W2, go in zip(groups, W2s, g_outs):
        if go is not None:
            tot = tot + (go @ W2) @ W1s[g]
    return tot


@contextlib.contextmanager
def patched(dtype=None):
    """Run the fused engine's host logic on CPU tensors: every kernel launch it makes goes to a torch stand-in."""
    saved = (chain.RUNNER[0], dense_ops.PACKER[0], dense_ops.dense_tc, ops._SegBilinear, ops._EdgeToAtom.forward,
             ops.edge_to_atom_grad_r, ops.materialize_term, ops.segment_sum_rows, ops.lincomb, ops._GatherRows,
             ops._MulScale, ops._f32)
    saved_fan = (ops.rbf_fan_fwd, ops.rbf_fan_bwd)
    ops.rbf_fan_fwd, ops.rbf_fan_bwd = _rbf_fan_fwd, _rbf_fan_bwd
    chain.RUNNER[0] = lambda prog, settled: emulate(prog, settled, dtype)
    dense_ops.PACKER[0] = "emulated"
    dense_ops.dense_tc = _dense_tc
    ops._SegBilinear = _SegBilinearCPU
    ops._EdgeToAtom.forward = staticmethod(_e2a_forward)
    ops.edge_to_atom_grad_r, ops.materialize_term = _grad_r, _materialize
    ops.segment_sum_rows, ops.lincomb = _segment_sum_rows, _lincomb
    ops._GatherRows, ops._MulScale = _GatherRowsCPU, _MulScaleCPU
    ops._f32 = lambda t: t
    dense_ops._f32 = lambda t: t
    dense_ops.invalidate()
    try:
        yield
    finally:
        (chain.RUNNER[0], dense_ops.PACKER[0], dense_ops.dense_tc, ops._SegBilinear, fwd, ops.edge_to_atom_grad_r,
         ops.materialize_term, ops.segment_sum_rows, ops.lincomb, ops._GatherRows, ops._MulScale, ops._f32) = saved
        ops._EdgeToAtom.forward = staticmethod(fwd) if not isinstance(fwd, staticmethod) else fwd
        ops.rbf_fan_fwd, ops.rbf_fan_bwd = saved_fan
        dense_ops._f32 = ops._f32
        dense_ops.invalidate()
