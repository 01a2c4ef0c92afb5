"""Fused dense layers on tcgen05 (include/gemnet_b200.h gemnet_dense_tc): forward/backward pairs in the protocol the
tape engine (engine.py) consumes — `forward(ctx, ...)` / `backward(ctx, *grads)` — for
  * Dense        y = f(x W^T)                      (base_layers.py:45-48)      backward  gx = (g .* dact) W
  * ResidualTC   y = (x + f(f(x W1^T) W2^T)) * s   (base_layers.py:61-89)      backward  gx = ((g .* d2) W2 .* d1) W1 + s g
Weights are packed once per (storage, version, orientation) into the kernel's resident shared-memory image.
"""
import os

import torch

from . import _lib
from .ops import _L, _check, _f32, _p, _stream

_packs = {}
PDL = [0 if os.environ.get("GEMNET_B200_PDL", "1") == "0" else 1]  # programmatic dependent launch of dense_tc


def supported(N, K):
    return 4 <= N <= 128 and 4 <= K <= 128 and N % 4 == 0 and K % 4 == 0


def packed_weight(W, transpose):
    """B operand of acc = A @ B^T.  transpose=False: B = W (N x K) -> y = x W^T;  True: B = W^T -> gx = g W.

    One cache entry per (storage, shape, strides, orientation); it is re-packed when the tensor's version counter moves
    (optimizer steps, copy_, load_state_dict).  In-place writes through `.data` do NOT move the counter
    (ExponentialMovingAverage.copy_to / restore of the reference, ema_decay.py:110,145): GemNet.train()/eval() mode
    changes, load_state_dict and .to() call `invalidate()`, which covers the reference Trainer's
    load_averaged_variables -> eval -> restore -> train cycle; anything else must call GemNet.invalidate_caches()."""
    key = (W.data_ptr(), tuple(W.shape), tuple(W.stride()), transpose)
    hit = _packs.get(key)
    if hit is not None and hit[6] == W._version:
        return hit
    N, K = (W.shape[1], W.shape[0]) if transpose else (W.shape[0], W.shape[1])
    sn, sk = (W.stride(1), W.stride(0)) if transpose else (W.stride(0), W.stride(1))
    if PACKER[0] is not None:  # CPU tests of the chain programs' algebra (tests/chain_emulator.py): no image
        return [None, N, K, W, False, transpose, W._version]
    n = _L().gemnet_dense_pack_floats(N, K)
    buf = hit[0] if hit is not None else torch.empty(n, dtype=torch.float32, device=W.device)
    _check(_L().gemnet_dense_pack_weight(_p(W), sn, sk, N, K, _p(buf), _stream()), "gemnet_dense_pack_weight")
    if hit is None and len(_packs) > 8192:
        _packs.clear()
    # keep W referenced: while the entry lives its storage cannot be freed and handed to another tensor with the
    # same data_ptr — which would silently serve a stale image
    # [4]: the pack kernel is the most recent launch -> the first consumer must not prefetch the image early
    _packs[key] = [buf, N, K, W, True, transpose, W._version]
    return _packs[key]


PACKER = [None]
_table = [None, None]   # (signature of the cached entries, device table of gemnet_pack_desc)


def refresh():
    """Re-pack EVERY cached image in one launch (gemnet_dense_pack_many) and mark them current.  Training calls this at
    the start of each step: the optimiser has changed every weight, and ~230 single-weight pack launches per step
    would cost more than the step's GEMMs.  Inside a captured CUDA graph the launch is part of the graph, so replays
    re-pack from the live weights."""
    if not _packs or PACKER[0] is not None:
        return
    import numpy as np
    entries = list(_packs.values())
    sig = tuple((e[0].data_ptr(), e[3].data_ptr(), e[5]) for e in entries)
    if _table[0] != sig:
        rows = np.zeros((len(entries), 6), dtype=np.int64)
        for i, e in enumerate(entries):
            buf, N, K, W, _, tr, _ = e
            sn, sk = (W.stride(1), W.stride(0)) if tr else (W.stride(0), W.stride(1))
            Np, Kp = (64 if N <= 64 else 128), (64 if K <= 64 else 128)
            rows[i] = (W.data_ptr(), buf.data_ptr(), sn, sk, N | (K << 32), Np | (Kp << 32))
        _table[0], _table[1] = sig, torch.from_numpy(rows).to(entries[0][0].device)
    _check(_L().gemnet_dense_pack_many(_table[1].data_ptr(), len(entries), _stream()), "gemnet_dense_pack_many")
    for e in entries:
        e[4], e[6] = True, e[3]._version


def invalidate():
    """Drop every packed weight image (they are rebuilt on next use)."""
    _packs.clear()
    from . import ops
    ops._segbil_packs.clear()


_fn = []


def dense_tc(x, pack, act=0, s_acc=1.0, res=None, s_res=0.0, xmul=None, want_dact=False):
    """y (M,N) [, dact] = gemnet_dense_tc(...).  x may be a row-strided view (inner stride 1).
    (Hot host path: ~230 calls per GemNet-T evaluation — kept free of helper calls.)"""
    buf, N, K = pack[0], pack[1], pack[2]
    M, ldx = x.shape[0], x.stride(0)
    if x.stride(1) != 1 or x.shape[1] != K or x.dtype is not torch.float32 or not x.is_cuda:
        raise _lib.GemnetB200Error("dense_tc: x must be a CUDA float32 (M, K) tensor with unit inner stride")
    if xmul is not None and (xmul.shape != x.shape or xmul.stride() != x.stride()):
        raise _lib.GemnetB200Error("dense_tc: xmul must match x in shape and strides")
    if want_dact:
        both = torch.empty((2, M, N), dtype=torch.float32, device=x.device)  # one allocation for y and dact
        y, dact = both[0], both[1]
        pd = dact.data_ptr()
    else:
        y, dact, pd = torch.empty((M, N), dtype=torch.float32, device=x.device), None, None
    if not _fn:
        _fn.append(_L().gemnet_dense_tc)
    flags = PDL[0]  # GEMNET_DENSE_WEIGHTS_SETTLED
    if pack[4]:
        pack[4], flags = False, 0
    rc = _fn[0](x.data_ptr(), None if xmul is None else xmul.data_ptr(), M, K, ldx, buf.data_ptr(), N, act, s_acc,
                None if res is None else res.data_ptr(), s_res, y.data_ptr(), pd, flags,
                torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    _lib.KERNEL_LAUNCHES[0] += 1
    if rc != 0:
        _check(rc, "gemnet_dense_tc")
    return (y, dact) if want_dact else y


class DenseTC:
    """forward(ctx, x, W, act) -> y ; backward(ctx, g) -> (gx, None, None)"""

    @staticmethod
    def forward(ctx, x, W, act):
        x = _f32(x) if x.stride(-1) != 1 else x
        need = ctx.needs_input_grad[0] or getattr(ctx, "train", False)
        ctx.W, ctx.act = W, act
        ctx.x = x if getattr(ctx, "train", False) else None
        if act and need:
            y, ctx.dact = dense_tc(x, packed_weight(W, False), act=1, want_dact=True)
        else:
            y, ctx.dact = dense_tc(x, packed_weight(W, False), act=act), None
        return y

    @staticmethod
    def backward(ctx, g):
        g = _f32(g)
        if getattr(ctx, "train", False):   # training: dW = (g .* dact)^T x (library GEMM over the rows)
            gz = g if ctx.dact is None else g * ctx.dact
            ctx.tape.add_wgrad(ctx.W, torch.mm(gz.t(), ctx.x))
            if not ctx.needs_input_grad[0]:
                return None, None, None
        return dense_tc(g, packed_weight(ctx.W, True), xmul=ctx.dact), None, None


class ResidualTC:
    """forward(ctx, x, W1, W2, s) -> (x + ssilu(ssilu(x W1^T) W2^T)) * s ; backward(ctx, g) -> (gx, None, None, None)"""

    @staticmethod
    def forward(ctx, x, W1, W2, s):
        if getattr(ctx, "train", False):
            raise _lib.GemnetB200Error("training needs the chain-fused residual stacks (engine.fuse_ok): unsupported block structure")
        x = _f32(x)
        need = ctx.needs_input_grad[0]
        ctx.W1, ctx.W2, ctx.s = W1, W2, s
        if need:
            y1, ctx.d1 = dense_tc(x, packed_weight(W1, False), act=1, want_dact=True)
            out, ctx.d2 = dense_tc(y1, packed_weight(W2, False), act=1, s_acc=s, res=x, s_res=s, want_dact=True)
        else:
            y1 = dense_tc(x, packed_weight(W1, False), act=1)
            out = dense_tc(y1, packed_weight(W2, False), act=1, s_acc=s, res=x, s_res=s)
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32(g)
        t = dense_tc(g, packed_weight(ctx.W2, True), xmul=ctx.d2)                 # (g .* s f'(z2)) W2
        gx = dense_tc(t, packed_weight(ctx.W1, True), xmul=ctx.d1, res=g, s_res=ctx.s)  # (t .* f'(z1)) W1 + s g
        return gx, None, None, None
