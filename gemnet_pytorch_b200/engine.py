"""Inference engine: energy + forces of a GemNet batch without torch.autograd.

The shipped checkpoints compute forces as -dE/dR (gemnet/model/gemnet.py:598-613), i.e. every evaluation is a forward
AND a backward pass.  At batch 64 the step is launch/host bound: ~650 kernels, and torch.autograd spends
more host time per custom Function (graph node creation, Python trampolines, dispatcher) than the kernels take on the
GPU.  This module runs the SAME kernels (the forward/backward bodies of the autograd Functions in ops.py are reused
verbatim through a stand-in ctx object) on a minimal reverse-mode tape: a list of closures, gradients keyed by tensor
identity, and a "depends on R" flag that prunes everything that does not (weights, atom embeddings).  Weight
gradients are never formed.  Used by GemNet.forward in eval mode; training keeps using torch.autograd.
"""
import os

import torch

from . import chain, dense_ops, fused, graph, layers, ops


class _Ctx:
    """Stand-in for torch.autograd.function.FunctionCtx (what the Function bodies in ops.py touch)."""

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors


class _Linear:
    """y = x @ W^T (base_layers.py:45-48, bias-free); backward w.r.t. x only."""

    @staticmethod
    def forward(ctx, x, W):
        ctx.W = W
        ctx.x = x if getattr(ctx, "train", False) else None
        return torch.mm(x, W.t())

    @staticmethod
    def backward(ctx, g):
        if getattr(ctx, "train", False):
            ctx.tape.add_wgrad(ctx.W, torch.mm(g.t(), ctx.x))
            if not ctx.needs_input_grad[0]:
                return None, None
        return torch.mm(g, ctx.W), None


class _Embedding:
    """h = weight[Z - 1] (embedding_block.py:29-34); training: the embedding table's gradient"""

    @staticmethod
    def forward(ctx, weight, Z):
        ctx.idx, ctx.n = Z - 1, weight.shape[0]
        return weight[ctx.idx]

    @staticmethod
    def backward(ctx, g):
        return torch.zeros(ctx.n, g.shape[1], dtype=g.dtype, device=g.device).index_add_(0, ctx.idx, g), None


class Tape:
    """Reverse-mode tape.  With `streams=(main, side)` every recorded op belongs to a lane (0 = main, 1 = side);
    its backward runs on the same lane, and a gradient that crosses lanes makes the consumer's stream wait for the
    producer's (an event over everything the producer lane has enqueued so far).  Cross-lane tensors are parked
    in `keep` until the evaluation ends, so the caching allocator cannot hand their memory to the other stream."""

    def __init__(self, enabled=True, streams=None, train=False):
        self.enabled = enabled
        self.train = train  # training: gradients w.r.t. the PARAMETERS are wanted too (weights are `require`d)
        self.closures = []
        self.grads = {}
        self.req = set()
        self.streams = streams
        self.lane = 0
        self.keep = []
        self.used = set()  # side lanes that have work (only those are joined: an idle stream is not part of a capture)

    # -- lanes
    def sync_from(self, src):
        """the current lane's stream waits for everything enqueued so far on lane `src`"""
        self.streams[self.lane].wait_stream(self.streams[src])

    def switch(self, lane):
        if self.streams is not None and lane != self.lane:
            self.lane = lane
            torch.cuda.set_stream(self.streams[lane])

    def on_side(self, fn, lane=1):
        """run fn() on side lane `lane`, after what the main lane has enqueued so far"""
        if self.streams is None or lane >= len(self.streams):
            return fn()
        self.used.add(lane)
        self.switch(lane)
        self.sync_from(0)
        try:
            return fn()
        finally:
            self.switch(0)

    def join(self, lane=None):
        """main lane waits for side lane `lane` (default: all of them)"""
        if self.streams is not None:
            self.switch(0)
            for src in (sorted(self.used) if lane is None else (lane,)):
                if src in self.used:
                    self.sync_from(src)

    def _take_terms(self, key, lazy_ok=False):
        """pending gradient summands [(tensor, scale, idx, mul)] of a tensor; makes this lane wait for the lanes that
        wrote them.  A summand with idx / mul is LAZY: it stands for scale * tensor[idx] * mul (row gather, Hadamard
        product) and is evaluated inside the consumer's operand fill (fused.py); consumers that cannot do that get it
        materialised here (lazy_ok=False)."""
        terms = self.grads.pop(key, None)
        if terms is None:
            return None
        out = []
        for g, sc, src, idx, mul in terms:
            if self.streams is not None and src != self.lane:
                self.sync_from(src)
                self.keep.append(g)
            if not lazy_ok and (idx is not None or mul is not None):
                g, sc, idx, mul = ops.materialize_term(g, idx, mul, sc), 1.0, None, None
            out.append((g, sc, idx, mul))
        return out

    def reduce_terms(self, terms, limit, plain=False):
        """at most `limit` summands (the plain ones are pre-summed by lincomb3 when there are too many); plain=True
        also materialises lazy summands"""
        if plain:
            terms = [(ops.materialize_term(g, idx, mul, sc), 1.0, None, None) if (idx is not None or mul is not None)
                     else (g, sc, idx, mul) for g, sc, idx, mul in terms]
        if len(terms) <= limit:
            return terms
        lazy = [t for t in terms if t[2] is not None or t[3] is not None]
        flat = [t for t in terms if t[2] is None and t[3] is None]
        while len(lazy) >= limit:  # cannot keep them all lazy
            g, sc, idx, mul = lazy.pop()
            flat.append((ops.materialize_term(g, idx, mul, sc), 1.0, None, None))
        keep = limit - len(lazy) - 1
        head, tail = flat[:keep], flat[keep:]
        summed = ops.lincomb([(g, sc) for g, sc, _, _ in tail]) if len(tail) > 1 or tail[0][1] != 1.0 else tail[0][0]
        return head + [(summed, 1.0, None, None)] + lazy

    def _take(self, key):
        """the gradient of a tensor, materialised: sum_i scale_i * term_i in ceil((n-1)/2) launches"""
        terms = self._take_terms(key)
        if terms is None:
            return None
        if len(terms) == 1 and terms[0][1] == 1.0:
            return terms[0][0]
        g0 = terms[0][0]
        if all(g.shape == g0.shape and g.dtype is torch.float32 and g.is_cuda and g.is_contiguous() for g, _, _, _ in terms):
            return ops.lincomb([(g, sc) for g, sc, _, _ in terms])
        tot = None
        for g, sc, _, _ in terms:
            tot = g * sc if tot is None else tot + g * sc
        return tot

    def require(self, t):
        self.req.add(id(t))

    def needs(self, t):
        if not isinstance(t, torch.Tensor):
            return False
        return id(t) in self.req or (self.train and t._base is not None and id(t._base) in self.req)

    def on_wgrad_stream(self, fn, tensors=()):
        """run fn() (weight-gradient GEMMs) on the side stream `wstream` after what the current stream has enqueued;
        backward_train joins it before the gradients are collected.  `tensors` are read there: kept alive until the
        join so that the caching allocator cannot recycle them under the side stream."""
        ws = getattr(self, "wstream", None)
        if ws is None:
            return fn()
        self.keep.extend(tensors)
        cur = torch.cuda.current_stream()
        ws.wait_stream(cur)
        with torch.cuda.stream(ws):
            fn()

    def add_wgrad(self, W, dW):
        """gradient of a weight operand; W may be a column / row slice (view) of a parameter (EdgeEmbedding's three
        blocks): the slice's gradient is embedded in a zero tensor of the parameter's shape"""
        base = W._base if W._base is not None else W
        if base is W:
            self.add_grad(W, dW)
            return
        full = torch.zeros_like(base)
        full.as_strided(W.shape, W.stride(), W.storage_offset() - base.storage_offset()).copy_(dW)
        self.add_grad(base, full)

    def add_grad(self, t, g, scale=1.0, idx=None, mul=None):
        """Gradients are kept as lists of (tensor, scale, lane, idx, mul) summands and summed only when they are
        consumed: a tensor with n consumers costs ceil((n-1)/2) launches instead of n-1 (none at all when the consumer
        is a fused chain, whose operand fill sums them), and the 1/sqrt(2) factors of the skip connections ride along
        as scales instead of costing a launch each."""
        self.grads.setdefault(id(t), []).append((g, scale, self.lane, idx, mul))

    def apply(self, Fn, *args):
        """Run Fn.forward; if any tensor argument depends on R, record Fn.backward."""
        ctx = _Ctx()
        ctx.needs_input_grad = tuple(self.enabled and self.needs(a) for a in args)
        ctx.tape, ctx.train = self, self.train
        fused_fn = getattr(Fn, "fused", False)
        out = Fn.forward(ctx, *args)
        if not any(ctx.needs_input_grad) and not (self.train and fused_fn):  # (a fused block always holds weights)
            return out
        outs = out if isinstance(out, tuple) else (out,)
        for o in outs:
            self.req.add(id(o))

        lane = self.lane

        def backward_add_scale():
            # y = (a [+ b [+ c]]) * s: the gradient terms of y pass to every input with their scale multiplied by s
            self.switch(lane)
            terms = self._take_terms(id(out), lazy_ok=True)
            if terms is None:
                return
            for a, need in zip(args[1:], ctx.needs_input_grad[1:]):
                if need:
                    for g, sc, idx, mul in terms:
                        self.add_grad(a, g, sc * args[0], idx, mul)

        def backward_fused():
            # fused chains (fused.py) take the pending summands as they are and may return lazy summands
            self.switch(lane)
            terms = [self._take_terms(id(o), lazy_ok=True) for o in outs]
            if all(t is None for t in terms):
                return
            gin = Fn.backward_terms(ctx, self, terms)
            for a, need, g in zip(args, ctx.needs_input_grad, gin):
                if need and g is not None:
                    if isinstance(g, list):
                        for gt, sc, idx, mul in g:
                            self.add_grad(a, gt, sc, idx, mul)
                    else:
                        self.add_grad(a, g)

        def backward():
            self.switch(lane)
            gs = [self._take(id(o)) for o in outs]
            if all(g is None for g in gs):
                return
            gs = [torch.zeros_like(o) if g is None else g for g, o in zip(gs, outs)]
            gin = Fn.backward(ctx, *gs)
            if not isinstance(gin, tuple):
                gin = (gin,)
            for a, need, g in zip(args, ctx.needs_input_grad, gin):
                if need and g is not None:
                    self.add_grad(a, g)

        # keep tensors alive: gradients are keyed by id()
        fn = backward_add_scale if Fn is ops._AddScale else (backward_fused if getattr(Fn, "fused", False) else backward)
        self.closures.append((fn, outs, args))
        return out

    def backward(self):
        for fn, _, _ in reversed(self.closures):
            fn()
        self.join()
        self.closures.clear()


# ---------------------------------------------------------------------------------------------- layer evaluation
class _Run:
    """One evaluation: thin wrappers that mirror the module tree of layers.py on the tape."""

    def __init__(self, model, tape, plan):
        self.m, self.t, self.p = model, tape, plan
        self.tc = getattr(model, "dense_algo", 1) == 1  # fused tcgen05 dense layers where the shape is supported
        # chain-fused blocks (fused.py): one launch per dense stack instead of one per layer
        self.fuse = self.tc and getattr(model, "chain_fusion", True) and fuse_ok(model)
        self.fuse_out = self.fuse and model.num_targets == 1
        self.fan_out = {}   # id(level-2 rbf layer) -> its precomputed output (fused.RbfFan)

    # -- primitives
    def _tc_ok(self, W, x):
        return self.tc and dense_ops.supported(W.shape[0], W.shape[1]) and x.shape[0] > 0

    def linear(self, dense, x, W=None):
        W = dense.linear.weight if W is None else W
        if self._tc_ok(W, x):
            return self.t.apply(dense_ops.DenseTC, x, W, 0)
        return self.t.apply(_Linear, x, W)

    def act(self, dense, z):
        return self.t.apply(ops._ScaledSiLU, z) if isinstance(dense._activation, layers.ScaledSiLU) else z

    def dense(self, dense, x):
        W = dense.linear.weight
        if isinstance(dense._activation, layers.ScaledSiLU) and self._tc_ok(W, x):
            return self.t.apply(dense_ops.DenseTC, x, W, 1)
        return self.act(dense, self.linear(dense, x))

    def residual(self, layer, x):
        if (self.tc and len(layer.dense_mlp) == 2 and x.shape[0] > 0
                and all(isinstance(d._activation, layers.ScaledSiLU) and
                        dense_ops.supported(*d.linear.weight.shape) for d in layer.dense_mlp)):
            return self.t.apply(dense_ops.ResidualTC, x, layer.dense_mlp[0].linear.weight,
                                layer.dense_mlp[1].linear.weight, layers.INV_SQRT_2)
        y = x
        for d in layer.dense_mlp[:-1]:
            y = self.dense(d, y)
        last = layer.dense_mlp[-1]
        z = self.linear(last, y)
        if isinstance(last._activation, layers.ScaledSiLU):
            return self.t.apply(ops._ResidualTail, z, x, layers.INV_SQRT_2)
        return self.add_scale(layers.INV_SQRT_2, x, z)

    def add_scale(self, s, a, b=None, c=None):
        return self.t.apply(ops._AddScale, float(s), a, b, c)

    def mul_scale(self, a, b, s):
        return self.t.apply(ops._MulScale, a, b, float(s))

    def gather(self, x, idx, perm, rowptr):
        return self.t.apply(ops._GatherRows, x, idx, perm, rowptr)

    # -- blocks
    def edge_embedding(self, emb, h, m_rbf):
        p = self.p
        W = emb.dense.linear.weight
        A = h.shape[1]
        p_c = self.linear(None, h, W[:, :A])
        p_a = self.linear(None, h, W[:, A:2 * A])
        y = self.linear(None, m_rbf, W[:, 2 * A:])
        z = self.add_scale(1.0, y, self.gather(p_c, p.id_c, p.src_perm, p.src_rowptr),
                           self.gather(p_a, p.id_a, p.atom_perm, p.atom_rowptr))
        return self.act(emb.dense, z)

    def atom_update(self, blk, m, rbf, mlp):
        p = self.p
        x = self.t.apply(ops._EdgeToAtom, m, self.linear(blk.dense_rbf, rbf), p.id_a, p.atom_perm, p.atom_rowptr,
                         blk.scale_sum.value(), None)
        x = self.dense(mlp[0], x)
        for layer in list(mlp)[1:]:
            x = self.residual(layer, x)
        return x

    def output_block(self, blk, m, rbf_out):
        p = self.p
        r = self.linear(blk.dense_rbf, rbf_out)
        if blk.direct_forces:
            x_E, x_F = self.t.apply(ops._EdgeToAtom, m, r, p.id_a, p.atom_perm, p.atom_rowptr, blk.scale_sum.value(),
                                    blk.scale_rbf.value())
        else:
            x_E = self.t.apply(ops._EdgeToAtom, m, r, p.id_a, p.atom_perm, p.atom_rowptr, blk.scale_sum.value(), None)
            x_F = None
        x_E = self.dense(blk.seq_energy[0], x_E)
        for layer in list(blk.seq_energy)[1:]:
            x_E = self.residual(layer, x_E)
        x_E = self.linear(blk.out_energy, x_E)
        if blk.direct_forces:
            x_F = self.dense(blk.seq_forces[0], x_F)
            for layer in list(blk.seq_forces)[1:]:
                x_F = self.residual(layer, x_F)
            x_F = self.linear(blk.out_forces, x_F)
        return x_E, x_F

    def bilinear(self, mod, rbf_W1, sph, x, expand, rowptr, scale, algo):
        S = rbf_W1.shape[2]
        if algo == 1 and (S, mod.emb_size, mod.emb_size_interm, mod.units_out) != (7, 64, 16, 64):
            algo = 0
        W = mod.weight if self.t.train else mod.weight.detach()
        return self.t.apply(ops._SegBilinear, x, sph, rbf_W1, W, expand, rowptr, float(scale), int(algo))

    def triplet_interaction(self, ti, m, basis, algo):
        p = self.p
        x = self.dense(ti.dense_ba, m)
        x = self.mul_scale(x, self.linear(ti.mlp_rbf, basis["rbf3"]), ti.scale_rbf.value())
        x = self.dense(ti.down_projection, x)
        x = self.bilinear(ti.mlp_cbf, basis["rbf_W1_3"], basis["sph3"], x, p.id3_expand, p.rowptr3,
                          ti.scale_cbf_sum.value(), algo)
        self.t.keep.append(x)
        x_ac = self.t.on_side(lambda: self.gather(self.dense(ti.up_projection_ac, x), p.id_swap, p.swap_perm,
                                                  p.swap_rowptr), lane=3)
        x_ca = self.dense(ti.up_projection_ca, x)
        self.t.join(3)
        return self.add_scale(layers.INV_SQRT_2, x_ca, x_ac)

    def quadruplet_interaction(self, qi, m, basis, algo):
        p = self.p
        x = self.dense(qi.dense_db, m)
        x = self.mul_scale(x, self.linear(qi.mlp_rbf, basis["rbf4"]), qi.scale_rbf.value())
        x = self.dense(qi.down_projection, x)
        x = self.gather(x, p.id4_expand_intm_db, p.intm_db_perm, p.intm_db_rowptr)
        x = self.mul_scale(x, self.linear(qi.mlp_cbf, basis["cbf4"]), qi.scale_cbf.value())
        x = self.bilinear(qi.mlp_sbf, basis["rbf_W1_4"], basis["ylm"], x, p.id4_expand_abd, p.rowptr4,
                          qi.scale_sbf_sum.value(), algo)
        x_ca = self.dense(qi.up_projection_ca, x)
        x_ac = self.gather(self.dense(qi.up_projection_ac, x), p.id_swap, p.swap_perm, p.swap_rowptr)
        return self.add_scale(layers.INV_SQRT_2, x_ca, x_ac)

    # -- chain-fused blocks (fused.py)
    def rlin(self, dense, x):
        """a projection of the radial basis: precomputed by the fan kernel, or a dense layer of its own"""
        hit = self.fan_out.get(id(dense))
        return hit if hit is not None else self.linear(dense, x)

    def quad_path_fused(self, qi, m, basis, algo):
        p = self.p
        r4 = self.rlin(qi.mlp_rbf, basis["rbf4"])
        x = self.t.apply(fused.TripPrologue, m, r4, qi.dense_db, qi.down_projection, qi.scale_rbf.value())
        x = self.gather(x, p.id4_expand_intm_db, p.intm_db_perm, p.intm_db_rowptr)
        x = self.mul_scale(x, self.linear(qi.mlp_cbf, basis["cbf4"]), qi.scale_cbf.value())
        return self.bilinear(qi.mlp_sbf, basis["rbf_W1_4"], basis["ylm"], x, p.id4_expand_abd, p.rowptr4,
                             qi.scale_sbf_sum.value(), algo)

    def interaction_block_fused(self, blk, h, m, basis, algo):
        """InteractionBlock.forward (interaction_block.py:186-233 / 385-421) as 5 chain launches + the bilinear layers
        + the edge->atom scatter."""
        t, p, A = self.t, self.p, self.t.apply
        ti = blk.trip_interaction
        t.keep.append(m)
        x4 = None
        if not blk.triplets_only:
            # the quadruplet branch fills the GPU; the triplet branch (small, latency-bound launches) hides next to it
            x4 = t.on_side(lambda: self.quad_path_fused(blk.quad_interaction, m, basis, algo), lane=2)
        r3 = self.rlin(ti.mlp_rbf, basis["rbf3"])
        x = A(fused.TripPrologue, m, r3, ti.dense_ba, ti.down_projection, ti.scale_rbf.value())
        x3 = self.bilinear(ti.mlp_cbf, basis["rbf_W1_3"], basis["sph3"], x, p.id3_expand, p.rowptr3,
                           ti.scale_cbf_sum.value(), algo)
        if x4 is not None:
            t.join(2)
            t.keep.append(x4)
        m2 = A(fused.Post, m, x3, x4, blk, p)
        au = blk.atom_update
        rh = self.rlin(au.dense_rbf, basis["rbf_h"])
        x0 = A(fused.EdgeToAtomLazy, m2, rh, p.id_a, p.atom_perm, p.atom_rowptr, au.scale_sum.value(), None)
        h2, p_c, p_a = A(fused.AtomChain, x0, h, blk)
        m3 = A(fused.EdgeEmb, m2, p_c, p_a, blk, p)
        return h2, m3

    def output_block_fused(self, blk, m, rbf_out, E_prev, F_prev):
        """OutputBlock.forward (atom_update_block.py:157-193) accumulated onto the previous blocks' sums."""
        p, A = self.p, self.t.apply
        ro = self.rlin(blk.dense_rbf, rbf_out)
        if blk.direct_forces:
            x_E, x_F = A(fused.EdgeToAtomLazy, m, ro, p.id_a, p.atom_perm, p.atom_rowptr, blk.scale_sum.value(),
                         blk.scale_rbf.value())
        else:
            x_E = A(fused.EdgeToAtomLazy, m, ro, p.id_a, p.atom_perm, p.atom_rowptr, blk.scale_sum.value(), None)
        E = A(fused.OutChain, x_E, E_prev, blk.seq_energy, blk.out_energy)
        F = A(fused.OutChain, x_F, F_prev, blk.seq_forces, blk.out_forces) if blk.direct_forces else F_prev
        return E, F

    def interaction_block(self, blk, h, m, basis, algo):
        if self.fuse:
            return self.interaction_block_fused(blk, h, m, basis, algo)
        self.t.keep.append(m)
        x_skip = self.t.on_side(lambda: self.dense(blk.dense_ca, m), lane=3)
        if blk.triplets_only:
            x3 = self.triplet_interaction(blk.trip_interaction, m, basis, algo)  # joins lane 3
            x = self.add_scale(layers.INV_SQRT_2, x_skip, x3)
        else:
            # the quadruplet branch fills the GPU; the triplet branch (small, latency-bound launches) hides next to it
            self.t.keep.append(m)
            x4 = self.t.on_side(lambda: self.quadruplet_interaction(blk.quad_interaction, m, basis, algo), lane=2)
            x3 = self.triplet_interaction(blk.trip_interaction, m, basis, algo)  # joins lane 3
            self.t.join(2)
            x = self.add_scale(layers.INV_SQRT_3, x_skip, x3, x4)
        for layer in blk.layers_before_skip:
            x = self.residual(layer, x)
        m = self.add_scale(layers.INV_SQRT_2, m, x)
        for layer in blk.layers_after_skip:
            m = self.residual(layer, m)
        h2 = self.atom_update(blk.atom_update, m, basis["rbf_h"], blk.atom_update.layers)
        h = self.add_scale(layers.INV_SQRT_2, h, h2)
        m2 = self.edge_embedding(blk.concat_layer, h, m)
        for layer in blk.residual_m:
            m2 = self.residual(layer, m2)
        m = self.add_scale(layers.INV_SQRT_2, m, m2)
        return h, m


def fuse_ok(model):
    """Can the dense stacks of this model run as chain programs (fused.py)?  Needs the published block structure
    (>= 1 residual layer before / after the skip and in the atom MLPs, one concat residual, bias-free ScaledSiLU dense
    layers with sizes that are multiples of 8 up to 128); anything else keeps the layer-by-layer kernels."""
    cached = getattr(model, "_fuse_ok", None)
    if cached is not None:
        return cached
    ok = True
    try:
        for blk in model.int_blocks:
            paths = [blk.trip_interaction] + ([] if blk.triplets_only else [blk.quad_interaction])
            nb, na = len(blk.layers_before_skip), len(blk.layers_after_skip)
            ok &= nb >= 1 and na >= 1 and len(blk.residual_m) == 1 and len(blk.atom_update.layers) >= 2
            ok &= 2 * len(paths) + 1 + 2 * (nb + na) <= chain.MAX_STEPS and 2 * len(blk.atom_update.layers) + 1 <= chain.MAX_STEPS
            res = list(blk.layers_before_skip) + list(blk.layers_after_skip) + list(blk.residual_m) + list(blk.atom_update.layers)[1:]
            ok &= all(fused.residual_ok(r) for r in res)
            dense = [blk.dense_ca, blk.atom_update.layers[0]]
            for pth in paths:
                dense += [pth.down_projection, pth.up_projection_ca, pth.up_projection_ac,
                          pth.dense_ba if hasattr(pth, "dense_ba") else pth.dense_db]
            ok &= all(fused.dense_ok(d) for d in dense)
            W = blk.concat_layer.dense.linear.weight
            ok &= blk.concat_layer.dense.linear.bias is None and isinstance(blk.concat_layer.dense._activation, layers.ScaledSiLU)
            A = blk.atom_update.layers[0].linear.weight.shape[0]
            ok &= chain.supported(W.shape[0], A, W.shape[1] - 2 * A)
        for ob in model.out_blocks:
            heads = [ob.seq_energy] + ([ob.seq_forces] if ob.direct_forces else [])
            for mlp in heads:
                ok &= len(mlp) >= 2 and fused.dense_ok(mlp[0]) and all(fused.residual_ok(r) for r in list(mlp)[1:])
                ok &= 2 * len(mlp) - 1 <= chain.MAX_STEPS
            ok &= ob.out_energy.linear.bias is None and chain.supported(ob.out_energy.linear.weight.shape[1])
    except AttributeError:
        ok = False
    model._fuse_ok = bool(ok)
    return model._fuse_ok


def side_stream(model):
    """the model's side streams (created on first use; graph capture calls this before it starts capturing):
    lane 1 = output blocks, lane 2 = the quadruplet branch of an interaction block, lane 3 = short independent
    branches inside an interaction block (the skip projection, the second up-projection)"""
    side = getattr(model, "_side_streams", None)
    if side is None:
        side = model._side_streams = (torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream())
    return side


def _lanes(model):
    """(main, side, side2) streams for the multi-lane evaluation, or None.  GEMNET_B200_LANES: "capture" (default) = only while
    a CUDA graph is being captured (eager evaluation is host-bound, a second stream only adds event calls),
    "1" = always, "0" = never."""
    mode = os.environ.get("GEMNET_B200_LANES", "capture")
    if mode == "0" or (mode == "capture" and not torch.cuda.is_current_stream_capturing()):
        return None
    n_side = int(os.environ.get("GEMNET_B200_SIDE_LANES", "2"))
    return (torch.cuda.current_stream(),) + side_stream(model)[:n_side]


def trainable(model):
    """Can this model's TRAINING step run on the engine?  direct-force models (no double backward needed) with the
    chain-fused block structure, one target, no force coupling."""
    return (model.direct_forces and not model.forces_coupled and model.num_targets == 1 and fuse_ok(model)
            and getattr(model, "chain_fusion", True) and getattr(model, "dense_algo", 1) == 1)


@torch.no_grad()
def backward_train(tape, E_mol, F_j, gE, gF, params):
    """dL/dparams from dL/dE, dL/dF through the tape of energy_and_forces(train=True)"""
    if gE is not None:
        tape.add_grad(E_mol, gE.contiguous())
    if gF is not None:
        tape.add_grad(F_j, gF.contiguous())
    tape.backward()
    ws = getattr(tape, "wstream", None)
    if ws is not None:
        torch.cuda.current_stream().wait_stream(ws)
    grads = tuple(tape._take(id(p)) for p in params)
    tape.keep.clear()
    return grads


@torch.no_grad()
def energy_and_forces(model, inputs, train=False):
    """GemNet.forward (gemnet.py:453-615): returns (E (nMol, T), F (nAtoms, 3) | (nAtoms, T, 3)); with train=True
    (direct-force models) also the tape, for backward_train."""
    plan = graph.get_plan(inputs, model.triplets_only)
    Z, R = inputs["Z"], inputs["R"]
    dev = R.device
    S, nR = model.num_spherical, model.num_radial
    need_forces = not model.direct_forces
    n_passes = model.num_targets if (need_forces and model.num_targets > 1) else 1
    forces = []
    E_mol = None
    streams = None if train else _lanes(model)
    if train:
        dense_ops.refresh()   # the optimiser changed every weight: ONE launch re-packs every tcgen05 image
    for target in range(n_passes):
        tape = Tape(enabled=need_forces or train, streams=streams, train=train)
        if need_forces:
            tape.require(R)
        if train:
            for prm in model.parameters():
                if prm.requires_grad:
                    tape.require(prm)
            if os.environ.get("GEMNET_B200_WGRAD_STREAM", "1") != "0":
                tape.wstream = side_stream(model)[0]
        wt = (lambda prm: prm) if train else (lambda prm: prm.detach())
        run = _Run(model, tape, plan)
        A = tape.apply

        D_ca, V_ca = A(ops._EdgeGeometry, R, plan.id_c, plan.id_a)
        basis = {}
        if not model.triplets_only:
            D_ab, _ = A(ops._EdgeGeometry, R, plan.id4_int_b, plan.id4_int_a)
            sph_abd, bd_proj = A(ops._QuadIntm, R, plan.abd_center, plan.abd_first, plan.abd_second, S, True)
            cs_cab, ac_proj = A(ops._QuadIntm, R, plan.cab_center, plan.cab_first, plan.cab_second, S, False)
            basis["ylm"] = A(ops._QuadYlm, cs_cab, ac_proj, bd_proj, plan.id4_reduce_cab, plan.id4_expand_abd, S)
            cb = model.cbf_basis
            rad_int = A(ops._SphBesselRadial, D_ab, S, nR, cb.cutoff, cb.p)                 # (nInt, S, R)
            rad_g = A(_Gather42, rad_int, plan.abd_int_edge, plan)                           # (M, S*R)
            cbf4 = A(_MulBroadcastSph, rad_g, sph_abd, nR)                                   # (M, S*R)
        rb = model.rbf_basis
        rbf = A(ops._BesselRbf, D_ca, wt(rb.frequencies), rb.cutoff, rb.p)
        c3 = model.cbf_basis3
        rad = A(ops._SphBesselRadial, D_ca, S, nR, c3.cutoff, c3.p)
        basis["sph3"] = A(ops._TripletSph, R, plan.id_c, plan.id_a, plan.id3_reduce, plan.id3_expand, S)

        h = A(_Embedding, model.atom_emb.embeddings.weight, Z) if train else model.atom_emb(Z)
        m = run.edge_embedding(model.edge_emb, h, rbf)
        use_fan = run.fuse and run.fuse_out and fused.rbf_fan_ok(model)
        if use_fan:  # every projection of the radial basis in one launch (13 / 17 K=16 GEMMs per direction otherwise)
            level1, uses = fused.fan_layout(model)
            outs = A(fused.RbfFan, rbf, level1, uses)
            run.fan_out = {id(d): o for (_, d), o in zip(uses, outs)}
        if not model.triplets_only:
            basis["rbf4"] = None if use_fan else run.linear(model.mlp_rbf4, rbf)
            basis["cbf4"] = run.linear(model.mlp_cbf4, cbf4)
            basis["rbf_W1_4"] = A(ops._DownProj, rad, wt(model.mlp_sbf4.weight), layers.l_of_s_tensor(S, True, dev))
        basis["rbf3"] = None if use_fan else run.linear(model.mlp_rbf3, rbf)
        basis["rbf_W1_3"] = A(ops._DownProj, rad, wt(model.mlp_cbf3.weight), layers.l_of_s_tensor(S, False, dev))
        basis["rbf_h"] = None if use_fan else run.linear(model.mlp_rbf_h, rbf)
        rbf_out = None if use_fan else run.linear(model.mlp_rbf_out, rbf)

        # The output blocks (mostly atom-level layers: a dozen CTAs per launch) only feed the final sums, so they run
        # on the side lane, next to the following interaction block — forward and backward.
        tape.keep.append(m)
        if run.fuse_out:
            E_a, F_ca = tape.on_side(lambda: run.output_block_fused(model.out_blocks[0], m, rbf_out, None, None))
        else:
            E_a, F_ca = tape.on_side(lambda: run.output_block(model.out_blocks[0], m, rbf_out))
        for i in range(model.num_blocks):
            h, m = run.interaction_block(model.int_blocks[i], h, m, basis, model.bilinear_algo)

            def out_i(m=m, E_a=E_a, F_ca=F_ca, blk=model.out_blocks[i + 1]):
                tape.keep.append(m)  # read on the side lane: must not be recycled by the main lane meanwhile
                if run.fuse_out:
                    return run.output_block_fused(blk, m, rbf_out, E_a, F_ca)
                E, F = run.output_block(blk, m, rbf_out)
                return run.add_scale(1.0, E_a, E), (F_ca + F if model.direct_forces else F_ca)

            E_a, F_ca = tape.on_side(out_i)
        tape.join()
        E_mol = A(ops._MoleculeSum, E_a, plan.batch_seg, plan.mol_rowptr, not model.extensive)

        if model.direct_forces:
            if model.forces_coupled:
                nE = plan.n_edges
                id_undir = inputs["id_undir"]
                summed = torch.zeros(int(nE / 2), F_ca.shape[1], dtype=F_ca.dtype, device=dev).index_add(0, id_undir, F_ca)
                cnt = torch.zeros(int(nE / 2), dtype=F_ca.dtype, device=dev).index_add(
                    0, id_undir, torch.ones(nE, dtype=F_ca.dtype, device=dev)).clamp(min=1)
                F_ca = (summed / cnt[:, None])[id_undir]
            F_j = A(ops._ForceReadout, F_ca, V_ca, plan.id_a, plan.atom_perm, plan.atom_rowptr)
            return (E_mol, F_j, tape) if train else (E_mol, F_j)

        seed = torch.zeros_like(E_mol)
        if model.num_targets > 1:
            seed[:, target] = 1.0
        else:
            seed.fill_(1.0)
        tape.add_grad(E_mol, seed)
        tape.backward()
        gR = tape._take(id(R))
        forces.append(-gR if gR is not None else torch.zeros_like(R))
    F_j = torch.stack(forces, dim=1) if model.num_targets > 1 else forces[0]
    return E_mol, F_j


# ---------------------------------------------------------------------------------------------- small extra ops (cbf4)
class _Gather42(torch.autograd.Function):
    """rows of the (nInt, S*R) radial table per intermediate triplet (basis_layers.py:134); S*R = 42 is not a
    multiple of 4, so this one gather stays an ATen index op (once per evaluation)."""

    @staticmethod
    def forward(ctx, table, idx, plan):
        # NB: the tape tracks tensors by identity, so views must be taken INSIDE ops, never between them
        ctx.idx, ctx.shape = idx.long(), table.shape
        return table.reshape(table.shape[0], -1)[ctx.idx]

    @staticmethod
    def backward(ctx, g):
        out = torch.zeros(ctx.shape[0], g.shape[1], dtype=g.dtype, device=g.device)
        return out.index_add_(0, ctx.idx, g).view(ctx.shape), None, None


class _MulBroadcastSph(torch.autograd.Function):
    """out[m, l*R + n] = rad[m, l*R + n] * sph[m, l]  (basis_layers.py:135-144)."""

    @staticmethod
    def forward(ctx, rad, sph, nR):
        ctx.save_for_backward(rad, sph)
        ctx.nR = nR
        return (rad.view(rad.shape[0], -1, nR) * sph[:, :, None]).reshape(rad.shape[0], -1)

    @staticmethod
    def backward(ctx, g):
        rad, sph = ctx.saved_tensors
        g3 = g.view(g.shape[0], -1, ctx.nR)
        return (g3 * sph[:, :, None]).reshape(g.shape[0], -1), (g3 * rad.view(rad.shape[0], -1, ctx.nR)).sum(-1), None
