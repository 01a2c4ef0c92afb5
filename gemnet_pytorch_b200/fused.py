"""Chain-fused blocks of the GemNet evaluation (N1): each class below is ONE gemnet_chain_tc launch in the forward
direction and ONE in the backward direction (forces are -dE/dR, gemnet.py:598-613), in the protocol the tape engine
(engine.py) consumes:

    forward(ctx, *args) -> tensor | tuple          backward_terms(ctx, tape, terms_per_output) -> grads per arg

`terms_per_output[i]` is the pending gradient of output i as a list of (tensor, scale, idx, mul) summands (idx / mul:
a lazily gathered / multiplied summand, see engine.Tape.add_grad) or None; a returned gradient is None, a tensor, or a
list of such summands — the consumer's operand fill sums them for free, so no add / scale / gather kernel runs
between the chains.

  TripPrologue   x  = down(ssilu(m W_ba^T) * r * s)                       interaction_block.py:653-676 / 517-560
  Post           m' = after_skip((m + before_skip((dense_ca(m) + sum_p (up_ca_p(x_p) + up_ac_p(x_p)[id_swap])/sqrt2) * c)) / sqrt2)
                                                                          interaction_block.py:186-221, 385-409, 684-696
  AtomChain      h' = (h + mlp(x0)) / sqrt2 ;  p_c = h' W_1^T, p_a = h' W_2^T
                                                                          atom_update_block.py:69-72, interaction_block.py:223-226,
                                                                          embedding_block.py:70-74 (atom part of the concat layer)
  EdgeEmb        m''= (m' + residual_m(ssilu(m' W_3^T + p_c[id_c] + p_a[id_a]))) / sqrt2
                                                                          embedding_block.py:70-74, interaction_block.py:228-233
  OutChain       E_a = E_prev + mlp(x0) w_out^T                           atom_update_block.py:175-193 (num_targets = 1)
  EdgeToAtomLazy the edge->atom scatter with a LAZY edge gradient         atom_update_block.py:64-68, 166-173
"""
import torch

from . import chain, dense_ops, layers, ops
from .chain import ACT, A_FROM_CARRY, CARRY_ADD, SET_CARRY, TO_A, Program, Step, Term

S2 = layers.INV_SQRT_2


def _new(rows, cols, like):
    return torch.empty((rows, cols), dtype=like.dtype, device=like.device)  # float32 on the device


def _pw(W, transpose=False):
    return dense_ops.packed_weight(W, transpose)


def _terms(tl):
    """pending gradient summands -> operand-fill terms"""
    return [Term(g, sc, idx=idx, mul=mul) for g, sc, idx, mul in tl]


def _wkey(st):
    W = st.weight[3]
    return (W.data_ptr(), tuple(W.shape), tuple(W.stride()))


def _save_operands(steps, rows, like, fwd=None):
    """training: every step stores its GEMM operand (the weight gradient's factor).  A step that reuses its
    predecessor's operand unchanged (no fill, predecessor without TO_A / A_FROM_CARRY) aliases the predecessor's."""
    prev = None
    for st in steps:
        if prev is not None and not st.terms and not (prev.flags & (TO_A | A_FROM_CARRY)):
            st.save_alias = prev if prev.save_x is not None else getattr(prev, "save_alias", None)
        else:
            st.save_x = _new(rows, st.K, like)
        prev = st
    return None


def _operand(st):
    if st.save_x is not None:
        return st.save_x
    al = getattr(st, "save_alias", None)
    return None if al is None else al.save_x


def _weight_grads(tape, fwd_steps, fstack, bwd_steps, gstack):
    """dW = (operand of the backward step)^T @ (operand of the forward step) for every weight both programs use:
    library GEMMs with the row count as reduction dimension (cuBLAS split-K), issued on the tape's weight-gradient
    stream so that they run beside the latency-bound chain kernels of the remaining backward pass.
    (One batched GEMM per chain was measured slower: 7 x (128 x 128) outputs with K = 8186 and no split-K.)"""
    fmap = {_wkey(st): _operand(st) for st in fwd_steps}
    pairs = []
    for st in bwd_steps:
        a, gz = fmap.get(_wkey(st)), _operand(st)
        if a is not None and gz is not None:
            pairs.append((st.weight[3], gz, a))

    def work():
        for W, gz, a in pairs:
            tape.add_wgrad(W, ops.weight_grad(gz, a))
    tape.on_wgrad_stream(work, [t for _, gz, a in pairs for t in (gz, a)])


def _is_silu(d):
    return isinstance(d._activation, layers.ScaledSiLU)


def residual_ok(layer):
    return (len(layer.dense_mlp) == 2 and all(_is_silu(d) and d.linear.bias is None and
                                               chain.supported(*d.linear.weight.shape) for d in layer.dense_mlp))


def dense_ok(d, act=True):
    return d.linear.bias is None and (_is_silu(d) if act else not _is_silu(d)) and chain.supported(*d.linear.weight.shape)


# --------------------------------------------------------------------------------------------------- residual stacks
def _residual_fwd(steps, saved, layer_list, need, rows, like, tail=None):
    """Append the two steps of every ResidualLayer of `layer_list` (base_layers.py:61-89): the input is the current
    operand AND the carry; afterwards the output is the operand and the carry again.  `tail` = dict of extra fields for
    the LAST step (a following skip connection folded in: scales multiplied, post term, stores)."""
    for j, layer in enumerate(layer_list):
        W1, W2 = layer.dense_mlp[0].linear.weight, layer.dense_mlp[1].linear.weight
        d1 = _new(rows, W1.shape[0], like) if need else None
        d2 = _new(rows, W2.shape[0], like) if need else None
        steps.append(Step(_pw(W1), flags=ACT | TO_A, dact=d1))
        kw = dict(flags=ACT | CARRY_ADD | SET_CARRY | TO_A, s_acc=S2, s_carry=S2, dact=d2)
        skip = S2
        if tail is not None and j == len(layer_list) - 1:
            extra = dict(tail)
            f = extra.pop("scale", 1.0)
            kw["s_acc"], kw["s_carry"], skip = S2 * f, S2 * f, S2 * f
            kw.update(extra)
        steps.append(Step(_pw(W2), **kw))
        saved.append((W1, W2, d1, d2, skip))


def _residual_bwd(steps, saved, first_terms, after):
    """Reverse of _residual_fwd.  Entry: either `first_terms` (the gradient arrives from global memory: operand fill
    with load_set_carry) or, if None, the previous step left the gradient in the carry and (times this stack's last d2)
    in the operand.  `after` = fields of the step that ends the reversed stack (v = gradient w.r.t. the stack's input):
    must provide what follows (next_amul / stores / flags)."""
    n = len(saved)
    for j in range(n - 1, -1, -1):
        W1, W2, d1, d2, skip = saved[j]
        kw = {}
        if j == n - 1 and first_terms is not None:
            kw = dict(terms=first_terms, load_set_carry=True, amul=d2)
        steps.append(Step(_pw(W2, True), flags=TO_A, next_amul=d1, **kw))
        if j > 0:
            steps.append(Step(_pw(W1, True), flags=CARRY_ADD | SET_CARRY | TO_A, s_carry=skip, next_amul=saved[j - 1][3]))
        else:
            a = dict(after)
            a["flags"] = a.get("flags", 0) | CARRY_ADD
            a["s_carry"] = skip
            steps.append(Step(_pw(W1, True), **a))


# --------------------------------------------------------------------------------------------------- fused blocks
class TripPrologue:
    """x = down_projection(ssilu(m W_ba^T) * r * s)   (interaction_block.py:658-671 triplets, 530-543 quadruplets)."""
    fused = True

    @staticmethod
    def forward(ctx, m, r, dense_in, down, s):
        Wb, Wd = dense_in.linear.weight, down.linear.weight
        E = m.shape[0]
        need = any(ctx.needs_input_grad[:2]) or ctx.train
        ctx.aux = _new(E, Wb.shape[0], m) if need else None
        ctx.d1 = _new(E, Wb.shape[0], m) if need else None
        ctx.d2 = _new(E, Wd.shape[0], m) if need else None
        ctx.Wb, ctx.Wd, ctx.E = Wb, Wd, E
        x = _new(E, Wd.shape[0], m)
        steps = [Step(_pw(Wb), terms=[Term(m)], flags=ACT | TO_A, s_acc=s, mul_early=r, aux=ctx.aux, dact=ctx.d1),
                 Step(_pw(Wd), flags=ACT, dact=ctx.d2, out=x)]
        ctx.fstack = _save_operands(steps, E, m) if ctx.train else None
        ctx.fsteps = steps
        chain.run(Program(E, steps))
        return x

    @staticmethod
    def backward_terms(ctx, tape, terms):
        gx = terms[0]
        if gx is None:
            return None, None, None, None, None
        need_m, need_r = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_r = _new(ctx.E, ctx.Wb.shape[0], ctx.aux) if need_r else None
        g_m = _new(ctx.E, ctx.Wb.shape[1], ctx.aux) if need_m else None
        steps = [Step(_pw(ctx.Wd, True), terms=_terms(gx), amul=ctx.d2, flags=TO_A, out=g_r, smul=ctx.aux,
                      next_amul=ctx.d1)]
        if need_m or ctx.train:
            steps.append(Step(_pw(ctx.Wb, True), out=g_m))
        gstack = _save_operands(steps, ctx.E, ctx.aux, (ctx.fsteps, ctx.fstack)) if ctx.train else None
        chain.run(Program(ctx.E, steps))
        if ctx.train:
            _weight_grads(tape, ctx.fsteps, ctx.fstack, steps, gstack)
        return g_m, g_r, None, None, None


class Post:
    """Everything of an InteractionBlock between the bilinear layers and the atom update (interaction_block.py:186-221,
    385-409, 684-696): up-projections of every path (the `x_ac[id_swap]` gather is a gathered operand fill of the
    up_projection_ac GEMM), dense_ca, the path merge, layers_before_skip, the skip with m, layers_after_skip."""
    fused = True

    @staticmethod
    def forward(ctx, m, x3, x4, blk, plan):
        paths = [(x3, blk.trip_interaction)] + ([(x4, blk.quad_interaction)] if x4 is not None else [])
        c = layers.INV_SQRT_2 if x4 is None else layers.INV_SQRT_3
        E = m.shape[0]
        need = any(ctx.needs_input_grad[:3]) or ctx.train
        H = blk.dense_ca.linear.weight.shape[0]
        steps, ups = [], []
        for j, (x, mod) in enumerate(paths):
            Wca, Wac = mod.up_projection_ca.linear.weight, mod.up_projection_ac.linear.weight
            da = _new(E, H, m) if need else None
            db = _new(E, H, m) if need else None
            steps.append(Step(_pw(Wca), terms=[Term(x)], flags=ACT | SET_CARRY | (CARRY_ADD if j else 0), s_acc=S2, dact=da))
            steps.append(Step(_pw(Wac), terms=[Term(x, idx=plan.id_swap)], flags=ACT | CARRY_ADD | SET_CARRY, s_acc=S2,
                              dact=db))
            ups.append((Wca, Wac, da, db))
        Wc = blk.dense_ca.linear.weight
        dc = _new(E, H, m) if need else None
        steps.append(Step(_pw(Wc), terms=[Term(m)], flags=ACT | CARRY_ADD | SET_CARRY | TO_A, s_acc=c, s_carry=c, dact=dc))
        before, after = [], []
        _residual_fwd(steps, before, list(blk.layers_before_skip), need, E, m, tail=dict(scale=S2, post0=m, s_post0=S2))
        out = _new(E, H, m)
        _residual_fwd(steps, after, list(blk.layers_after_skip), need, E, m, tail=dict(out=out))
        ctx.fstack = _save_operands(steps, E, m) if ctx.train else None
        ctx.fsteps = steps
        chain.run(Program(E, steps))
        ctx.save = (ups, Wc, dc, before, after, c, E, H, plan.id_swap, x4 is not None)
        return out

    @staticmethod
    def backward_terms(ctx, tape, terms):
        G = terms[0]
        if G is None:
            return None, None, None, None, None
        ups, Wc, dc, before, after, c, E, H, swap, has4 = ctx.save
        like = dc
        gm_part = _new(E, H, like)
        steps = []
        # after-skip stack reversed: ends with v = dL/dm1; m1 = (m + x1)/sqrt2 -> store the m-branch, carry on with x1
        _residual_bwd(steps, after, _terms(G),
                      dict(flags=SET_CARRY | TO_A, out=gm_part, s_store=S2, next_amul=before[-1][3]))
        # before-skip stack reversed: ends with v = dL/dx0 ; x0 = ssilu(m Wc^T)*c + c*sum_p(...)
        _residual_bwd(steps, before, None, dict(flags=SET_CARRY | TO_A, next_amul=dc))
        g_m = _new(E, H, like) if (ctx.needs_input_grad[0] or ctx.train) else None
        grads_x = []
        n_up = len(ups)
        # dense_ca branch, then the up-projections straight from the carry (dL/dx0)
        steps.append(Step(_pw(Wc, True), out=g_m, post0=gm_part, flags=A_FROM_CARRY, s_a=c, next_amul=ups[0][2]))
        for j, (Wca, Wac, da, db) in enumerate(ups):
            Cx = Wca.shape[1]
            ga, gb = _new(E, Cx, like), _new(E, Cx, like)
            steps.append(Step(_pw(Wca, True), out=ga, flags=A_FROM_CARRY, s_a=c, next_amul=db))
            nxt = ups[j + 1][2] if j + 1 < n_up else None
            steps.append(Step(_pw(Wac, True), out=gb, out_idx=swap, flags=(A_FROM_CARRY if nxt is not None else 0),
                              s_a=c, next_amul=nxt))
            grads_x.append([(ga, 1.0, None, None), (gb, 1.0, None, None)])
        gstack = _save_operands(steps, E, like, (ctx.fsteps, ctx.fstack)) if ctx.train else None
        chain.run(Program(E, steps))
        if ctx.train:
            _weight_grads(tape, ctx.fsteps, ctx.fstack, steps, gstack)
        return g_m, grads_x[0], (grads_x[1] if has4 else None), None, None


class AtomChain:
    """h' = (h + AtomUpdateBlock MLP(x0)) / sqrt2 and the atom part of the following concat layer, p_c = h' W_1^T,
    p_a = h' W_2^T (atom_update_block.py:69-72; interaction_block.py:223-226; embedding_block.py:70-74 with the weight
    split into its three column blocks as in layers.EdgeEmbedding)."""
    fused = True

    @staticmethod
    def forward(ctx, x0, h, blk):
        mlp = list(blk.atom_update.layers)
        W0 = mlp[0].linear.weight
        W = blk.concat_layer.dense.linear.weight
        A = h.shape[1]
        Wpc, Wpa = W[:, :A], W[:, A:2 * A]
        n = x0.shape[0]
        need = any(ctx.needs_input_grad[:2]) or ctx.train
        d0 = _new(n, W0.shape[0], x0) if need else None
        h2 = _new(n, A, x0)
        p_c, p_a = _new(n, W.shape[0], x0), _new(n, W.shape[0], x0)
        steps = [Step(_pw(W0), terms=[Term(x0)], flags=ACT | SET_CARRY | TO_A, dact=d0)]
        saved = []
        _residual_fwd(steps, saved, mlp[1:], need, n, x0, tail=dict(scale=S2, post0=h, s_post0=S2, out=h2))
        steps.append(Step(_pw(Wpc), out=p_c))
        steps.append(Step(_pw(Wpa), out=p_a))
        ctx.fstack = _save_operands(steps, n, x0) if ctx.train else None
        ctx.fsteps = steps
        chain.run(Program(n, steps))
        ctx.save = (W0, d0, saved, Wpc, Wpa, n, A)
        return h2, p_c, p_a

    @staticmethod
    def backward_terms(ctx, tape, terms):
        W0, d0, saved, Wpc, Wpa, n, A = ctx.save
        t_h, t_pc, t_pa = terms
        assert t_pc is not None and t_pa is not None
        t_h = tape.reduce_terms(t_h, 2, plain=True) if t_h else []
        need_x, need_h = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_h = _new(n, A, d0) if need_h else None
        g_x = _new(n, W0.shape[1], d0) if need_x else None
        post = {}
        for k, (g, sc, _, _) in enumerate(t_h):
            post[f"post{k}"], post[f"s_post{k}"] = g, sc
        steps = [Step(_pw(Wpc, True), terms=_terms(t_pc), flags=SET_CARRY),
                 Step(_pw(Wpa, True), terms=_terms(t_pa), flags=CARRY_ADD | SET_CARRY | TO_A, out=g_h, s_store=S2,
                      next_amul=saved[-1][3], **post)]
        _residual_bwd(steps, saved, None, dict(flags=TO_A, next_amul=d0))
        steps.append(Step(_pw(W0, True), out=g_x))
        gstack = _save_operands(steps, n, d0, (ctx.fsteps, ctx.fstack)) if ctx.train else None
        chain.run(Program(n, steps))
        if ctx.train:
            _weight_grads(tape, ctx.fsteps, ctx.fstack, steps, gstack)
        return g_x, g_h, None


class EdgeEmb:
    """m'' = (m' + residual_m(ssilu(m' W_3^T + p_c[id_c] + p_a[id_a]))) / sqrt2   (embedding_block.py:70-74 without
    the concatenation, interaction_block.py:228-233); num_concat = 1."""
    fused = True

    @staticmethod
    def forward(ctx, m, p_c, p_a, blk, plan):
        W = blk.concat_layer.dense.linear.weight
        A = (W.shape[1] - m.shape[1]) // 2
        Wm = W[:, 2 * A:]
        E, H = m.shape[0], W.shape[0]
        need = any(ctx.needs_input_grad[:3]) or ctx.train
        d0 = _new(E, H, m) if need else None
        out = _new(E, H, m)
        steps = [Step(_pw(Wm), terms=[Term(m)], pre0=p_c, pre0_idx=plan.id_c, pre1=p_a, pre1_idx=plan.id_a,
                      flags=ACT | SET_CARRY | TO_A, dact=d0)]
        saved = []
        _residual_fwd(steps, saved, list(blk.residual_m), need, E, m, tail=dict(scale=S2, post0=m, s_post0=S2, out=out))
        ctx.fstack = _save_operands(steps, E, m) if ctx.train else None
        ctx.fsteps = steps
        chain.run(Program(E, steps))
        ctx.save = (Wm, d0, saved, E, H, plan, p_c.shape[0])
        return out

    @staticmethod
    def backward_terms(ctx, tape, terms):
        G = terms[0]
        if G is None:
            return None, None, None, None, None
        Wm, d0, saved, E, H, plan, n_atoms = ctx.save
        G = tape.reduce_terms(G, chain.MAX_TERMS)
        g_z, g_m = _new(E, H, d0), _new(E, Wm.shape[1], d0)
        steps = []
        # carry = dL/dm'' for the whole program (one residual layer: nothing in between overwrites it)
        _residual_bwd(steps, saved, _terms(G), dict(flags=TO_A, mul_late=d0, out=g_z))
        steps.append(Step(_pw(Wm, True), flags=CARRY_ADD, s_carry=S2, out=g_m))
        gstack = _save_operands(steps, E, d0, (ctx.fsteps, ctx.fstack)) if ctx.train else None
        chain.run(Program(E, steps))
        if ctx.train:
            _weight_grads(tape, ctx.fsteps, ctx.fstack, steps, gstack)
        g_pc = ops.segment_sum_rows(g_z, plan.src_perm, plan.src_rowptr, n_atoms)
        g_pa = ops.segment_sum_rows(g_z, plan.atom_perm, plan.atom_rowptr, n_atoms)
        return g_m, g_pc, g_pa, None, None


class OutChain:
    """E_new = E_prev + MLP(x0) w_out^T for one target (atom_update_block.py:175-178 energy, :184-190 force branch)."""
    fused = True

    @staticmethod
    def forward(ctx, x0, E_prev, mlp, out_dense):
        mlp = list(mlp)
        W0 = mlp[0].linear.weight
        w_out = out_dense.linear.weight
        n = x0.shape[0]
        need = ctx.needs_input_grad[0] or ctx.train
        d0 = _new(n, W0.shape[0], x0) if need else None
        E_new = _new(n, 1, x0)
        steps = [Step(_pw(W0), terms=[Term(x0)], flags=ACT | SET_CARRY | TO_A, dact=d0)]
        saved = []
        ctx.v_last = _new(n, w_out.shape[1], x0) if ctx.train else None   # training: features the output layer sees
        _residual_fwd(steps, saved, mlp[1:], need, n, x0,
                      tail=dict(dot_w=w_out, dot_out=E_new, dot_in=E_prev, out=ctx.v_last))
        ctx.fstack = _save_operands(steps, n, x0) if ctx.train else None
        ctx.fsteps = steps
        chain.run(Program(n, steps))
        ctx.save = (W0, d0, saved, w_out, n)
        return E_new

    @staticmethod
    def backward_terms(ctx, tape, terms):
        gE = terms[0]
        if gE is None:
            return None, None, None, None
        W0, d0, saved, w_out, n = ctx.save
        gE = tape.reduce_terms(gE, 1, plain=True)
        g, sc, _, _ = gE[0]
        g_x = None
        if ctx.needs_input_grad[0] or ctx.train:
            g_x = _new(n, W0.shape[1], d0)
            steps = []
            _residual_bwd(steps, saved, [Term(g, sc, mul=w_out, kind=1)], dict(flags=TO_A, next_amul=d0))
            steps.append(Step(_pw(W0, True), out=g_x))
            gstack = _save_operands(steps, n, d0, (ctx.fsteps, ctx.fstack)) if ctx.train else None
            chain.run(Program(n, steps))
            if ctx.train:
                _weight_grads(tape, ctx.fsteps, ctx.fstack, steps, gstack)
                tape.add_wgrad(w_out, ops.weight_grad(g if sc == 1.0 else g * sc, ctx.v_last))
        return g_x, ([(g, sc, None, None)] if ctx.needs_input_grad[1] else None), None, None


class EdgeToAtomLazy:
    """ops._EdgeToAtom whose edge gradient stays LAZY: dL/dm[e] = scale * g_out[id_a[e]] * r[e] is handed on as a
    (tensor, scale, idx, mul) summand and evaluated inside the consumer's operand fill."""
    fused = True

    @staticmethod
    def forward(ctx, m, r, id_a, perm, rowptr, scale, had_scale):
        return ops._EdgeToAtom.forward(ctx, m, r, id_a, perm, rowptr, scale, had_scale)

    @staticmethod
    def backward_terms(ctx, tape, terms):
        m, r, id_a = ctx.saved_tensors
        scale, had_scale = ctx.cfg
        g_out = tape.reduce_terms(terms[0], 1, plain=True)[0] if terms[0] else None
        g_had = tape.reduce_terms(terms[1], 1, plain=True)[0] if (had_scale is not None and terms[1]) else None
        if g_out is None and g_had is None:
            return (None,) * 7
        g_m, g_r = [], None
        if g_out is not None:
            g_m.append((g_out[0], scale * g_out[1], id_a, r))
        if g_had is not None:
            g_m.append((g_had[0], had_scale * g_had[1], None, r))
        if ctx.needs_input_grad[1]:
            g_r = ops.edge_to_atom_grad_r(m, id_a, None if g_out is None else g_out[0],
                                          0.0 if g_out is None else scale * g_out[1],
                                          None if g_had is None else g_had[0],
                                          0.0 if g_had is None else had_scale * g_had[1])
        return (g_m if ctx.needs_input_grad[0] else None), g_r, None, None, None, None, None


class RbfFan:
    """Every projection of the Bessel basis in one launch (rbf_fan.cu): level 1 = the shared mlp_rbf* layers
    (gemnet.py:206-247), level 2 = the per-consumer layers (interaction_block.py:458-463,601-606;
    atom_update_block.py:36-41,128).  Returns one (E, H) tensor per `uses` entry."""
    fused = True

    @staticmethod
    def forward(ctx, rbf, level1, uses):
        W1s = [d.linear.weight for d in level1]
        W2s = [d.linear.weight for _, d in uses]
        groups = [g for g, _ in uses]
        out = ops.rbf_fan_fwd(rbf, W1s, W2s, groups)
        ctx.save = (rbf.shape[0], rbf.shape[1], W1s, W2s, groups)
        ctx.rbf = rbf if ctx.train else None
        return tuple(out[j] for j in range(len(uses)))

    @staticmethod
    def backward_terms(ctx, tape, terms):
        E, R, W1s, W2s, groups = ctx.save
        gs = []
        for t in terms:
            if not t:
                gs.append(None)
                continue
            g, sc, _, _ = tape.reduce_terms(t, 1, plain=True)[0]
            gs.append(g if sc == 1.0 else ops.lincomb([(g, sc)]))
        if all(g is None for g in gs):
            return None, None, None
        if ctx.train:   # weight gradients of both levels (plain GEMMs over the edges: library calls)
            rbf = ctx.rbf
            t = [torch.mm(rbf, W1.t()) for W1 in W1s]
            acc = [None] * len(W1s)
            for g, W2, gi in zip(gs, W2s, groups):
                if g is None:
                    continue
                tape.add_wgrad(W2, ops.weight_grad(g, t[gi]))
                tg = torch.mm(g, W2)
                acc[gi] = tg if acc[gi] is None else acc[gi] + tg
            for W1, tg in zip(W1s, acc):
                if tg is not None:
                    tape.add_wgrad(W1, ops.weight_grad(tg, rbf))
        g_rbf = ops.rbf_fan_bwd(E, R, W1s, W2s, groups, gs) if ctx.needs_input_grad[0] else None
        return g_rbf, None, None


def rbf_fan_ok(model):
    """the fan kernel covers emb_size_rbf = 16, equal consumer widths (multiple of 32, <= 512), <= 24 consumers"""
    try:
        uses = fan_layout(model)[1]
        H = {d.linear.weight.shape[0] for _, d in uses}
        return (model.mlp_rbf3.linear.weight.shape[0] == 16 and len(H) == 1 and next(iter(H)) % 32 == 0
                and next(iter(H)) <= 512 and len(uses) <= 24 and model.mlp_rbf3.linear.weight.shape[1] <= 64
                and all(d.linear.bias is None and not _is_silu(d) for _, d in uses))
    except AttributeError:
        return False


def fan_layout(model):
    """(level-1 layers, [(group, level-2 layer)]) of a GemNet: groups 0 rbf3, 1 rbf_h, 2 rbf_out, 3 rbf4"""
    level1 = [model.mlp_rbf3, model.mlp_rbf_h, model.mlp_rbf_out]
    if not model.triplets_only:
        level1.append(model.mlp_rbf4)
    uses = []
    for blk in model.int_blocks:
        uses.append((0, blk.trip_interaction.mlp_rbf))
        uses.append((1, blk.atom_update.dense_rbf))
        if not blk.triplets_only:
            uses.append((3, blk.quad_interaction.mlp_rbf))
    for ob in model.out_blocks:
        uses.append((2, ob.dense_rbf))
    return level1, uses
