"""TEST INFRASTRUCTURE — torch-CPU restatement of the reference model forward (rows H1-H13 of SURVEY §8a).

`OracleGemNet(kwargs, state_dict)` evaluates GemNet.forward on a parameter dict that uses the
reference's own state_dict key names, so the shipped `pretrained/GemNet-{T,Q}/model.pth` (when
present) and the seeded synthetic weights of `synthetic_state_dict` load as they are.

Follows (file:line relative to the reference root):
  gemnet/model/gemnet.py:262-286       calculate_interatomic_vectors        -> edge_vectors
  gemnet/model/gemnet.py:289-311       calculate_neighbor_angles            -> neighbor_angles
  gemnet/model/gemnet.py:314-332       vector_rejection                     -> vector_rejection
  gemnet/model/gemnet.py:335-418       calculate_angles (quadruplets)       -> quad_angles
  gemnet/model/gemnet.py:421-451       calculate_angles3                    -> triplet_angles
  gemnet/model/gemnet.py:453-615       GemNet.forward                       -> OracleGemNet.forward
  gemnet/model/layers/base_layers.py:5-89       Dense / ScaledSiLU / ResidualLayer   -> dense, residual
  gemnet/model/layers/embedding_block.py:7-75   AtomEmbedding / EdgeEmbedding        -> edge_embedding
  gemnet/model/layers/efficient.py:41-57        EfficientInteractionDownProjection   -> down_projection
  gemnet/model/layers/efficient.py:159-189      EfficientInteractionBilinear         -> efficient_bilinear
  gemnet/model/layers/interaction_block.py:158-234,363-422,517-566,653-696  blocks   -> interaction_block, ...
  gemnet/model/layers/atom_update_block.py:55-72,157-193  AtomUpdateBlock/OutputBlock -> atom_update, output_block
  gemnet/model/layers/scaling.py:170-174        ScalingFactor.forward (y * scale)

torch_scatter.scatter (un-vendored third-party dependency, unpinned in requirements.txt:10) is
restated as index_add (sum) / index_add + count divide (mean): its published semantics for
`dim=0, dim_size=n, reduce="add"|"mean"`.

The bilinear layer keeps the reference's *computational* form (zero-padded (E,Kmax,·) tensors and
three batched matmuls) so that timing this oracle on CPU is representative of the reference's
own CPU path (bench.py cpu_baseline / --impl reference).
"""
import math

import numpy as np
import torch

from . import basis_oracle as bo

INV_SQRT_2 = 1 / (2.0 ** 0.5)
INV_SQRT_3 = 1 / (3.0 ** 0.5)


# ----------------------------------------------------------------------------- small pieces
def scatter_sum(src, index, n):
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.index_add(0, index, src)


def scatter_mean(src, index, n):
    s = scatter_sum(src, index, n)
    cnt = torch.zeros(n, dtype=src.dtype, device=src.device).index_add(
        0, index, torch.ones_like(index, dtype=src.dtype)).clamp(min=1)
    return s / cnt.view((-1,) + (1,) * (src.dim() - 1))


def scaled_silu(x):
    return torch.nn.functional.silu(x) * (1 / 0.6)  # base_layers.py:51-58


def edge_vectors(R, id_s, id_t):
    V = R[id_t] - R[id_s]
    D = torch.sqrt(torch.sum(V ** 2, dim=1))
    return D, V / D[..., None]


def neighbor_angles(R_ac, R_ab):
    x = torch.sum(R_ac * R_ab, dim=1)
    y = torch.linalg.cross(R_ac, R_ab, dim=-1).norm(dim=-1)  # reference: torch.cross w/o dim
    y = torch.max(y, torch.tensor(1e-9, dtype=y.dtype))
    return torch.atan2(y, x)


def vector_rejection(R_ab, P_n):
    a_x_b = torch.sum(R_ab * P_n, dim=-1)
    b_x_b = torch.sum(P_n * P_n, dim=-1)
    return R_ab - (a_x_b / b_x_b)[:, None] * P_n


def triplet_angles(R, id_c, id_a, id3_reduce_ca, id3_expand_ba):
    Rc = R[id_c[id3_reduce_ca]]
    Ra = R[id_a[id3_reduce_ca]]
    Rb = R[id_c[id3_expand_ba]]
    return neighbor_angles(Rc - Ra, Rb - Ra)


def quad_angles(R, id_c, id_a, id4_int_b, id4_int_a, id4_expand_abd, id4_reduce_cab,
                id4_expand_intm_db, id4_reduce_intm_ca, id4_expand_intm_ab, id4_reduce_intm_ab):
    Ra = R[id4_int_a[id4_expand_intm_ab]]
    Rb = R[id4_int_b[id4_expand_intm_ab]]
    Rd = R[id_c[id4_expand_intm_db]]
    R_ba = Ra - Rb
    R_bd = Rd - Rb
    angle_abd = neighbor_angles(R_ba, R_bd)
    R_bd_proj = vector_rejection(R_bd, R_ba)[id4_expand_abd]

    Rc = R[id_c[id4_reduce_intm_ca]]
    Ra = R[id_a[id4_reduce_intm_ca]]
    Rb = R[id4_int_b[id4_reduce_intm_ab]]
    R_ac = Rc - Ra
    R_ab = Rb - Ra
    angle_cab = neighbor_angles(R_ab, R_ac)[id4_reduce_cab]
    R_ac_proj = vector_rejection(R_ac, R_ab)[id4_reduce_cab]
    angle_cabd = neighbor_angles(R_ac_proj, R_bd_proj)
    return angle_cab, angle_abd, angle_cabd


def pad_by_segment(values, id_reduce, Kidx, n_rows):
    """Zero padded dense (n_rows, Kmax, F) tensor (efficient.py:173-175, basis_layers.py:153-159)."""
    Kmax = 0 if values.shape[0] == 0 else int(torch.max(Kidx).item()) + 1
    out = torch.zeros(n_rows, Kmax, values.shape[1], dtype=values.dtype, device=values.device)
    out[id_reduce, Kidx] = values
    return out


def down_projection(rbf_env, sph_padded, weight):
    """efficient.py:41-57: rbf_env (S,E,R), sph (E,Kmax,S), weight (S,R,I) -> (E,I,S), (E,S,Kmax)."""
    rbf_W1 = torch.matmul(rbf_env, weight).permute(1, 2, 0)
    return rbf_W1, torch.transpose(sph_padded, 1, 2)


def efficient_bilinear(basis, m, id_reduce, Kidx, weight):
    """efficient.py:159-189: weight (C,I,O); m (U,C) already gathered; returns (E,O)."""
    rbf_W1, sph = basis
    nEdges = rbf_W1.shape[0]
    m2 = pad_by_segment(m, id_reduce, Kidx, nEdges)  # (E,Kmax,C)
    sum_k = torch.matmul(sph, m2)  # (E,S,C)
    rbf_W1_sum_k = torch.matmul(rbf_W1, sum_k)  # (E,I,C)
    m_ca = torch.matmul(rbf_W1_sum_k.permute(2, 0, 1), weight)  # (C,E,O)
    return torch.sum(m_ca, dim=0)


# ----------------------------------------------------------------------------- parameter bookkeeping
def param_shapes(kw):
    """Names and shapes of the UNIQUE tensors of a reference GemNet state_dict (the reference also stores
    every Dense weight twice, `X.weight` and `X.linear.weight`, and aliases `out_blocks.N.layers` as
    `out_blocks.N.seq_energy`; base_layers.py:24-27, atom_update_block.py:130)."""
    S, Rn = kw["num_spherical"], kw["num_radial"]
    A, Ed = kw["emb_size_atom"], kw["emb_size_edge"]
    Ct, Cq = kw["emb_size_trip"], kw["emb_size_quad"]
    rbf, cbf, sbf = kw["emb_size_rbf"], kw["emb_size_cbf"], kw["emb_size_sbf"]
    Ot, Oq = kw["emb_size_bil_trip"], kw["emb_size_bil_quad"]
    T = kw.get("num_targets", 1)
    trip_only = kw["triplets_only"]
    direct = kw.get("direct_forces", False)
    sh = {"rbf_basis.frequencies": (Rn,)}
    if not trip_only:
        sh["mlp_rbf4.weight"] = (rbf, Rn)
        sh["mlp_cbf4.weight"] = (cbf, Rn * S)
        sh["mlp_sbf4.weight"] = (S * S, Rn, sbf)
    sh["mlp_rbf3.weight"] = (rbf, Rn)
    sh["mlp_cbf3.weight"] = (S, Rn, cbf)
    sh["mlp_rbf_h.weight"] = (rbf, Rn)
    sh["mlp_rbf_out.weight"] = (rbf, Rn)
    sh["atom_emb.embeddings.weight"] = (93, A)
    sh["edge_emb.dense.weight"] = (Ed, 2 * A + Rn)

    def res(prefix, units):
        sh[prefix + ".dense_mlp.0.weight"] = (units, units)
        sh[prefix + ".dense_mlp.1.weight"] = (units, units)

    def atom_upd(prefix, units, key="layers"):
        sh[prefix + ".dense_rbf.weight"] = (Ed, rbf)
        sh[prefix + ".scale_sum.scale_factor"] = ()
        sh[prefix + f".{key}.0.weight"] = (units, Ed)
        for i in range(kw["num_atom"]):
            res(prefix + f".{key}.{i + 1}", units)

    for i in range(kw["num_blocks"] + 1):
        p = f"out_blocks.{i}"
        atom_upd(p, A)
        sh[p + ".out_energy.weight"] = (T, A)
        if direct:
            sh[p + ".scale_rbf.scale_factor"] = ()
            sh[p + ".seq_forces.0.weight"] = (Ed, Ed)
            for j in range(kw["num_atom"]):
                res(p + f".seq_forces.{j + 1}", Ed)
            sh[p + ".out_forces.weight"] = (T, Ed)
    for i in range(kw["num_blocks"]):
        p = f"int_blocks.{i}"
        sh[p + ".dense_ca.weight"] = (Ed, Ed)
        if not trip_only:
            q = p + ".quad_interaction"
            sh[q + ".dense_db.weight"] = (Ed, Ed)
            sh[q + ".mlp_rbf.weight"] = (Ed, rbf)
            sh[q + ".scale_rbf.scale_factor"] = ()
            sh[q + ".mlp_cbf.weight"] = (Cq, cbf)
            sh[q + ".scale_cbf.scale_factor"] = ()
            sh[q + ".mlp_sbf.weight"] = (Cq, sbf, Oq)
            sh[q + ".scale_sbf_sum.scale_factor"] = ()
            sh[q + ".down_projection.weight"] = (Cq, Ed)
            sh[q + ".up_projection_ca.weight"] = (Ed, Oq)
            sh[q + ".up_projection_ac.weight"] = (Ed, Oq)
        t = p + ".trip_interaction"
        sh[t + ".dense_ba.weight"] = (Ed, Ed)
        sh[t + ".mlp_rbf.weight"] = (Ed, rbf)
        sh[t + ".scale_rbf.scale_factor"] = ()
        sh[t + ".mlp_cbf.weight"] = (Ct, cbf, Ot)
        sh[t + ".scale_cbf_sum.scale_factor"] = ()
        sh[t + ".down_projection.weight"] = (Ct, Ed)
        sh[t + ".up_projection_ca.weight"] = (Ed, Ot)
        sh[t + ".up_projection_ac.weight"] = (Ed, Ot)
        for j in range(kw["num_before_skip"]):
            res(p + f".layers_before_skip.{j}", Ed)
        for j in range(kw["num_after_skip"]):
            res(p + f".layers_after_skip.{j}", Ed)
        atom_upd(p + ".atom_update", A)
        sh[p + ".concat_layer.dense.weight"] = (Ed, 2 * A + Ed)
        for j in range(kw["num_concat"]):
            res(p + f".residual_m.{j}", Ed)
    return sh


def expand_aliases(unique_sd):
    """Unique tensors -> full reference-format state_dict (duplicate `.linear.weight`, `seq_energy` alias)."""
    full = {}
    for k, v in unique_sd.items():
        full[k] = v
        is_dense = k.endswith(".weight") and v.dim() == 2 and not k.startswith("atom_emb.")
        if is_dense:
            full[k[: -len("weight")] + "linear.weight"] = v
    for k in list(full):
        if k.startswith("out_blocks.") and ".layers." in k:
            full[k.replace(".layers.", ".seq_energy.")] = full[k]
    return full


def synthetic_state_dict(kw, seed=1234, dtype=torch.float32):
    """Deterministic, reference-loadable weights (full state_dict format) from a seeded CPU generator.
    Variance ~ 1/fan_in like the reference's He-orthogonal init (initializers.py:27-40); scale factors in
    [0.4, 0.8] and 0.1x read-out weights, chosen so that energies are O(1) eV, forces O(10) eV/A and the model's
    own float32-vs-float64 deviation stays ~1e-5 relative (a Gaussian-weight network with larger scale
    factors amplifies rounding noise past the 1e-4 parity gate); Bessel frequencies perturbed off n*pi
    like the trained checkpoints."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = {}
    for name, shape in sorted(param_shapes(kw).items()):
        if name == "rbf_basis.frequencies":
            n = shape[0]
            v = math.pi * torch.arange(1, n + 1, dtype=torch.float64) * (
                1 + 0.05 * (torch.rand(n, generator=g, dtype=torch.float64) - 0.5))
        elif name.endswith("scale_factor"):
            v = 0.4 + 0.4 * torch.rand((), generator=g, dtype=torch.float64)
        elif name == "atom_emb.embeddings.weight":
            v = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * math.sqrt(3)
        else:
            fan_in = int(np.prod(shape[:-1])) if len(shape) == 3 else shape[1]
            v = torch.randn(shape, generator=g, dtype=torch.float64) / math.sqrt(fan_in)
            if ".out_energy." in name or ".out_forces." in name:
                v = v * 0.1
        sd[name] = v.to(dtype)
    return expand_aliases(sd)


# ----------------------------------------------------------------------------- the model
class OracleGemNet:
    def __init__(self, kwargs, state_dict, dtype=torch.float32):
        self.kw = dict(kwargs)
        self.dtype = dtype
        self.p = {k: v.detach().to(dtype).clone() for k, v in state_dict.items()}
        self.triplets_only = self.kw["triplets_only"]
        self.direct_forces = self.kw.get("direct_forces", False)
        self.num_blocks = self.kw["num_blocks"]
        self.cutoff = self.kw.get("cutoff", 5.0)
        self.int_cutoff = self.kw.get("int_cutoff", 10.0)
        self.envp = self.kw.get("envelope_exponent", 5)
        self.extensive = self.kw.get("extensive", True)
        self.forces_coupled = self.kw.get("forces_coupled", False)
        self.num_targets = self.kw.get("num_targets", 1)
        act = self.kw.get("activation", "swish")
        if act is not None and act.lower() not in ("swish", "silu"):
            raise NotImplementedError("Activation function not implemented for GemNet (yet).")
        self.trace = None  # optional dict collecting per-row intermediates for the parity tests

    # -- layers
    def dense(self, name, x, act):
        y = torch.nn.functional.linear(x, self.p[name + ".weight"])
        return scaled_silu(y) if act else y

    def residual(self, name, x):
        y = self.dense(name + ".dense_mlp.0", x, True)
        y = self.dense(name + ".dense_mlp.1", y, True)
        return (x + y) * INV_SQRT_2

    def scale(self, name, y):
        return y * self.p[name + ".scale_factor"]

    def edge_embedding(self, name, h, m_rbf, id_c_arg, id_a_arg):
        # embedding_block.py:70-74 — NB the call sites pass (id_c, id_a) into parameters named (idnb_a, idnb_c)
        x = torch.cat([h[id_c_arg], h[id_a_arg], m_rbf], dim=-1)
        return self.dense(name + ".dense", x, True)

    def atom_update(self, name, h, m, rbf, id_j, key="layers"):
        nAtoms = h.shape[0]
        x = m * self.dense(name + ".dense_rbf", rbf, False)
        x2 = scatter_sum(x, id_j, nAtoms)
        x = self.scale(name + ".scale_sum", x2)
        x = self.dense(name + f".{key}.0", x, True)
        for i in range(self.kw["num_atom"]):
            x = self.residual(name + f".{key}.{i + 1}", x)
        return x

    def output_block(self, name, h, m, rbf, id_j):
        nAtoms = h.shape[0]
        x = m * self.dense(name + ".dense_rbf", rbf, False)
        x_E = scatter_sum(x, id_j, nAtoms)
        x_E = self.scale(name + ".scale_sum", x_E)
        x_E = self.dense(name + ".layers.0", x_E, True)
        for i in range(self.kw["num_atom"]):
            x_E = self.residual(name + f".layers.{i + 1}", x_E)
        x_E = self.dense(name + ".out_energy", x_E, False)
        if self.direct_forces:
            x_F = self.scale(name + ".scale_rbf", x)
            x_F = self.dense(name + ".seq_forces.0", x_F, True)
            for i in range(self.kw["num_atom"]):
                x_F = self.residual(name + f".seq_forces.{i + 1}", x_F)
            x_F = self.dense(name + ".out_forces", x_F, False)
        else:
            x_F = 0
        return x_E, x_F

    def triplet_interaction(self, name, m, rbf3, cbf3, Kidx3, id_swap, id3_expand_ba, id3_reduce_ca):
        x_ba = self.dense(name + ".dense_ba", m, True)
        x_ba = self.scale(name + ".scale_rbf", x_ba * self.dense(name + ".mlp_rbf", rbf3, False))
        x_ba = self.dense(name + ".down_projection", x_ba, True)
        x_ba = x_ba[id3_expand_ba]
        x = efficient_bilinear(cbf3, x_ba, id3_reduce_ca, Kidx3, self.p[name + ".mlp_cbf.weight"])
        x = self.scale(name + ".scale_cbf_sum", x)
        if self.trace is not None:
            self.trace.setdefault("trip_bilinear_scaled", []).append(x.detach())
        x_ca = self.dense(name + ".up_projection_ca", x, True)
        x_ac = self.dense(name + ".up_projection_ac", x, True)[id_swap]
        return (x_ca + x_ac) * INV_SQRT_2

    def quadruplet_interaction(self, name, m, rbf, cbf, sbf, Kidx4, id_swap, id4_reduce_ca,
                               id4_expand_intm_db, id4_expand_abd):
        x_db = self.dense(name + ".dense_db", m, True)
        x_db = self.scale(name + ".scale_rbf", x_db * self.dense(name + ".mlp_rbf", rbf, False))
        x_db = self.dense(name + ".down_projection", x_db, True)
        x_db = x_db[id4_expand_intm_db]
        x_db = self.scale(name + ".scale_cbf", x_db * self.dense(name + ".mlp_cbf", cbf, False))
        x_db = x_db[id4_expand_abd]
        x = efficient_bilinear(sbf, x_db, id4_reduce_ca, Kidx4, self.p[name + ".mlp_sbf.weight"])
        x = self.scale(name + ".scale_sbf_sum", x)
        if self.trace is not None:
            self.trace.setdefault("quad_bilinear_scaled", []).append(x.detach())
        x_ca = self.dense(name + ".up_projection_ca", x, True)
        x_ac = self.dense(name + ".up_projection_ac", x, True)[id_swap]
        return (x_ca + x_ac) * INV_SQRT_2

    def interaction_block(self, name, h, m, g):
        x_ca_skip = self.dense(name + ".dense_ca", m, True)
        x3 = self.triplet_interaction(name + ".trip_interaction", m, g["rbf3"], g["cbf3"], g["Kidx3"],
                                      g["id_swap"], g["id3_expand_ba"], g["id3_reduce_ca"])
        if self.triplets_only:
            x = (x_ca_skip + x3) * INV_SQRT_2
        else:
            x4 = self.quadruplet_interaction(name + ".quad_interaction", m, g["rbf4"], g["cbf4"], g["sbf4"],
                                             g["Kidx4"], g["id_swap"], g["id4_reduce_ca"],
                                             g["id4_expand_intm_db"], g["id4_expand_abd"])
            x = (x_ca_skip + x3 + x4) * INV_SQRT_3
        for i in range(self.kw["num_before_skip"]):
            x = self.residual(name + f".layers_before_skip.{i}", x)
        m = (m + x) * INV_SQRT_2
        for i in range(self.kw["num_after_skip"]):
            m = self.residual(name + f".layers_after_skip.{i}", m)
        h2 = self.atom_update(name + ".atom_update", h, m, g["rbf_h"], g["id_a"])
        h = (h + h2) * INV_SQRT_2
        m2 = self.edge_embedding(name + ".concat_layer", h, m, g["id_c"], g["id_a"])
        for i in range(self.kw["num_concat"]):
            m2 = self.residual(name + f".residual_m.{i}", m2)
        m = (m + m2) * INV_SQRT_2
        return h, m

    # -- forward
    def forward(self, inputs, create_graph=False):
        kw = self.kw
        S, Rn = kw["num_spherical"], kw["num_radial"]
        Z = inputs["Z"]
        R = inputs["R"].to(self.dtype)
        id_a, id_c, id_undir, id_swap = inputs["id_a"], inputs["id_c"], inputs["id_undir"], inputs["id_swap"]
        id3_expand_ba, id3_reduce_ca, Kidx3 = inputs["id3_expand_ba"], inputs["id3_reduce_ca"], inputs["Kidx3"]
        batch_seg = inputs["batch_seg"]
        tr = self.trace
        if not self.direct_forces:
            R = R.detach().requires_grad_(True)

        D_ca, V_ca = edge_vectors(R, id_c, id_a)
        g = {"id_swap": id_swap, "id3_expand_ba": id3_expand_ba, "id3_reduce_ca": id3_reduce_ca,
             "Kidx3": Kidx3, "id_c": id_c, "id_a": id_a}
        nE = D_ca.shape[0]

        if not self.triplets_only:
            Kidx4 = inputs["Kidx4"]
            id4_reduce_ca = inputs["id4_reduce_ca"]
            D_ab, _ = edge_vectors(R, inputs["id4_int_b"], inputs["id4_int_a"])
            Phi_cab, Phi_abd, Theta_cabd = quad_angles(
                R, id_c, id_a, inputs["id4_int_b"], inputs["id4_int_a"], inputs["id4_expand_abd"],
                inputs["id4_reduce_cab"], inputs["id4_expand_intm_db"], inputs["id4_reduce_intm_ca"],
                inputs["id4_expand_intm_ab"], inputs["id4_reduce_intm_ab"])
            # cbf4: non-efficient SphericalBasisLayer at int_cutoff (basis_layers.py:133-144)
            rad_int = bo.radial_sph_bessel(D_ab, S, Rn, self.int_cutoff, self.envp)  # (nInt,S,R)
            cbf4 = (rad_int[inputs["id4_expand_intm_ab"]] * bo.legendre_y_l0(Phi_abd, S)[:, :, None]).reshape(-1, S * Rn)
            # sbf4: efficient TensorBasisLayer (basis_layers.py:241-295)
            rad = bo.radial_sph_bessel(D_ca, S, Rn, self.cutoff, self.envp)  # (E,S,R)
            deg = torch.arange(S) * 2 + 1
            rad49 = torch.repeat_interleave(rad, deg, dim=1).transpose(0, 1)  # (S^2,E,R)
            Ylm = bo.real_sph_harm_all(Phi_cab, Theta_cabd, S)  # (Q,S^2)
            sbf4 = (rad49, pad_by_segment(Ylm, id4_reduce_ca, Kidx4, nE))
            if tr is not None:
                tr.update(D_ab=D_ab.detach(), Phi_cab=Phi_cab.detach(), Phi_abd=Phi_abd.detach(),
                          Theta_cabd=Theta_cabd.detach(), cbf4=cbf4.detach(), Ylm=Ylm.detach())

        rbf = bo.bessel_rbf(D_ca, self.p["rbf_basis.frequencies"], self.cutoff, self.envp)
        Angles3 = triplet_angles(R, id_c, id_a, id3_reduce_ca, id3_expand_ba)
        rad3 = bo.radial_sph_bessel(D_ca, S, Rn, self.cutoff, self.envp)
        sph3 = bo.legendre_y_l0(Angles3, S)
        cbf3 = (rad3.transpose(0, 1), pad_by_segment(sph3, id3_reduce_ca, Kidx3, nE))
        if tr is not None:
            tr.update(D_ca=D_ca.detach(), V_ca=V_ca.detach(), rbf=rbf.detach(), Angles3=Angles3.detach(),
                      rad3=rad3.detach(), sph3=sph3.detach())

        h = self.p["atom_emb.embeddings.weight"][Z - 1]
        m = self.edge_embedding("edge_emb", h, rbf, id_c, id_a)

        if not self.triplets_only:
            g["rbf4"] = self.dense("mlp_rbf4", rbf, False)
            g["cbf4"] = self.dense("mlp_cbf4", cbf4, False)
            g["sbf4"] = down_projection(sbf4[0], sbf4[1], self.p["mlp_sbf4.weight"])
            g.update(Kidx4=Kidx4, id4_reduce_ca=id4_reduce_ca,
                     id4_expand_intm_db=inputs["id4_expand_intm_db"], id4_expand_abd=inputs["id4_expand_abd"])
        g["rbf3"] = self.dense("mlp_rbf3", rbf, False)
        g["cbf3"] = down_projection(cbf3[0], cbf3[1], self.p["mlp_cbf3.weight"])
        g["rbf_h"] = self.dense("mlp_rbf_h", rbf, False)
        rbf_out = self.dense("mlp_rbf_out", rbf, False)
        if tr is not None:
            tr.update(rbf_W1_3=g["cbf3"][0].detach(), m0=m.detach())

        E_a, F_ca = self.output_block("out_blocks.0", h, m, rbf_out, id_a)
        for i in range(self.num_blocks):
            h, m = self.interaction_block(f"int_blocks.{i}", h, m, g)
            E, F = self.output_block(f"out_blocks.{i + 1}", h, m, rbf_out, id_a)
            F_ca = F_ca + F
            E_a = E_a + E
        if tr is not None:
            tr.update(h_final=h.detach(), m_final=m.detach(), E_atom=E_a.detach())

        nMol = int(torch.max(batch_seg).item()) + 1
        E_mol = scatter_sum(E_a, batch_seg, nMol) if self.extensive else scatter_mean(E_a, batch_seg, nMol)

        if self.direct_forces:
            nAtoms = Z.shape[0]
            if self.forces_coupled:
                F_ca = scatter_mean(F_ca, id_undir, int(nE / 2))[id_undir]
            F_ji = F_ca[:, :, None] * V_ca[:, None, :]
            F_j = scatter_sum(F_ji, id_a, nAtoms)
        else:
            if self.num_targets > 1:
                F_j = torch.stack([-torch.autograd.grad(E_mol[:, i].sum(), R, create_graph=create_graph,
                                                        retain_graph=True)[0]
                                   for i in range(self.num_targets)], dim=1)
            else:
                F_j = -torch.autograd.grad(E_mol.sum(), R, create_graph=create_graph)[0]
        return E_mol, F_j

    def predict(self, inputs):
        E, F = self.forward(inputs)
        return E.detach(), F.detach()
