"""Host-side mirror of the reference layer interface for the hot path.

Module / parameter names reproduce the reference's state_dict so that the shipped checkpoints load with
strict=True: every Dense stores its weight twice (`X.weight` and `X.linear.weight`,
base_layers.py:24-27), OutputBlock aliases `layers` as `seq_energy` (atom_update_block.py:130) and every
ScalingFactor is a scalar non-trainable Parameter (scaling.py:165-167).  Construction order and initialisers
follow the reference so that a seeded random init gives the same weights.

These `forward` methods are the LAYER-BY-LAYER path (`GemNet.forward_autograd`: torch.autograd over one kernel per
op from ops.py, dense layers as library GEMMs): used with `use_engine=False`, while a scale factor is being fitted
(scale_fit.py: the observed tensors must exist in HBM) and as the cross-check of the engine in the tests.  Inference and
direct-force training run on the tape engine instead (engine.py / fused.py: chain-fused tcgen05 kernels),
which reads the same modules' parameters.
"""

import numpy as np
import torch

from . import basis_consts, ops

INV_SQRT_2 = 1 / (2.0 ** 0.5)
INV_SQRT_3 = 1 / (3.0 ** 0.5)


def _w(module, param):
    """Parameters of the custom kernels: in eval mode they are constants (forces only need dE/dR), so their
    (expensive, reduction-over-all-edges) gradients are not computed; in training mode they are differentiated."""
    return param if (module.training and torch.is_grad_enabled()) else param.detach()


# ----------------------------------------------------------------------------- init (initializers.py:4-40)
def _standardize(kernel):
    eps = 1e-6
    axis = [0, 1] if kernel.dim() == 3 else 1
    var, mean = torch.var_mean(kernel, dim=axis, unbiased=True, keepdim=True)
    return (kernel - mean) / (var + eps) ** 0.5


def he_orthogonal_init(tensor):
    tensor = torch.nn.init.orthogonal_(tensor)
    fan_in = tensor.shape[:-1].numel() if tensor.dim() == 3 else tensor.shape[1]
    with torch.no_grad():
        tensor.data = _standardize(tensor.data)
        tensor.data *= (1 / fan_in) ** 0.5
    return tensor


# ----------------------------------------------------------------------------- dense stack (base_layers.py)
class ScaledSiLU(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.scale_factor = 1 / 0.6

    def forward(self, x):
        return ops.scaled_silu(x)  # silu(x) / 0.6 in one kernel (fwd) / one kernel (bwd)


class Dense(torch.nn.Module):
    def __init__(self, in_features, out_features, bias=False, activation=None, name=None):
        super().__init__()
        self.linear = torch.nn.Linear(in_features, out_features, bias=bias)
        self.reset_parameters()
        self.weight = self.linear.weight  # same Parameter under a second name, as the reference
        self.bias = self.linear.bias
        if isinstance(activation, str):
            activation = activation.lower()
        if activation in ["swish", "silu"]:
            self._activation = ScaledSiLU()
        elif activation is None:
            self._activation = torch.nn.Identity()
        else:
            raise NotImplementedError("Activation function not implemented for GemNet (yet).")

    def reset_parameters(self):
        he_orthogonal_init(self.linear.weight)
        if self.linear.bias is not None:
            self.linear.bias.data.fill_(0)

    def forward(self, x):
        return self._activation(self.linear(x))


class ResidualLayer(torch.nn.Module):
    def __init__(self, units, nLayers=2, activation=None, name=None):
        super().__init__()
        self.dense_mlp = torch.nn.Sequential(
            *[Dense(units, units, activation=activation, bias=False) for _ in range(nLayers)])

    def forward(self, inputs):
        x = inputs
        for layer in self.dense_mlp[:-1]:
            x = layer(x)
        last = self.dense_mlp[-1]
        z = last.linear(x)
        if isinstance(last._activation, ScaledSiLU):
            return ops.residual_tail(z, inputs, INV_SQRT_2)  # (inputs + silu(z)/0.6) / sqrt(2), one kernel
        return ops.add_scale(INV_SQRT_2, inputs, z)


class AtomEmbedding(torch.nn.Module):
    def __init__(self, emb_size, name=None):
        super().__init__()
        self.emb_size = emb_size
        self.embeddings = torch.nn.Embedding(93, emb_size)
        torch.nn.init.uniform_(self.embeddings.weight, a=-np.sqrt(3), b=np.sqrt(3))

    def forward(self, Z):
        return self.embeddings(Z - 1)


class EdgeEmbedding(torch.nn.Module):
    def __init__(self, atom_features, edge_features, out_features, activation=None, name=None):
        super().__init__()
        self.dense = Dense(2 * atom_features + edge_features, out_features, activation=activation, bias=False)

    def forward(self, h, m_rbf, plan):
        """dense(cat([h[id_c], h[id_a], m_rbf]))  (embedding_block.py:70-74; the call sites pass (id_c, id_a) for
        the parameters named (idnb_a, idnb_c)).  Evaluated without the concatenation: the weight is split into its
        three column blocks, the two atom blocks are applied on the nAtoms rows (10x fewer than nEdges) and gathered:
            z[e] = (h W_1^T)[id_c[e]] + (h W_2^T)[id_a[e]] + m_rbf[e] W_3^T."""
        W = self.dense.linear.weight
        A = h.shape[1]
        p_c = torch.nn.functional.linear(h, W[:, :A])
        p_a = torch.nn.functional.linear(h, W[:, A:2 * A])
        y = torch.nn.functional.linear(m_rbf, W[:, 2 * A:])
        z = ops.add_scale(1.0, y, ops.gather_rows(p_c, plan.id_c, plan.src_perm, plan.src_rowptr),
                          ops.gather_rows(p_a, plan.id_a, plan.atom_perm, plan.atom_rowptr))
        return self.dense._activation(z)


# ----------------------------------------------------------------------------- scaling (scaling.py:150-174)
from .scale_fit import ScalingFactor  # noqa: E402,F401  (the module tree keeps `layers.ScalingFactor`)


# ----------------------------------------------------------------------------- basis layers (basis_layers.py)
class BesselBasisLayer(torch.nn.Module):
    def __init__(self, num_radial, cutoff, envelope_exponent=5, name="bessel_basis"):
        super().__init__()
        self.num_radial = num_radial
        self.cutoff = float(cutoff)
        self.p = int(envelope_exponent)
        assert self.p > 0
        self.frequencies = torch.nn.Parameter(
            data=torch.Tensor(np.pi * np.arange(1, num_radial + 1, dtype=np.float32)), requires_grad=True)

    def forward(self, d):
        return ops.bessel_rbf(d, _w(self, self.frequencies), self.cutoff, self.p)


class SphericalBasisLayer(torch.nn.Module):
    """2-D basis.  The radial part (E,S,R) and the angular part (T,S) are returned separately and un-padded:
    the reference's zero-padded (E,Kmax,S) tensor (basis_layers.py:153-159) never exists here."""

    def __init__(self, num_spherical, num_radial, cutoff, envelope_exponent=5, efficient=False, name="spherical_basis"):
        super().__init__()
        assert num_radial <= 64
        self.num_spherical, self.num_radial = num_spherical, num_radial
        self.cutoff, self.p, self.efficient = float(cutoff), int(envelope_exponent), efficient

    def radial(self, D):
        return ops.sph_bessel_radial(D, self.num_spherical, self.num_radial, self.cutoff, self.p)


class TensorBasisLayer(SphericalBasisLayer):
    def __init__(self, num_spherical, num_radial, cutoff, envelope_exponent=5, efficient=False, name="tensor_basis"):
        super().__init__(num_spherical, num_radial, cutoff, envelope_exponent, efficient, name)


# ----------------------------------------------------------------------------- efficient.py
class EfficientInteractionDownProjection(torch.nn.Module):
    def __init__(self, num_spherical, num_radial, emb_size_interm, name="EfficientDownProj"):
        super().__init__()
        self.num_spherical, self.num_radial, self.emb_size_interm = num_spherical, num_radial, emb_size_interm
        self.reset_parameters()

    def reset_parameters(self):
        self.weight = torch.nn.Parameter(
            torch.empty((self.num_spherical, self.num_radial, self.emb_size_interm)), requires_grad=True)
        he_orthogonal_init(self.weight)

    def forward(self, rad, l_of_s):
        """rad (E, L, R) -> rbf_W1 (E, I, S)   (efficient.py:41-57)"""
        return ops.down_projection(rad, _w(self, self.weight), l_of_s)


class EfficientInteractionBilinear(torch.nn.Module):
    def __init__(self, emb_size, emb_size_interm, units_out, name="EfficientBilinear"):
        super().__init__()
        self.emb_size, self.emb_size_interm, self.units_out = emb_size, emb_size_interm, units_out
        self.reset_parameters()

    def reset_parameters(self):
        self.weight = torch.nn.Parameter(
            torch.empty((self.emb_size, self.emb_size_interm, self.units_out), requires_grad=True))
        he_orthogonal_init(self.weight)

    def forward(self, rbf_W1, sph, x, expand, rowptr, scale=1.0, algo=0):
        """x: UN-gathered (n_x, C) table; expand/rowptr: CSR of the units (efficient.py:159-189 + gather + scale)."""
        S = rbf_W1.shape[2]
        if algo == 1 and (S, self.emb_size, self.emb_size_interm, self.units_out) != (7, 64, 16, 64):
            algo = 0  # the tcgen05 kernel is built for the published triplet shape; other shapes use the fp32 kernel
        return ops.segmented_bilinear(x, sph, rbf_W1, _w(self, self.weight), expand, rowptr, scale, algo)


# ----------------------------------------------------------------------------- atom_update_block.py
class AtomUpdateBlock(torch.nn.Module):
    def __init__(self, emb_size_atom, emb_size_edge, emb_size_rbf, nHidden, activation=None, scale_file=None,
                 name="atom_update"):
        super().__init__()
        self.name = name
        self.emb_size_edge = emb_size_edge
        self.dense_rbf = Dense(emb_size_rbf, emb_size_edge, activation=None, bias=False)
        self.scale_sum = ScalingFactor(scale_file=scale_file, name=name + "_sum")
        self.layers = self.get_mlp(emb_size_atom, nHidden, activation)

    def get_mlp(self, units, nHidden, activation):
        dense1 = Dense(self.emb_size_edge, units, activation=activation, bias=False)
        res = [ResidualLayer(units, nLayers=2, activation=activation) for _ in range(nHidden)]
        return torch.nn.ModuleList([dense1] + res)

    def forward(self, h, m, rbf, plan):
        x = ops.edge_to_atom(m, self.dense_rbf(rbf), plan.id_a, plan.atom_perm, plan.atom_rowptr,
                             scale=self.scale_sum.value())
        self.scale_sum.observe(m, x)  # atom_update_block.py:68 (a no-op unless this factor is being fitted)
        for layer in self.layers:
            x = layer(x)
        return x


class OutputBlock(AtomUpdateBlock):
    def __init__(self, emb_size_atom, emb_size_edge, emb_size_rbf, nHidden, num_targets, activation=None,
                 direct_forces=True, output_init="HeOrthogonal", scale_file=None, name="output", **kwargs):
        super().__init__(name=name, emb_size_atom=emb_size_atom, emb_size_edge=emb_size_edge,
                         emb_size_rbf=emb_size_rbf, nHidden=nHidden, activation=activation, scale_file=scale_file,
                         **kwargs)
        assert isinstance(output_init, str)
        self.output_init = output_init
        self.direct_forces = direct_forces
        self.dense_rbf = Dense(emb_size_rbf, emb_size_edge, activation=None, bias=False)
        self.seq_energy = self.layers
        self.out_energy = Dense(emb_size_atom, num_targets, bias=False, activation=None)
        if self.direct_forces:
            self.scale_rbf = ScalingFactor(scale_file=scale_file, name=name + "_had")
            self.seq_forces = self.get_mlp(emb_size_edge, nHidden, activation)
            self.out_forces = Dense(emb_size_edge, num_targets, bias=False, activation=None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.output_init.lower() == "heorthogonal":
            he_orthogonal_init(self.out_energy.weight)
            if self.direct_forces:
                he_orthogonal_init(self.out_forces.weight)
        elif self.output_init.lower() == "zeros":
            torch.nn.init.zeros_(self.out_energy.weight)
            if self.direct_forces:
                torch.nn.init.zeros_(self.out_forces.weight)
        else:
            raise UserWarning(f"Unknown output_init: {self.output_init}")

    def forward(self, h, m, rbf, plan):
        r = self.dense_rbf(rbf)
        if self.direct_forces:
            x_E, x_F = ops.edge_to_atom(m, r, plan.id_a, plan.atom_perm, plan.atom_rowptr,
                                        scale=self.scale_sum.value(), had_scale=self.scale_rbf.value())
        else:
            x_E = ops.edge_to_atom(m, r, plan.id_a, plan.atom_perm, plan.atom_rowptr, scale=self.scale_sum.value())
            x_F = 0
        self.scale_sum.observe(m, x_E)        # atom_update_block.py:173
        if self.direct_forces:
            self.scale_rbf.observe(m, x_F)    # :181
        for layer in self.seq_energy:
            x_E = layer(x_E)
        x_E = self.out_energy(x_E)
        if self.direct_forces:
            for layer in self.seq_forces:
                x_F = layer(x_F)
            x_F = self.out_forces(x_F)
        return x_E, x_F


# ----------------------------------------------------------------------------- interaction_block.py
class TripletInteraction(torch.nn.Module):
    def __init__(self, emb_size_edge, emb_size_trip, emb_size_bilinear, emb_size_rbf, emb_size_cbf, activation=None,
                 scale_file=None, name="TripletInteraction", **kwargs):
        super().__init__()
        self.name = name
        self.dense_ba = Dense(emb_size_edge, emb_size_edge, activation=activation, bias=False)
        self.mlp_rbf = Dense(emb_size_rbf, emb_size_edge, activation=None, bias=False)
        self.scale_rbf = ScalingFactor(scale_file=scale_file, name=name + "_had_rbf")
        self.mlp_cbf = EfficientInteractionBilinear(emb_size_trip, emb_size_cbf, emb_size_bilinear)
        self.scale_cbf_sum = ScalingFactor(scale_file=scale_file, name=name + "_sum_cbf")
        self.down_projection = Dense(emb_size_edge, emb_size_trip, activation=activation, bias=False)
        self.up_projection_ca = Dense(emb_size_bilinear, emb_size_edge, activation=activation, bias=False)
        self.up_projection_ac = Dense(emb_size_bilinear, emb_size_edge, activation=activation, bias=False)

    def forward(self, m, rbf3, rbf_W1, sph3, plan, algo=0):
        x_in = self.dense_ba(m)
        x_ba = ops.mul_scale(x_in, self.mlp_rbf(rbf3), self.scale_rbf.value())
        self.scale_rbf.observe(x_in, x_ba)     # interaction_block.py:663
        x_ba = self.down_projection(x_ba)  # (E, emb_size_trip) — NOT gathered: the kernel gathers
        x = self.mlp_cbf(rbf_W1, sph3, x_ba, plan.id3_expand, plan.rowptr3, scale=self.scale_cbf_sum.value(),
                         algo=algo)
        self.scale_cbf_sum.observe(lambda: x_ba[plan.id3_expand.long()], x)  # :682 (x_ref = the gathered rows)
        x_ca = self.up_projection_ca(x)
        x_ac = ops.gather_rows(self.up_projection_ac(x), plan.id_swap, plan.swap_perm, plan.swap_rowptr)
        return ops.add_scale(INV_SQRT_2, x_ca, x_ac)


class QuadrupletInteraction(torch.nn.Module):
    def __init__(self, emb_size_edge, emb_size_quad, emb_size_bilinear, emb_size_rbf, emb_size_cbf, emb_size_sbf,
                 activation=None, scale_file=None, name="QuadrupletInteraction", **kwargs):
        super().__init__()
        self.name = name
        self.dense_db = Dense(emb_size_edge, emb_size_edge, activation=activation, bias=False)
        self.mlp_rbf = Dense(emb_size_rbf, emb_size_edge, activation=None, bias=False)
        self.scale_rbf = ScalingFactor(scale_file=scale_file, name=name + "_had_rbf")
        self.mlp_cbf = Dense(emb_size_cbf, emb_size_quad, activation=None, bias=False)
        self.scale_cbf = ScalingFactor(scale_file=scale_file, name=name + "_had_cbf")
        self.mlp_sbf = EfficientInteractionBilinear(emb_size_quad, emb_size_sbf, emb_size_bilinear)
        self.scale_sbf_sum = ScalingFactor(scale_file=scale_file, name=name + "_sum_sbf")
        self.down_projection = Dense(emb_size_edge, emb_size_quad, activation=activation, bias=False)
        self.up_projection_ca = Dense(emb_size_bilinear, emb_size_edge, activation=activation, bias=False)
        self.up_projection_ac = Dense(emb_size_bilinear, emb_size_edge, activation=activation, bias=False)

    def forward(self, m, rbf4, cbf4, rbf_W1, ylm, plan, algo=0):
        x_in = self.dense_db(m)
        x_db = ops.mul_scale(x_in, self.mlp_rbf(rbf4), self.scale_rbf.value())
        self.scale_rbf.observe(x_in, x_db)     # interaction_block.py:535
        x_db = self.down_projection(x_db)
        x_in = ops.gather_rows(x_db, plan.id4_expand_intm_db, plan.intm_db_perm, plan.intm_db_rowptr)  # (M, C4)
        x_db = ops.mul_scale(x_in, self.mlp_cbf(cbf4), self.scale_cbf.value())
        self.scale_cbf.observe(x_in, x_db)     # :549
        # second gather (id4_expand_abd) happens inside the kernel
        x = self.mlp_sbf(rbf_W1, ylm, x_db, plan.id4_expand_abd, plan.rowptr4, scale=self.scale_sbf_sum.value(),
                         algo=algo)
        self.scale_sbf_sum.observe(lambda: x_db[plan.id4_expand_abd.long()], x)  # :555
        x_ca = self.up_projection_ca(x)
        x_ac = ops.gather_rows(self.up_projection_ac(x), plan.id_swap, plan.swap_perm, plan.swap_rowptr)
        return ops.add_scale(INV_SQRT_2, x_ca, x_ac)


class InteractionBlock(torch.nn.Module):
    """GemNet-Q/dQ block (interaction_block.py:11-234)."""

    def __init__(self, emb_size_atom, emb_size_edge, emb_size_trip, emb_size_quad, emb_size_rbf, emb_size_cbf,
                 emb_size_sbf, emb_size_bil_trip, emb_size_bil_quad, num_before_skip, num_after_skip, num_concat,
                 num_atom, activation=None, scale_file=None, name="Interaction"):
        super().__init__()
        self.name = name
        block_nr = name.split("_")[-1]
        self.dense_ca = Dense(emb_size_edge, emb_size_edge, activation=activation, bias=False)
        self.quad_interaction = QuadrupletInteraction(
            emb_size_edge=emb_size_edge, emb_size_quad=emb_size_quad, emb_size_bilinear=emb_size_bil_quad,
            emb_size_rbf=emb_size_rbf, emb_size_cbf=emb_size_cbf, emb_size_sbf=emb_size_sbf, activation=activation,
            scale_file=scale_file, name=f"QuadInteraction_{block_nr}")
        self.trip_interaction = TripletInteraction(
            emb_size_edge=emb_size_edge, emb_size_trip=emb_size_trip, emb_size_bilinear=emb_size_bil_trip,
            emb_size_rbf=emb_size_rbf, emb_size_cbf=emb_size_cbf, activation=activation, scale_file=scale_file,
            name=f"TripInteraction_{block_nr}")
        self.layers_before_skip = torch.nn.ModuleList(
            [ResidualLayer(emb_size_edge, activation=activation) for _ in range(num_before_skip)])
        self.layers_after_skip = torch.nn.ModuleList(
            [ResidualLayer(emb_size_edge, activation=activation) for _ in range(num_after_skip)])
        self.atom_update = AtomUpdateBlock(emb_size_atom=emb_size_atom, emb_size_edge=emb_size_edge,
                                           emb_size_rbf=emb_size_rbf, nHidden=num_atom, activation=activation,
                                           scale_file=scale_file, name=f"AtomUpdate_{block_nr}")
        self.concat_layer = EdgeEmbedding(emb_size_atom, emb_size_edge, emb_size_edge, activation=activation)
        self.residual_m = torch.nn.ModuleList(
            [ResidualLayer(emb_size_edge, activation=activation) for _ in range(num_concat)])
        self.triplets_only = False

    def forward(self, h, m, basis, plan, algo=0):
        x_ca_skip = self.dense_ca(m)
        if self.triplets_only:
            x3 = self.trip_interaction(m, basis["rbf3"], basis["rbf_W1_3"], basis["sph3"], plan, algo)
            x = ops.add_scale(INV_SQRT_2, x_ca_skip, x3)
        else:
            x4 = self.quad_interaction(m, basis["rbf4"], basis["cbf4"], basis["rbf_W1_4"], basis["ylm"], plan, algo)
            x3 = self.trip_interaction(m, basis["rbf3"], basis["rbf_W1_3"], basis["sph3"], plan, algo)
            x = ops.add_scale(INV_SQRT_3, x_ca_skip, x3, x4)
        for layer in self.layers_before_skip:
            x = layer(x)
        m = ops.add_scale(INV_SQRT_2, m, x)
        for layer in self.layers_after_skip:
            m = layer(m)
        h2 = self.atom_update(h, m, basis["rbf_h"], plan)
        h = ops.add_scale(INV_SQRT_2, h, h2)
        m2 = self.concat_layer(h, m, plan)
        for layer in self.residual_m:
            m2 = layer(m2)
        m = ops.add_scale(INV_SQRT_2, m, m2)
        return h, m


class InteractionBlockTripletsOnly(InteractionBlock):
    """GemNet-T/dT block (interaction_block.py:237-422)."""

    def __init__(self, emb_size_atom, emb_size_edge, emb_size_trip, emb_size_quad, emb_size_rbf, emb_size_cbf,
                 emb_size_bil_trip, num_before_skip, num_after_skip, num_concat, num_atom, activation=None,
                 scale_file=None, name="Interaction", **kwargs):
        torch.nn.Module.__init__(self)
        self.name = name
        block_nr = name.split("_")[-1]
        self.dense_ca = Dense(emb_size_edge, emb_size_edge, activation=activation, bias=False)
        self.trip_interaction = TripletInteraction(
            emb_size_edge=emb_size_edge, emb_size_trip=emb_size_trip, emb_size_bilinear=emb_size_bil_trip,
            emb_size_rbf=emb_size_rbf, emb_size_cbf=emb_size_cbf, activation=activation, scale_file=scale_file,
            name=f"TripInteraction_{block_nr}")
        self.layers_before_skip = torch.nn.ModuleList(
            [ResidualLayer(emb_size_edge, activation=activation) for _ in range(num_before_skip)])
        self.layers_after_skip = torch.nn.ModuleList(
            [ResidualLayer(emb_size_edge, activation=activation) for _ in range(num_after_skip)])
        self.atom_update = AtomUpdateBlock(emb_size_atom=emb_size_atom, emb_size_edge=emb_size_edge,
                                           emb_size_rbf=emb_size_rbf, nHidden=num_atom, activation=activation,
                                           scale_file=scale_file, name=f"AtomUpdate_{block_nr}")
        self.concat_layer = EdgeEmbedding(emb_size_atom, emb_size_edge, emb_size_edge, activation=activation)
        self.residual_m = torch.nn.ModuleList(
            [ResidualLayer(emb_size_edge, activation=activation) for _ in range(num_concat)])
        self.triplets_only = True


def l_of_s_tensor(num_spherical, tensor_basis, device):
    return ops.const_tensor(f"l_of_s_{num_spherical}_{int(tensor_basis)}",
                            lambda: basis_consts.l_of_s(num_spherical, tensor_basis), device, torch.int32)
