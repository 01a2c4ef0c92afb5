"""GemNet with the reference's public interface, running its hot path on the sm_100a kernels.

Drop-in boundary: same constructor kwargs as gemnet/model/gemnet.py:82-113, same
`forward(inputs: dict) -> (E (nMol, num_targets), F (nAtoms, 3) | (nAtoms, num_targets, 3))` contract
(gemnet.py:453-615) including the `inputs["R"].requires_grad` side effect (:493-494,613), `predict`,
`load_weights`, `save_weights` (:780-790), the attributes `Trainer` touches (trainer.py:118-129,263-278,417) and a
state_dict that strict-loads the shipped checkpoints.

What differs inside: no zero-padded (E,Kmax,·) tensors, no gathered (T,C)/(Q,C) operands, angles are never
materialised (cos/sin go straight into the Legendre / Y_lm recurrences), spherical Bessel functions are evaluated in
fp64, scatters are CSR segment sums without atomics, and scale factors are folded into kernel epilogues.
"""
import torch

from . import _lib, dense_ops, engine, graph, layers, ops, scale_fit
from .layers import (AtomEmbedding, BesselBasisLayer, Dense, EdgeEmbedding, EfficientInteractionDownProjection,
                     InteractionBlock, InteractionBlockTripletsOnly, OutputBlock, SphericalBasisLayer,
                     TensorBasisLayer)


class GemNet(torch.nn.Module):
    def __init__(
        self,
        num_spherical: int,
        num_radial: int,
        num_blocks: int,
        emb_size_atom: int,
        emb_size_edge: int,
        emb_size_trip: int,
        emb_size_quad: int,
        emb_size_rbf: int,
        emb_size_cbf: int,
        emb_size_sbf: int,
        emb_size_bil_quad: int,
        emb_size_bil_trip: int,
        num_before_skip: int,
        num_after_skip: int,
        num_concat: int,
        num_atom: int,
        triplets_only: bool,
        num_targets: int = 1,
        direct_forces: bool = False,
        cutoff: float = 5.0,
        int_cutoff: float = 10.0,
        envelope_exponent: int = 5,
        extensive=True,
        forces_coupled: bool = False,
        output_init="HeOrthogonal",
        activation: str = "swish",
        scale_file=None,
        name="gemnet",
        **kwargs,
    ):
        super().__init__()
        assert num_blocks > 0
        self.num_targets = num_targets
        self.num_blocks = num_blocks
        self.extensive = extensive
        self.forces_coupled = forces_coupled
        self.direct_forces = direct_forces
        self.triplets_only = triplets_only
        self.num_spherical, self.num_radial = num_spherical, num_radial
        self.cutoff, self.int_cutoff = float(cutoff), float(int_cutoff)
        # 1 = tcgen05 forward kernel where built (published triplet shape), 0 = fp32 CUDA-core kernels everywhere
        self.bilinear_algo = int(kwargs.get("bilinear_algo", 1))
        # 1 = fused tcgen05 dense layers in the engine (dense_tc.cu: exact 3-way bf16 split, short accumulation
        # chains in separate TMEM accumulators -> unbiased, rms error 3e-8 vs cuBLAS fp32 1.5e-7); 0 = cuBLAS fp32 GEMMs
        # + fused pointwise kernels
        self.dense_algo = int(kwargs.get("dense_algo", 1))
        # eval(): evaluate energy+forces on the tape engine (engine.py) instead of torch.autograd (same kernels)
        self.use_engine = bool(kwargs.get("use_engine", True))

        # basis layers (no random state)
        self.rbf_basis = BesselBasisLayer(num_radial, cutoff=cutoff, envelope_exponent=envelope_exponent)
        if not self.triplets_only:
            self.cbf_basis = SphericalBasisLayer(num_spherical, num_radial, cutoff=int_cutoff,
                                                 envelope_exponent=envelope_exponent, efficient=False)
            self.sbf_basis = TensorBasisLayer(num_spherical, num_radial, cutoff=cutoff,
                                              envelope_exponent=envelope_exponent, efficient=True)
        self.cbf_basis3 = SphericalBasisLayer(num_spherical, num_radial, cutoff=cutoff,
                                              envelope_exponent=envelope_exponent, efficient=True)

        # shared down projections (construction order = reference, for seeded-init parity)
        if not self.triplets_only:
            self.mlp_rbf4 = Dense(num_radial, emb_size_rbf, activation=None, bias=False)
            self.mlp_cbf4 = Dense(num_radial * num_spherical, emb_size_cbf, activation=None, bias=False)
            self.mlp_sbf4 = EfficientInteractionDownProjection(num_spherical ** 2, num_radial, emb_size_sbf)
        self.mlp_rbf3 = Dense(num_radial, emb_size_rbf, activation=None, bias=False)
        self.mlp_cbf3 = EfficientInteractionDownProjection(num_spherical, num_radial, emb_size_cbf)
        self.mlp_rbf_h = Dense(num_radial, emb_size_rbf, activation=None, bias=False)
        self.mlp_rbf_out = Dense(num_radial, emb_size_rbf, activation=None, bias=False)

        self.atom_emb = AtomEmbedding(emb_size_atom)
        self.edge_emb = EdgeEmbedding(emb_size_atom, num_radial, emb_size_edge, activation=activation)

        block_cls = InteractionBlockTripletsOnly if self.triplets_only else InteractionBlock
        int_blocks = [
            block_cls(emb_size_atom=emb_size_atom, emb_size_edge=emb_size_edge, emb_size_trip=emb_size_trip,
                      emb_size_quad=emb_size_quad, emb_size_rbf=emb_size_rbf, emb_size_cbf=emb_size_cbf,
                      emb_size_sbf=emb_size_sbf, emb_size_bil_trip=emb_size_bil_trip,
                      emb_size_bil_quad=emb_size_bil_quad, num_before_skip=num_before_skip,
                      num_after_skip=num_after_skip, num_concat=num_concat, num_atom=num_atom, activation=activation,
                      scale_file=scale_file, name=f"IntBlock_{i + 1}")
            for i in range(num_blocks)
        ]
        out_blocks = [
            OutputBlock(emb_size_atom=emb_size_atom, emb_size_edge=emb_size_edge, emb_size_rbf=emb_size_rbf,
                        nHidden=num_atom, num_targets=num_targets, activation=activation, output_init=output_init,
                        direct_forces=direct_forces, scale_file=scale_file, name=f"OutBlock_{i}")
            for i in range(num_blocks + 1)
        ]
        self.out_blocks = torch.nn.ModuleList(out_blocks)
        self.int_blocks = torch.nn.ModuleList(int_blocks)
        # after the reference's own argument checks (same exceptions as gemnet.py / basis_layers.py for what IT rejects)
        _check_supported_shapes(num_spherical, num_radial, emb_size_atom, emb_size_edge, emb_size_trip, emb_size_quad,
                                emb_size_rbf, emb_size_cbf, emb_size_sbf, emb_size_bil_trip, emb_size_bil_quad,
                                triplets_only)

    # ------------------------------------------------------------------------------------------ weight caches
    def invalidate_caches(self):
        """Drop the packed (bf16-split) weight images of the tcgen05 kernels.  They are keyed on the tensors' version
        counters, which optimizer steps / copy_ / load_state_dict move but in-place writes through `.data` do not
        (ExponentialMovingAverage.copy_to / restore, ema_decay.py:110,145).  train()/eval() mode changes,
        load_state_dict and .to()/.cuda() call this, which covers the reference Trainer's cycle
        (load_averaged_variables -> eval -> restore_variable_backups -> train); call it yourself after any other `.data`
        write."""
        dense_ops.invalidate()
        for mod in self.modules():
            if isinstance(mod, layers.ScalingFactor):
                mod._cached = None
        self._fuse_ok = None

    def train(self, mode=True):
        if mode != self.training:
            self.invalidate_caches()
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_caches()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_caches()
        return out

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, inputs):
        if scale_fit.fitting_active():
            # fit_scaling.py: a scale factor is collecting Var(x_ref) / Var(y) — only the layer-by-layer path has those
            # tensors in HBM
            return self.forward_autograd(inputs)
        if self.use_engine and self.training and torch.is_grad_enabled() and engine.trainable(self):
            # training step of a direct-force model on the engine: the chain-fused kernels forward and backward, weight
            # gradients from the operands they saved (library GEMMs over the rows), bridged into torch.autograd so that
            # the reference Trainer's loss.backward() (trainer.py:346) works unchanged
            params = [p for p in self.parameters() if p.requires_grad]
            return _EngineTrainFn.apply(self, inputs, *params)
        if self.use_engine and not self.training:
            E_mol, F_j = engine.energy_and_forces(self, inputs)
            if not self.direct_forces:
                inputs["R"].requires_grad = False  # net effect of gemnet.py:493-494,613
            return E_mol, F_j
        return self.forward_autograd(inputs)

    def forward_autograd(self, inputs):
        """The same computation through torch.autograd (training, or use_engine=False)."""
        Z, R = inputs["Z"], inputs["R"]
        plan = graph.get_plan(inputs, self.triplets_only)
        S, nR = self.num_spherical, self.num_radial
        dev = R.device

        if not self.direct_forces:
            if self.training and torch.is_grad_enabled():
                # the reference differentiates the force THROUGH the force (create_graph=True, gemnet.py:598-613); the
                # kernels' backward passes are first-order only, so the force loss would silently contribute no weight
                # gradient: refuse instead of training on the energy alone
                raise NotImplementedError(
                    "training a direct_forces=False GemNet needs a double backward (SURVEY 8f N3), which the B200 "
                    "kernels do not provide: train with direct_forces=True, or evaluate with model.eval()")
            inputs["R"].requires_grad = True

        # H1: distances and unit vectors
        D_ca, V_ca = ops.edge_geometry(R, plan.id_c, plan.id_a)
        basis = {}

        if not self.triplets_only:
            # H2/H4/H5 for quadruplets: intermediate triplets, then per-quadruplet Y_lm
            D_ab, _ = ops.edge_geometry(R, plan.id4_int_b, plan.id4_int_a)
            sph_abd, bd_proj = ops.quad_intermediate(R, plan.abd_center, plan.abd_first, plan.abd_second, S, True)
            cs_cab, ac_proj = ops.quad_intermediate(R, plan.cab_center, plan.cab_first, plan.cab_second, S, False)
            ylm = ops.quad_ylm(cs_cab, ac_proj, bd_proj, plan.id4_reduce_cab, plan.id4_expand_abd, S)  # (Q, S^2)
            # cbf4 = non-efficient 2-D basis on the interaction edges (basis_layers.py:133-144): (M, S*R)
            rad_int = self.cbf_basis.radial(D_ab)  # (nInt, S, R)
            cbf4 = (rad_int[plan.abd_int_edge.long()] * sph_abd[:, :, None]).reshape(-1, S * nR)

        rbf = self.rbf_basis(D_ca)                                                   # H3  (E, R)
        rad = self.cbf_basis3.radial(D_ca)                                           # H4  (E, S, R), shared with H5
        sph3 = ops.triplet_sph(R, plan.id_c, plan.id_a, plan.id3_reduce, plan.id3_expand, S)  # H2+H4 (T, S)

        h = self.atom_emb(Z)
        m = self.edge_emb(h, rbf, plan)

        if not self.triplets_only:
            basis["rbf4"] = self.mlp_rbf4(rbf)
            basis["cbf4"] = self.mlp_cbf4(cbf4)
            basis["rbf_W1_4"] = self.mlp_sbf4(rad, layers.l_of_s_tensor(S, True, dev))      # H6 (E, I4, S^2)
            basis["ylm"] = ylm
        basis["rbf3"] = self.mlp_rbf3(rbf)
        basis["rbf_W1_3"] = self.mlp_cbf3(rad, layers.l_of_s_tensor(S, False, dev))          # H6 (E, I3, S)
        basis["sph3"] = sph3
        basis["rbf_h"] = self.mlp_rbf_h(rbf)
        rbf_out = self.mlp_rbf_out(rbf)

        E_a, F_ca = self.out_blocks[0](h, m, rbf_out, plan)
        for i in range(self.num_blocks):
            h, m = self.int_blocks[i](h, m, basis, plan, self.bilinear_algo)
            E, F = self.out_blocks[i + 1](h, m, rbf_out, plan)
            F_ca = F_ca + F
            E_a = E_a + E

        # H12: read-out
        E_mol = ops.molecule_sum(E_a, plan.batch_seg, plan.mol_rowptr, mean=not self.extensive)

        if self.direct_forces:
            if self.forces_coupled:  # enforce |F_ac| = |F_ca| (gemnet.py:588-592)
                nE = plan.n_edges
                id_undir = inputs["id_undir"]
                summed = torch.zeros(int(nE / 2), F_ca.shape[1], dtype=F_ca.dtype, device=dev).index_add(0, id_undir, F_ca)
                cnt = torch.zeros(int(nE / 2), dtype=F_ca.dtype, device=dev).index_add(
                    0, id_undir, torch.ones(nE, dtype=F_ca.dtype, device=dev)).clamp(min=1)
                F_ca = (summed / cnt[:, None])[id_undir]
            F_j = ops.force_readout(F_ca, V_ca, plan.id_a, plan.atom_perm, plan.atom_rowptr)  # (A, T, 3)
        else:
            # first-order kernels only: create_graph (double backward) is not available
            if self.num_targets > 1:
                forces = [-torch.autograd.grad(E_mol[:, i].sum(), inputs["R"], retain_graph=True)[0]
                          for i in range(self.num_targets)]
                F_j = torch.stack(forces, dim=1)
            else:
                F_j = -torch.autograd.grad(E_mol.sum(), inputs["R"], retain_graph=self.training)[0]
            inputs["R"].requires_grad = False
            if not self.training:
                E_mol = E_mol.detach()
        return E_mol, F_j

    def predict(self, inputs):
        E, F = self(inputs)
        return E.detach().cpu(), F.detach().cpu()

    def make_graphed(self, inputs, warmup=3):
        """Capture the whole energy+force evaluation of ONE batch topology (forward, and the autograd backward that
        produces the forces) into a CUDA graph.  Batch-64 GemNet is launch-latency bound; replaying
        ~1000 kernels as one graph removes the per-launch host cost.  Returns a GraphedGemNet whose call replays
        the graph on the static device buffers of `inputs` (new coordinates of the same topology can be copied into
        `inputs["R"]` first)."""
        return GraphedGemNet(self, inputs, warmup)

    def make_serving(self, capacity, warmup=3, device_build=False):
        """Static-shape serving: ONE CUDA graph (H2D of the batch, index plan, energy+force evaluation, D2H of E and
        F) that serves every batch fitting `capacity` (graph.BatchCapacity).  Returns a ServingGemNet; its call
        takes the reference-format HOST batch and returns (E, F) on the host.  `device_build=True`: the graph starts
        from raw coordinates instead (`call_molecules(R_list, Z_list)`; index construction on the GPU, row N2)."""
        return ServingGemNet(self, capacity, warmup, device_build)

    def make_serving_pool(self, molecules, max_atoms, growth=1.12, max_graphs=16, warmup=2):
        """Serving for a stream of batches of unknown size: capacities on a geometric ladder (each batch runs in
        the smallest bucket that holds it, so padding costs at most ~`growth` - 1 in edges and triplets), one captured
        graph per bucket, captured on first use.  Returns a ServingPool (same call convention as ServingGemNet)."""
        return ServingPool(self, molecules, max_atoms, growth, max_graphs, warmup)

    def make_molecule_server(self, molecules, max_atoms, workers=2, index_build="host", **pool_kwargs):
        """Serving from raw molecules: (R_list, Z_list) on the host -> (E, F) on the host.  The index construction
        (data_container.py:156-489) runs either on the host (`index_build="host"`, row H0: the native builder behind
        graph.build_batch on `workers` threads, pipelined with the GPU evaluation by MoleculeServer.stream()) or on the
        GPU (`index_build="device"`, row N2: csrc/graph_device.cu — inside the captured graph for triplet models); the
        evaluation is the static-shape CUDA graph pool of make_serving_pool.  This is what ase_calculator.py:148-170 does
        per MD step (rebuild indices, predict) as one call."""
        return MoleculeServer(self, molecules, max_atoms, workers, index_build, **pool_kwargs)

    def load_weights(self, path):
        self.load_state_dict(torch.load(path))

    def save_weights(self, path):
        torch.save(self.state_dict(), path)


def _check_supported_shapes(num_spherical, num_radial, emb_size_atom, emb_size_edge, emb_size_trip, emb_size_quad,
                            emb_size_rbf, emb_size_cbf, emb_size_sbf, emb_size_bil_trip, emb_size_bil_quad, triplets_only):
    """The reference accepts any hyperparameters (gemnet.py:82-113); this library's kernels are compiled for the
    published GemNet configurations.  Fail at construction with the list of what is built, not at the first forward
    after the weights were loaded."""
    bad = []
    if not 1 <= num_spherical <= 8:
        bad.append(f"num_spherical={num_spherical} (register-resident Legendre / Y_lm recurrences: 1..8)")
    if not 1 <= num_radial <= 16:
        bad.append(f"num_radial={num_radial} (radial-basis kernels: 1..16)")
    trip = (num_spherical, emb_size_trip, emb_size_cbf, emb_size_bil_trip)
    if trip != (7, 64, 16, 64):
        bad.append(f"(num_spherical, emb_size_trip, emb_size_cbf, emb_size_bil_trip)={trip}: the triplet bilinear "
                   "kernels (segbilinear*.cu) are built for (7, 64, 16, 64)")
    if not triplets_only:
        quad = (num_spherical ** 2, emb_size_quad, emb_size_sbf, emb_size_bil_quad)
        if quad != (49, 32, 32, 32):
            bad.append(f"(num_spherical^2, emb_size_quad, emb_size_sbf, emb_size_bil_quad)={quad}: the quadruplet "
                       "bilinear kernels are built for (49, 32, 32, 32)")
    for name, v in (("emb_size_atom", emb_size_atom), ("emb_size_edge", emb_size_edge), ("emb_size_rbf", emb_size_rbf),
                    ("emb_size_trip", emb_size_trip), ("emb_size_quad", emb_size_quad),
                    ("emb_size_bil_trip", emb_size_bil_trip), ("emb_size_bil_quad", emb_size_bil_quad)):
        if not (4 <= v <= 128 and v % 4 == 0):
            bad.append(f"{name}={v} (tcgen05 dense kernels: multiples of 4 in [4, 128])")
    if bad:
        raise _lib.GemnetB200Error("GemNet (B200): unsupported hyperparameters — " + "; ".join(bad))


class _EngineTrainFn(torch.autograd.Function):
    """(E, F) = engine.energy_and_forces(model, inputs, train=True) as one autograd node over all parameters."""

    @staticmethod
    def forward(ctx, model, inputs, *params):
        E, F, tape = engine.energy_and_forces(model, inputs, train=True)
        ctx.tape, ctx.outs, ctx.params = tape, (E, F), params
        return E, F

    @staticmethod
    def backward(ctx, gE, gF):
        grads = engine.backward_train(ctx.tape, ctx.outs[0], ctx.outs[1], gE, gF, ctx.params)
        ctx.tape = None
        return (None, None) + grads


class GraphedGemNet:
    """CUDA-graph replay of GemNet.forward for a fixed batch topology (inference / force evaluation)."""

    def __init__(self, model, inputs, warmup=3):
        assert not model.training, "graph capture is for inference (model.eval())"
        self.model, self.inputs = model, inputs
        engine.side_stream(model)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                model(inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.E, self.F = model(inputs)

    def __call__(self, R=None):
        if R is not None:
            self.inputs["R"].copy_(R)
        self.graph.replay()
        return self.E, self.F



class ServingGemNet:
    """CUDA-graph serving of GemNet-T/dT for batches of varying topology.

    Batch-64 GemNet is launch-latency bound, and an eager step spends more time in the host's ~480 launches than on
    the GPU.  A captured graph has all sizes baked in, so the batch is padded to a fixed capacity with inert
    entries (graph.pad_batch: a pad molecule whose atoms/edges/triplets only reference each other) directly in
    pinned staging buffers; the captured graph copies them to the device, builds the index plan (CSR row pointers,
    edge->atom orderings) with the plan kernels, evaluates energies and forces and copies E and F to pinned host
    buffers.  One cudaGraphLaunch per batch; the pad rows are cut off on return."""

    def __init__(self, model, capacity, warmup=3, device_build=False):
        assert not model.training, "serving capture is for inference (model.eval())"
        if not model.triplets_only:
            raise NotImplementedError("static-shape serving is built for triplets_only models (GemNet-T / dT)")
        self.model, self.cap = model, capacity
        dev = next(model.parameters()).device
        self.builder = None
        if device_build:
            # row N2: the captured graph starts from coordinates — H2D of (R float64, Z, atoms per molecule, sizes), the
            # index-construction kernels (csrc/graph_device.cu) and the pad kernel write the capacity-sized input dict
            self.builder = graph.DeviceStaticBuilder(capacity, model.cutoff, device=dev)
            self.builder.stage([], [])  # an all-pad batch for warm-up and capture
            self.host = self.builder.host
            self.hdr_host = torch.zeros(graph.HDR_WORDS, dtype=torch.int32).pin_memory()
        else:
            self.host = graph.serving_buffers(capacity, pin_memory=True)
            self.dev = {k: torch.empty_like(v, device=dev) for k, v in self.host.items()}
            empty = {k: v[:0] for k, v in self.host.items()}
            graph.pad_batch(empty, capacity, out=self.host)  # an all-pad batch for warm-up and capture
        self.E_host = self.F_host = self.flag_host = None

        def step():
            if self.builder is not None:
                self.builder.enqueue()
                inputs = self.builder.inputs()
                self.hdr_host.copy_(self.builder.a["hdr"], non_blocking=True)
            else:
                for k, h in self.host.items():
                    self.dev[k].copy_(h, non_blocking=True)
                inputs = dict(self.dev)  # a fresh dict: the index plan is rebuilt (inside the graph) for every batch
            inputs[graph._VALIDATED_KEY] = True
            inputs[graph._REBUILD_KEY] = True   # same device buffers, new contents: the plan kernels are part of the graph
            E, F = model(inputs)
            inputs[graph._REBUILD_KEY] = False
            flags = torch.cat(graph.get_plan(inputs, model.triplets_only).flags)  # device-side order/range check of the index lists
            if self.E_host is None:
                self.E_host = torch.empty(E.shape, dtype=E.dtype).pin_memory()
                self.F_host = torch.empty(F.shape, dtype=F.dtype).pin_memory()
                self.flag_host = torch.zeros(flags.shape, dtype=flags.dtype).pin_memory()
            self.E_host.copy_(E, non_blocking=True)
            self.F_host.copy_(F, non_blocking=True)
            self.flag_host.copy_(flags, non_blocking=True)

        engine.side_stream(model)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            step()
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.d2h_bytes = self.E_host.numel() * 4 + self.F_host.numel() * 4 + (0 if self.builder is None else 64)

    def call_molecules(self, R_list, Z_list):
        """(R_list, Z_list) on the HOST -> (E, F) host tensors; the index construction runs inside the captured graph
        (needs device_build=True).  Raises if the batch did not fit the capacity (the index arrays are then untouched)."""
        return self.call_arrays(*graph.concat_molecules(R_list, Z_list))

    def call_arrays(self, R64, Z, N):
        """the same from the concatenated batch (R float64 (A,3), Z (A,), N int32 (M,))"""
        self.launch_arrays(R64, Z, N)
        return self.finish()

    def launch_arrays(self, R64, Z, N):
        """first half of call_arrays: stage the batch and start the graph (asynchronous); `finish()` returns the result.
        Between the two the host is free — MoleculeServer.stream prepares the next batch there."""
        if self.builder is None:
            raise _lib.GemnetB200Error("this runner was captured for host-built batches (device_build=False)")
        self._sizes = self.builder.stage_arrays(R64, Z, N)
        self.graph.replay()

    def finish(self):
        A, M = self._sizes
        torch.cuda.current_stream().synchronize()
        ov = int(self.hdr_host[graph.HDR_OVERFLOW])
        if ov:
            raise _lib.GemnetB200Error(
                f"device index builder: batch does not fit {self.cap} (flags {ov}: 1 sizes inconsistent, 2 edges, "
                f"4 triplets, 64 no room for the pad molecule); counted edges {int(self.hdr_host[graph.HDR_EDGES])}, "
                f"triplets {int(self.hdr_host[graph.HDR_TRIPLETS])}")
        return self.E_host[:M].clone(), self.F_host[:A].clone()

    def __call__(self, inputs):
        """inputs: reference-format batch on the HOST (torch CPU tensors or numpy arrays) -> (E, F) host tensors."""
        if self.builder is not None:
            raise _lib.GemnetB200Error("this runner builds its indices on the device: use call_molecules(R_list, Z_list)")
        _, (A, _, _, M) = graph.pad_batch(inputs, self.cap, out=self.host, check=False)
        self.graph.replay()
        torch.cuda.current_stream().synchronize()
        if self.flag_host.any():  # the plan kernels validate what the host-side check would (no extra sync: same D2H)
            raise _lib.GemnetB200Error(graph.INDEX_ERROR)
        return self.E_host[:M].clone(), self.F_host[:A].clone()


class ServingPool:
    """ServingGemNet runners keyed by (edge capacity, triplet capacity) on a geometric ladder; least-recently-used
    eviction beyond `max_graphs` (a captured bucket keeps its activations' memory pool, ~1 GB at batch 64)."""

    def __init__(self, model, molecules, max_atoms, growth=1.12, max_graphs=16, warmup=2):
        assert growth > 1.0
        self.model, self.molecules, self.growth, self.max_graphs, self.warmup = model, molecules, growth, max_graphs, warmup
        self.atoms = -(-(int(max_atoms) + graph.BatchCapacity.PAD_ATOMS) // 8) * 8
        self.runners = {}  # insertion order = recency

    def _ladder(self, n, base):
        cap = base
        while cap < n:
            cap = int(cap * self.growth) + 1
        return -(-cap // 8) * 8

    def capacity_for(self, n_edges, n_triplets):
        return graph.BatchCapacity(self.atoms, self._ladder(n_edges + graph.BatchCapacity.PAD_EDGES, 1024),
                                   self._ladder(max(n_triplets, 1), 4096), self.molecules)

    def __call__(self, inputs):
        cap = self.capacity_for(int(inputs["id_c"].shape[0]), int(inputs["id3_reduce_ca"].shape[0]))
        key = (cap.edges, cap.triplets)
        runner = self.runners.pop(key, None)
        if runner is None:
            while len(self.runners) >= self.max_graphs:
                self.runners.pop(next(iter(self.runners)))
            runner = ServingGemNet(self.model, cap, self.warmup)
        self.runners[key] = runner
        return runner(inputs)

    def call_molecules(self, R_list, Z_list):
        """Raw molecules on the host -> (E, F): the bucket is chosen from the host's edge / triplet COUNT (one pass
        over the distance tests, graph.count_batch), the index arrays themselves are built inside the captured graph."""
        runner, arrays = self.prepare_molecules(R_list, Z_list)
        return runner.call_arrays(*arrays)

    def prepare_molecules(self, R_list, Z_list):
        """host work of a device-built batch that does not touch the staging buffers: concatenate, count edges and
        triplets, pick (or capture) the runner of the capacity bucket.  Returns (runner, (R64, Z, N))."""
        R64, Z, N = graph.concat_molecules(R_list, Z_list)
        n_edges, n_triplets = graph.count_arrays(R64, N, self.model.cutoff)
        cap = self.capacity_for(n_edges, n_triplets)
        key = (cap.edges, cap.triplets, "device")
        runner = self.runners.pop(key, None)
        if runner is None:
            while len(self.runners) >= self.max_graphs:
                self.runners.pop(next(iter(self.runners)))
            runner = ServingGemNet(self.model, cap, self.warmup, device_build=True)
        self.runners[key] = runner
        return runner, (R64, Z, N)

    @property
    def h2d_bytes(self):
        return max((r.h2d_bytes for r in self.runners.values()), default=0)

    @property
    def d2h_bytes(self):
        return max((r.d2h_bytes for r in self.runners.values()), default=0)


class MoleculeServer:
    """Host molecules in, energies and forces out: H0 (index construction) on worker threads + ServingPool on the GPU.

    __call__(R_list, Z_list)       one batch, H0 inline (latency path)
    stream(batches)                generator over an iterable of (R_list, Z_list): the index construction of the next
                                   `workers` batches overlaps the GPU evaluation of the current one (the native builder
                                   releases the GIL), results are yielded in order"""

    def __init__(self, model, molecules, max_atoms, workers=2, index_build="host", **pool_kwargs):
        from concurrent.futures import ThreadPoolExecutor
        if index_build not in ("host", "device"):
            raise ValueError("index_build must be 'host' or 'device'")
        self.model, self.index_build = model, index_build
        if model.triplets_only:
            self.pool = ServingPool(model, molecules, max_atoms, **pool_kwargs)
        else:
            self.pool = None  # quadruplet models: eager evaluation of the built batch
        self.workers = max(1, int(workers))
        self.exec = ThreadPoolExecutor(max_workers=self.workers)
        self.dev = next(model.parameters()).device

    def build(self, R_list, Z_list):
        m = self.model
        return graph.build_batch(R_list, Z_list, m.cutoff, m.int_cutoff, m.triplets_only)

    def evaluate(self, host_batch):
        if self.pool is not None:
            return self.pool(host_batch)
        return self.model.predict({k: v.to(self.dev, non_blocking=True) for k, v in host_batch.items()})

    def __call__(self, R_list, Z_list):
        if self.index_build == "device":
            # row N2: only coordinates / atomic numbers / molecule sizes cross PCIe.  Triplet models: index kernels inside
            # the captured graph; quadruplet models: dynamic shapes (two reads of the 64-byte size header), eager evaluation
            if self.pool is not None:
                return self.pool.call_molecules(R_list, Z_list)
            m = self.model
            return m.predict(graph.build_batch_device(R_list, Z_list, m.cutoff, m.int_cutoff, m.triplets_only, self.dev))
        return self.evaluate(self.build(R_list, Z_list))

    def stream(self, batches):
        if self.index_build == "device":
            if self.pool is None:  # quadruplet models: dynamic shapes, evaluated eagerly
                for R_list, Z_list in batches:
                    yield self(R_list, Z_list)
                return
            # triplet models: while the graph of batch n runs, the host concatenates / counts batch n + 1 (the staging
            # buffers are only written after batch n's results have been read)
            it = iter(batches)
            nxt = next(it, None)
            prep = None if nxt is None else self.pool.prepare_molecules(*nxt)
            while prep is not None:
                runner, arrays = prep
                runner.launch_arrays(*arrays)
                nxt = next(it, None)
                try:
                    prep = None if nxt is None else self.pool.prepare_molecules(*nxt)
                finally:
                    out = runner.finish()  # (also when preparing the next batch raised: never leave a replay un-synchronised)
                yield out
            return
        import collections
        pending = collections.deque()
        it = iter(batches)
        for _ in range(self.workers + 1):
            nxt = next(it, None)
            if nxt is None:
                break
            pending.append(self.exec.submit(self.build, *nxt))
        while pending:
            hb = pending.popleft().result()
            nxt = next(it, None)
            if nxt is not None:
                pending.append(self.exec.submit(self.build, *nxt))
            yield self.evaluate(hb)

    @property
    def h2d_bytes(self):
        return 0 if self.pool is None else self.pool.h2d_bytes

    @property
    def d2h_bytes(self):
        return 0 if self.pool is None else self.pool.d2h_bytes
