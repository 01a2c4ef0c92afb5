"""Index handling: the native host builder (row H0) and the per-batch device plan (CSR row pointers).

`build_batch` is the drop-in for DataContainer.__getitem__ / ase_calculator.Molecule.get
(gemnet/training/data_container.py:156-408; ase_calculator.py:23-99): same dict keys, int64 index tensors
and float32 R, built by the C++ builder in libgemnet_b200.so instead of scipy.sparse + numba.

`GraphPlan` converts the reference-format index tensors of one batch into what the kernels consume:
int32 copies, CSR row pointers over the reduce-sorted triplet / quadruplet arrays (instead of the reference's
Kidx + zero padding, efficient.py:173-175), the edge->atom CSR for deterministic atomic-free scatters, and the
atom index triples of the quadruplet geometry.  It is cached on the inputs dict.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib, ops

_PLAN_KEY = "_gemnet_b200_plan"
_CHECK_SORTED = os.environ.get("GEMNET_B200_CHECK_INDICES", "1") != "0"

TRIPLET_KEYS = ("id_undir", "id_swap", "id_c", "id_a", "id3_expand_ba", "id3_reduce_ca", "Kidx3")
QUAD_KEYS = ("id4_int_b", "id4_int_a", "id4_reduce_ca", "id4_expand_db", "id4_reduce_cab", "id4_expand_abd",
             "Kidx4", "id4_reduce_intm_ca", "id4_expand_intm_db", "id4_reduce_intm_ab", "id4_expand_intm_ab")


def build_batch(R_list, Z_list, cutoff=5.0, int_cutoff=10.0, triplets_only=True, as_torch=True, pin_memory=False):
    """Host index construction for a list of molecules (R_i (n_i,3) float64, Z_i (n_i,) int).

    Returns the reference's input dict (without E/F/id): N, batch_seg, Z, R, id_* ..., Kidx*.
    """
    lib = _lib.load()
    N = np.array([len(z) for z in Z_list], dtype=np.int32)
    R64 = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, 3) for r in R_list], axis=0))
    handle = ctypes.c_void_p()
    _lib.check(lib.gemnet_graph_build_host(R64.ctypes.data, N.ctypes.data, len(N), float(cutoff), float(int_cutoff),
                                           int(bool(triplets_only)), ctypes.byref(handle)), "gemnet_graph_build_host")
    try:
        data = {"N": N.astype(np.int64),
                "Z": np.concatenate([np.asarray(z) for z in Z_list]).astype(np.int64),
                "R": R64.astype(np.float32)}
        keys = ("batch_seg",) + TRIPLET_KEYS + (() if triplets_only else QUAD_KEYS)
        for k in keys:
            n = lib.gemnet_graph_size(handle, k.encode())
            if n < 0:
                raise _lib.GemnetB200Error(f"graph builder did not produce '{k}'")
            arr = np.empty(n, dtype=np.int32)
            _lib.check(lib.gemnet_graph_copy_host(handle, k.encode(), arr.ctypes.data), "gemnet_graph_copy_host")
            data[k] = arr
    finally:
        lib.gemnet_graph_free(handle)
    if not as_torch:
        return data
    out = {}
    for k, v in data.items():
        t = torch.from_numpy(v).to(torch.float32 if k == "R" else torch.int64)
        out[k] = t.pin_memory() if pin_memory else t
    return out


class Molecule:
    """Single-molecule input builder with the interface of ase_calculator.Molecule / predict.ipynb cell 2
    (ase_calculator.py:23-99): Molecule(R, Z, cutoff, int_cutoff, triplets_only).get()."""

    def __init__(self, R, Z, cutoff, int_cutoff, triplets_only=False):
        self.R = np.asarray(R, dtype=np.float64).reshape(-1, 3)
        self.Z = np.asarray(Z)
        self.cutoff, self.int_cutoff, self.triplets_only = cutoff, int_cutoff, triplets_only

    def update(self, R):
        self.R = np.asarray(R, dtype=np.float64).reshape(-1, 3)

    def get(self):
        return build_batch([self.R], [self.Z], self.cutoff, self.int_cutoff, self.triplets_only)


class GraphPlan:
    def __init__(self, inputs, triplets_only, check=_CHECK_SORTED):
        dev = inputs["R"].device
        if dev.type != "cuda":
            raise _lib.GemnetB200Error("GemNet (B200) needs its inputs on a CUDA device; there is no CPU fallback")

        def i32(k):
            return inputs[k].to(device=dev, dtype=torch.int32).contiguous()

        self.triplets_only = triplets_only
        self.n_atoms = int(inputs["Z"].shape[0])
        self.id_c, self.id_a, self.id_swap = i32("id_c"), i32("id_a"), i32("id_swap")
        self.id_c64, self.id_a64, self.id_swap64 = inputs["id_c"], inputs["id_a"], inputs["id_swap"]
        self.n_edges = int(self.id_c.numel())
        self.batch_seg = i32("batch_seg")
        if "N" in inputs:
            self.n_mol = int(inputs["N"].shape[0])
        else:  # the reference does the same device->host sync (gemnet.py:578)
            self.n_mol = int(inputs["batch_seg"].max().item()) + 1 if self.n_atoms > 0 else 0
        flags = []
        self.mol_rowptr, f = ops.csr_from_sorted(self.batch_seg, self.n_mol)
        flags.append(f)
        self.atom_rowptr, self.atom_perm = ops.csr_by_key(self.id_a, self.n_atoms)   # edges by target atom
        self.src_rowptr, self.src_perm = ops.csr_by_key(self.id_c, self.n_atoms)     # edges by source atom
        self.swap_rowptr, self.swap_perm = ops.csr_by_key(self.id_swap, self.n_edges)  # inverse of id_swap

        self.id3_reduce, self.id3_expand = i32("id3_reduce_ca"), i32("id3_expand_ba")
        self.n_triplets = int(self.id3_reduce.numel())
        self.rowptr3, f = ops.csr_from_sorted(self.id3_reduce, self.n_edges)
        flags.append(f)

        if not triplets_only:
            self.id4_reduce_ca = i32("id4_reduce_ca")
            self.id4_reduce_cab, self.id4_expand_abd = i32("id4_reduce_cab"), i32("id4_expand_abd")
            self.id4_expand_intm_db = i32("id4_expand_intm_db")
            self.intm_db_rowptr, self.intm_db_perm = ops.csr_by_key(self.id4_expand_intm_db, self.n_edges)
            self.n_quads = int(self.id4_reduce_ca.numel())
            self.rowptr4, f = ops.csr_from_sorted(self.id4_reduce_ca, self.n_edges)
            flags.append(f)
            self.id4_int_a, self.id4_int_b = i32("id4_int_a"), i32("id4_int_b")
            ia, ib = inputs["id4_int_a"], inputs["id4_int_b"]
            # atoms of the intermediate triplets (gemnet.py:385-411)
            e_ab, e_db = inputs["id4_expand_intm_ab"], inputs["id4_expand_intm_db"]
            r_ab, r_ca = inputs["id4_reduce_intm_ab"], inputs["id4_reduce_intm_ca"]
            # a - b <- d : center b, first a, second d
            self.abd_center = ib[e_ab].to(torch.int32).contiguous()
            self.abd_first = ia[e_ab].to(torch.int32).contiguous()
            self.abd_second = inputs["id_c"][e_db].to(torch.int32).contiguous()
            self.abd_int_edge = e_ab.to(torch.int32).contiguous()  # row of D_ab / rad_int per intermediate triplet
            # c -> a - b : center a, first b, second c
            self.cab_center = inputs["id_a"][r_ca].to(torch.int32).contiguous()
            self.cab_first = ib[r_ab].to(torch.int32).contiguous()
            self.cab_second = inputs["id_c"][r_ca].to(torch.int32).contiguous()
        if check:
            bad = torch.cat(flags).max().item()
            if bad:
                raise _lib.GemnetB200Error(
                    "index tensors are not grouped/sorted as the reference DataContainer produces them "
                    "(batch_seg, id3_reduce_ca, id4_reduce_ca must be non-decreasing and in range)")


def get_plan(inputs, triplets_only):
    plan = inputs.get(_PLAN_KEY)
    key = (inputs["id_c"].data_ptr(), inputs["id3_reduce_ca"].data_ptr(), int(inputs["id_c"].numel()),
           int(inputs["id3_reduce_ca"].numel()), triplets_only)
    if plan is None or plan[0] != key:
        plan = (key, GraphPlan(inputs, triplets_only))
        inputs[_PLAN_KEY] = plan
    return plan[1]
