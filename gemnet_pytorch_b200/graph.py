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
_VALIDATED_KEY = "_gemnet_b200_indices_validated"  # set by pad_batch: the host already checked the index order
_REBUILD_KEY = "_gemnet_b200_rebuild_plan"  # serving: the static device buffers get new contents, rebuild the plan
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
            if as_torch:  # the reference's tensors are int64: widened by the library (multi-threaded), not by numpy
                arr = np.empty(n, dtype=np.int64)
                _lib.check(lib.gemnet_graph_copy_host_i64(handle, k.encode(), arr.ctypes.data),
                           "gemnet_graph_copy_host_i64")
            else:
                arr = np.empty(n, dtype=np.int32)
                _lib.check(lib.gemnet_graph_copy_host(handle, k.encode(), arr.ctypes.data), "gemnet_graph_copy_host")
            data[k] = arr
    finally:
        lib.gemnet_graph_free(handle)
    if not as_torch:
        return data
    out = {}
    for k, v in data.items():
        # widen in numpy (single-threaded memcpy speed); a torch .to() here fans out over every host thread and costs
        # milliseconds per array on many-core hosts
        t = torch.from_numpy(v if (k == "R" or v.dtype == np.int64) else v.astype(np.int64))
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


INDEX_ERROR = ("index tensors are not grouped/sorted as the reference DataContainer produces them (batch_seg, "
               "id3_reduce_ca, id4_reduce_ca must be non-decreasing) or an index lies outside its table "
               "(atom ids < n_atoms, edge ids < n_edges, ...)")


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
        # one flag for every range check of the gather / scatter lists (an id outside its table is skipped + reported)
        rng = torch.zeros(1, dtype=torch.int32, device=dev)
        flags.append(rng)
        self.atom_rowptr, self.atom_perm = ops.csr_by_key(self.id_a, self.n_atoms, rng)   # edges by target atom
        self.src_rowptr, self.src_perm = ops.csr_by_key(self.id_c, self.n_atoms, rng)     # edges by source atom
        self.swap_rowptr, self.swap_perm = ops.csr_by_key(self.id_swap, self.n_edges, rng)  # inverse of id_swap

        self.id3_reduce, self.id3_expand = i32("id3_reduce_ca"), i32("id3_expand_ba")
        self.n_triplets = int(self.id3_reduce.numel())
        self.rowptr3, f = ops.csr_from_sorted(self.id3_reduce, self.n_edges)
        flags.append(f)

        if not triplets_only:
            self.id4_reduce_ca = i32("id4_reduce_ca")
            self.id4_reduce_cab, self.id4_expand_abd = i32("id4_reduce_cab"), i32("id4_expand_abd")
            self.id4_expand_intm_db = i32("id4_expand_intm_db")
            self.intm_db_rowptr, self.intm_db_perm = ops.csr_by_key(self.id4_expand_intm_db, self.n_edges, rng)
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
        # the gather lists that no CSR construction has looked at (one launch)
        lists = [(self.id3_expand, self.n_edges)]
        if not triplets_only:
            lists += [(self.id4_reduce_cab, inputs["id4_reduce_intm_ca"].numel()),
                      (self.id4_expand_abd, self.id4_expand_intm_db.numel()),
                      (self.id4_int_a, self.n_atoms), (self.id4_int_b, self.n_atoms)]
        ops.index_range_check(lists, rng)
        self.flags = flags  # int32[1] each, non-zero = not grouped/sorted/in range; read by callers that skip the check here
        if check:
            bad = torch.cat(flags).max().item()
            if bad:
                raise _lib.GemnetB200Error(INDEX_ERROR)


_plans = {}  # (pointers, sizes, model kind) -> (weakref of the id_c tensor, plan); a few entries, least recent out


def get_plan(inputs, triplets_only):
    """The device plan of a batch, cached per index-tensor identity.  The cache is NOT kept in the caller's dict: the
    reference's Trainer.dict2device (trainer.py:312-318) calls `.to(device)` on every value of it."""
    import weakref
    id_c, red = inputs["id_c"], inputs["id3_reduce_ca"]
    key = (id_c.data_ptr(), red.data_ptr(), int(id_c.numel()), int(red.numel()), triplets_only)
    hit = None if inputs.get(_REBUILD_KEY, False) else _plans.get(key)
    if hit is not None and hit[0]() is id_c:       # same tensor object (not merely a recycled address)
        _plans[key] = _plans.pop(key)              # most recent last
        return hit[1]
    check = _CHECK_SORTED and not inputs.get(_VALIDATED_KEY, False)  # the device check ends in a host sync
    plan = GraphPlan(inputs, triplets_only, check=check)
    _plans.pop(key, None)
    while len(_plans) >= 4:
        _plans.pop(next(iter(_plans)))
    _plans[key] = (weakref.ref(id_c), plan)
    return plan


# ------------------------------------------------------------------------------------------------------------------
# Static-shape batches for CUDA-graph serving.
#
# A captured graph has every array size baked in, so a batch is padded up to a fixed capacity with INERT entries:
# a "pad molecule" (id n_molecules) of >= 3 atoms far away from everything, pad edges that only connect pad atoms
# and pad triplets that only connect pad edges.  Real entries keep their positions (pads are appended), messages
# never cross between real and pad entries (every scatter/gather follows the index lists), and the model has no
# batch-global reduction, so energies of the real molecules and forces on the real atoms are unchanged; the pad
# rows of the outputs are cut off by the caller.  Pad geometry is a non-degenerate triangle so that every pad value
# is finite as well.
class BatchCapacity:
    """Fixed sizes of a serving bucket: atoms, (directed) edges, triplets and molecules per batch."""

    PAD_ATOMS = 3
    PAD_EDGES = 6

    def __init__(self, atoms, edges, triplets, molecules):
        self.atoms, self.edges, self.triplets, self.molecules = int(atoms), int(edges), int(triplets), int(molecules)
        if self.edges % 2:
            raise ValueError("edge capacity must be even (edges come in directed pairs)")

    @classmethod
    def for_batches(cls, batches, margin=1.05):
        """Capacity that holds every batch of `batches` (reference-format dicts) with `margin` head-room."""
        a = max(int(b["Z"].shape[0]) for b in batches)
        e = max(int(b["id_c"].shape[0]) for b in batches)
        t = max(int(b["id3_reduce_ca"].shape[0]) for b in batches)
        m = max(int(b["N"].shape[0]) for b in batches)
        up = lambda n, q: -(-int(n * margin + 1) // q) * q  # noqa: E731
        return cls(up(a + cls.PAD_ATOMS, 8), up(e + cls.PAD_EDGES, 8), up(t, 8), m)

    def fits(self, n_atoms, n_edges, n_triplets, n_molecules):
        return (n_atoms + self.PAD_ATOMS <= self.atoms and n_edges + self.PAD_EDGES <= self.edges
                and n_triplets <= self.triplets and n_molecules <= self.molecules)

    def __repr__(self):
        return (f"BatchCapacity(atoms={self.atoms}, edges={self.edges}, triplets={self.triplets}, "
                f"molecules={self.molecules})")


_PAD_C = np.array([0, 1, 0, 2, 1, 2])        # source pad atom of pad-edge type j % 6
_PAD_A = np.array([1, 0, 2, 0, 2, 1])        # target pad atom; types (2k, 2k+1) are each other's reverse
_PAD_PARTNER = np.array([5, 3, 4, 1, 2, 0])  # an edge type with the same target and another source (a valid triplet)
_PAD_XYZ = np.array([[0.0, 0.0, 0.0], [1.2, 0.0, 0.0], [0.5, 1.1, 0.0]], dtype=np.float32)
SERVING_KEYS = ("Z", "R", "batch_seg", "N", "id_c", "id_a", "id_swap", "id_undir", "id3_expand_ba", "id3_reduce_ca")


def serving_buffers(cap, pin_memory=False):
    """Capacity-sized host arrays for `pad_batch` (torch tensors; pinned when they feed captured H2D copies)."""
    shapes = {"Z": (cap.atoms,), "R": (cap.atoms, 3), "batch_seg": (cap.atoms,), "N": (cap.molecules + 1,),
              "id_c": (cap.edges,), "id_a": (cap.edges,), "id_swap": (cap.edges,), "id_undir": (cap.edges,),
              "id3_expand_ba": (cap.triplets,), "id3_reduce_ca": (cap.triplets,)}
    out = {}
    for k, shp in shapes.items():
        # int32 staging: half the host copy and H2D bytes of the reference's int64 lists, and what the kernels consume
        t = torch.zeros(shp, dtype=torch.float32 if k == "R" else torch.int32)
        out[k] = t.pin_memory() if pin_memory else t
    return out


def _pad_tables(cap):
    """capacity-sized index tables of the pad pattern (computed once per capacity)"""
    t = getattr(cap, "_tables", None)
    if t is None:
        j = np.arange(cap.edges, dtype=np.int64)
        t = cap._tables = {"j": j, "c": _PAD_C[j % 6], "a": _PAD_A[j % 6], "grp": j // 6, "swap": j ^ 1, "half": j // 2,
                           "partner": (j // 6) * 6 + _PAD_PARTNER[j % 6],
                           "atom": np.arange(cap.atoms, dtype=np.int64)}
    return t


_NATIVE_PAD = os.environ.get("GEMNET_B200_NATIVE_PAD", "1") != "0"
_NATIVE_KEYS = ("Z", "R", "batch_seg", "N", "id_c", "id_a", "id_swap", "id_undir", "id3_expand_ba", "id3_reduce_ca")


def _pad_batch_native(inputs, cap, out):
    """gemnet_pad_batch_host (one native pass per array) when the batch is in the reference's tensor format: int64
    index tensors and float32 R, contiguous, on the host.  Returns the real sizes, or None (caller falls back)."""
    ptrs = []
    for k in _NATIVE_KEYS:
        v = inputs.get(k)
        want = torch.float32 if k == "R" else torch.int64
        if not (isinstance(v, torch.Tensor) and v.dtype is want and v.device.type == "cpu" and v.is_contiguous()):
            return None
        ptrs.append(v.data_ptr())
    for k in _NATIVE_KEYS:
        o = out[k]
        if not (o.dtype is (torch.float32 if k == "R" else torch.int32) and o.is_contiguous()):
            return None
    A, E, T, M = (int(inputs["Z"].shape[0]), int(inputs["id_c"].shape[0]), int(inputs["id3_reduce_ca"].shape[0]),
                  int(inputs["N"].shape[0]))
    lib = _lib.load()
    _lib.check(lib.gemnet_pad_batch_host(*ptrs, A, E, T, M, cap.atoms, cap.edges, cap.triplets, cap.molecules,
                                         *[out[k].data_ptr() for k in _NATIVE_KEYS]), "gemnet_pad_batch_host")
    return A, E, T, M


def pad_batch(inputs, cap, out=None, check=_CHECK_SORTED):
    """Write the triplet-model batch `inputs` (reference format, host) into capacity-sized buffers (see above).

    Returns (buffers, (n_atoms, n_edges, n_triplets, n_molecules)) — the real sizes, for cutting the outputs.
    `check`: validate on the host what GraphPlan otherwise validates on the device (grouping / index ranges)."""
    if out is not None and not check and _NATIVE_PAD:
        sizes = _pad_batch_native(inputs, cap, out)
        if sizes is not None:
            return out, sizes

    def arr(k):
        v = inputs[k]
        return v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)

    Z, R, seg, N = arr("Z"), arr("R"), arr("batch_seg"), arr("N")
    red, exp = arr("id3_reduce_ca"), arr("id3_expand_ba")
    A, E, T, M = Z.shape[0], inputs["id_c"].shape[0], red.shape[0], N.shape[0]
    if not cap.fits(A, E, T, M):
        raise _lib.GemnetB200Error(f"batch (atoms={A}, edges={E}, triplets={T}, molecules={M}) exceeds {cap}")
    if check and A:
        ok = ((seg[1:] >= seg[:-1]).all() and seg[0] >= 0 and seg[-1] < M and (red[1:] >= red[:-1]).all()
              and (T == 0 or (red[0] >= 0 and red[-1] < E)))
        if not ok:
            raise _lib.GemnetB200Error(
                "index tensors are not grouped/sorted as the reference DataContainer produces them "
                "(batch_seg, id3_reduce_ca must be non-decreasing and in range)")
    if out is None:
        out = serving_buffers(cap)
    tb = _pad_tables(cap)
    views = tb.setdefault("views", {})  # numpy views of the (long-lived) staging buffers, keyed by dict identity
    hit = views.get(id(out))
    if hit is None or hit[0] is not out:
        if len(views) > 8:
            views.clear()
        hit = views[id(out)] = (out, {k: v.numpy() for k, v in out.items()})
    o = hit[1]
    o["Z"][:A] = Z
    o["R"][:A] = R
    o["batch_seg"][:A] = seg
    for k in ("id_c", "id_a", "id_swap", "id_undir"):
        o[k][:E] = arr(k)
    o["id3_expand_ba"][:T] = exp
    o["id3_reduce_ca"][:T] = red
    # pad molecule: index cap.molecules (molecules M..cap-1 of a short batch stay empty)
    o["N"][:M] = N
    o["N"][M:cap.molecules] = 0
    o["N"][cap.molecules] = cap.atoms - A
    o["Z"][A:] = 1
    o["batch_seg"][A:] = cap.molecules
    # pad atoms form triangles (3 atoms each, 10 A apart, far from the real atoms); left-over atoms stay isolated
    n_pad = cap.atoms - A
    n_tri = max(1, n_pad // 3)
    far = (float(np.abs(R).max()) if A else 0.0) + 100.0
    k = tb["atom"][:n_pad]
    pos = _PAD_XYZ[np.where(k < 3 * n_tri, k % 3, 0)] + far
    pos[:, 0] += 10.0 * np.minimum(k // 3, n_tri - 1)
    o["R"][A:] = pos
    # pad edges: groups of 6 = the directed edges of one triangle (so the per-atom degree stays molecule-like)
    n_pe = cap.edges - E
    base = A + 3 * (tb["grp"][:n_pe] % n_tri)
    o["id_c"][E:] = base + tb["c"][:n_pe]
    o["id_a"][E:] = base + tb["a"][:n_pe]
    o["id_swap"][E:] = E + tb["swap"][:n_pe]
    o["id_undir"][E:] = E // 2 + tb["half"][:n_pe]
    # pad triplets: `per` of them on each pad edge of the complete groups, CSR order kept (reduce ids
    # non-decreasing); the expand edge is the edge of the same group with the same target atom and the other source
    nt = cap.triplets - T
    if nt:
        n_full = (n_pe // 6) * 6
        per = -(-nt // n_full)
        r_local = np.repeat(tb["j"][:-(-nt // per)], per)[:nt]
        o["id3_reduce_ca"][T:] = E + r_local
        o["id3_expand_ba"][T:] = E + tb["partner"][r_local]
    return out, (A, E, T, M)


# ------------------------------------------------------------------------------------------------------------------
# N2: index construction on the device (csrc/graph_device.cu; include/gemnet_b200.h, gemnet_device_graph).
class DeviceGraphDesc(ctypes.Structure):
    """ctypes mirror of `gemnet_device_graph`."""
    _fields_ = ([("R", ctypes.c_void_p), ("n_per_mol", ctypes.c_void_p), ("hdr", ctypes.c_void_p),
                 ("cutoff", ctypes.c_double), ("int_cutoff", ctypes.c_double), ("triplets_only", ctypes.c_int32)]
                + [(k, ctypes.c_int64) for k in ("cap_atoms", "cap_molecules", "cap_edges", "cap_triplets", "cap_int_edges",
                                       "cap_intm", "cap_quads")]
                + [(k, ctypes.c_void_p) for k in ("mol_off", "atom_cnt", "upper_off", "in_ptr", "int_off", "incoming", "incoming_pos",
                                     "intm_off_a", "intm_off_b", "quad_ptr", "R32", "batch_seg", "id_c", "id_a",
                                     "id_undir", "id_swap", "id3_expand_ba", "id3_reduce_ca", "Kidx3", "rowptr3",
                                     "id4_int_a", "id4_int_b", "id4_reduce_intm_ca", "id4_reduce_intm_ab",
                                     "id4_expand_intm_db", "id4_expand_intm_ab", "id4_reduce_ca", "id4_expand_db",
                                     "id4_reduce_cab", "id4_expand_abd", "Kidx4")])


HDR_WORDS = 16
HDR_ATOMS, HDR_MOLS, HDR_UPPER, HDR_EDGES, HDR_TRIPLETS, HDR_INT_EDGES, HDR_INTM, HDR_QUADS, HDR_OVERFLOW = range(9)
_I32_MAX = 2 ** 31 - 1
_EDGE_ARRAYS = ("id_c", "id_a", "id_undir", "id_swap", "incoming", "incoming_pos")
_TRIPLET_ARRAYS = ("id3_expand_ba", "id3_reduce_ca", "Kidx3")
_INT_ARRAYS = ("id4_int_a", "id4_int_b")
_INTM_ARRAYS = ("id4_reduce_intm_ca", "id4_reduce_intm_ab", "id4_expand_intm_db", "id4_expand_intm_ab")
_QUAD_ARRAYS = ("id4_reduce_ca", "id4_expand_db", "id4_reduce_cab", "id4_expand_abd", "Kidx4")


class _CudaBackend:
    """Arrays are torch CUDA tensors, the stages are the C-ABI entry points on the current stream."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.GemnetB200Error("the device index builder needs a CUDA device; there is no CPU fallback "
                                       "(graph.build_batch is the host builder)")
        self.lib = _lib.load()

    def empty(self, n, dtype=torch.int32):
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)[:int(n)]  # never a null pointer

    @staticmethod
    def host_buffer(n, dtype=torch.int32):
        return torch.zeros(max(int(n), 1), dtype=dtype).pin_memory()[:int(n)]

    @staticmethod
    def np_view(t):
        return t.numpy()

    @staticmethod
    def copy_(dst, src):
        dst.copy_(src, non_blocking=True)

    def upload(self, arr):
        return torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)

    @staticmethod
    def ptr(t):
        # an empty view has a null data_ptr: hand the library the (never empty) allocation behind it instead
        return t.data_ptr() or t.untyped_storage().data_ptr() or None

    @staticmethod
    def host(t):
        return t.cpu().numpy()

    @staticmethod
    def widen(t):
        return t.to(torch.int64)

    def run(self, stage, desc, *extra):
        fn = getattr(self.lib, "gemnet_graph_device_" + stage)
        kernels = {"count": 3, "edges": 4 if desc.triplets_only else 6, "quads": 1, "pad": 1}[stage]
        _lib.check(fn(ctypes.byref(desc), *extra, ops._stream()), "gemnet_graph_device_" + stage, kernels=kernels)


def _device_graph_build(be, R64, N, cutoff, int_cutoff, triplets_only):
    """Dynamic-shape use of the device builder (count -> allocate -> edges -> allocate -> quads) on backend `be`.
    Returns (arrays dict, header as numpy)."""
    A, M = int(R64.shape[0]), int(N.shape[0])
    a = {"R64": be.upload(R64.reshape(-1)), "n_per_mol": be.upload(N.astype(np.int32)),
         "hdr": be.upload(np.array([A, M] + [0] * (HDR_WORDS - 2), dtype=np.int32)),
         "mol_off": be.empty(M + 1), "atom_cnt": be.empty(3 * A), "upper_off": be.empty(A + 1), "in_ptr": be.empty(A + 1),
         "int_off": be.empty(A + 1), "R32": be.empty(3 * A, torch.float32), "batch_seg": be.empty(A)}
    d = DeviceGraphDesc()
    d.R, d.n_per_mol, d.hdr = be.ptr(a["R64"]), be.ptr(a["n_per_mol"]), be.ptr(a["hdr"])
    d.cutoff, d.int_cutoff, d.triplets_only = float(cutoff), float(int_cutoff), int(bool(triplets_only))
    d.cap_atoms, d.cap_molecules = A, M
    d.cap_edges = d.cap_triplets = d.cap_int_edges = d.cap_intm = d.cap_quads = _I32_MAX  # sizes not known yet
    for k in ("mol_off", "atom_cnt", "upper_off", "in_ptr", "int_off", "R32", "batch_seg"):
        setattr(d, k, be.ptr(a[k]))
    be.run("count", d)
    hdr = be.host(a["hdr"])
    if hdr[HDR_OVERFLOW]:
        raise _lib.GemnetB200Error(f"device index builder: inconsistent sizes (flags {int(hdr[HDR_OVERFLOW])}): the atoms "
                                   "per molecule do not add up to the number of coordinates, or a count exceeds 2^31-1")
    nE, nT, nI, nIm = (int(hdr[k]) for k in (HDR_EDGES, HDR_TRIPLETS, HDR_INT_EDGES, HDR_INTM))
    sizes = [(_EDGE_ARRAYS, nE), (_TRIPLET_ARRAYS, nT)]
    if not triplets_only:
        sizes += [(_INT_ARRAYS, nI), (_INTM_ARRAYS, nIm)]
        a["intm_off_a"], a["intm_off_b"], a["quad_ptr"] = be.empty(nI + 1), be.empty(nI + 1), be.empty(nE + 1)
    a["rowptr3"] = be.empty(nE + 1)
    for names, n in sizes:
        for k in names:
            a[k] = be.empty(n)
    d.cap_edges, d.cap_triplets, d.cap_int_edges, d.cap_intm = nE, nT, nI, nIm
    for k in a:
        if k not in ("R64", "n_per_mol", "hdr"):
            setattr(d, k, be.ptr(a[k]))
    be.run("edges", d)
    if not triplets_only:
        hdr = be.host(a["hdr"])
        if hdr[HDR_OVERFLOW]:
            raise _lib.GemnetB200Error("device index builder: more than 2^31-1 quadruplets in one batch")
        nQ = int(hdr[HDR_QUADS])
        for k in _QUAD_ARRAYS:
            a[k] = be.empty(nQ)
            setattr(d, k, be.ptr(a[k]))
        d.cap_quads = nQ
        be.run("quads", d)
    return a, hdr


def build_batch_device(R_list, Z_list, cutoff=5.0, int_cutoff=10.0, triplets_only=True, device="cuda",
                       index_dtype=torch.int32, _backend=None):
    """Index construction ON THE GPU: the drop-in for DataContainer.__getitem__ /
    ase_calculator.Molecule.get (data_container.py:156-408) that uploads only the coordinates (float64, as the
    reference's distance test uses them), atomic numbers and molecule sizes, and builds every index array of the
    reference's input dict with the kernels of csrc/graph_device.cu — the same arrays, element for element, as
    `build_batch` (tests/test_gpu_graph_device.py).  Returns the dict on `device`: float32 R, `index_dtype` index
    tensors (int32 = what the kernels consume; torch.int64 = the reference's dtype, data_container.py:505-512).
    Two device->host reads of the 64-byte size header (one for triplet models) are the only synchronisations."""
    if index_dtype not in (None, torch.int32, torch.int64):
        raise ValueError("index_dtype must be torch.int32 or torch.int64")
    be = _backend or _CudaBackend(device)
    N = np.array([len(z) for z in Z_list], dtype=np.int32)
    R64 = (np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, 3) for r in R_list], axis=0) if len(R_list)
           else np.zeros((0, 3)))
    a, _ = _device_graph_build(be, np.ascontiguousarray(R64), N, cutoff, int_cutoff, triplets_only)
    Z = np.concatenate([np.asarray(z) for z in Z_list]).astype(np.int64) if len(Z_list) else np.zeros(0, np.int64)
    out = {"N": be.upload(N.astype(np.int64)), "Z": be.upload(Z), "R": a["R32"].reshape(-1, 3)}
    keys = ("batch_seg",) + TRIPLET_KEYS + (() if triplets_only else QUAD_KEYS)
    for k in keys:
        out[k] = a[k] if index_dtype in (None, torch.int32) else be.widen(a[k])
    return out


def concat_molecules(R_list, Z_list):
    """(R float64 (A,3) contiguous, Z (A,), N int32 (M,)) of a list of molecules"""
    M = len(Z_list)
    if M == 0:
        return np.zeros((0, 3)), np.zeros(0, np.int64), np.zeros(0, np.int32)
    N = np.fromiter((len(z) for z in Z_list), dtype=np.int32, count=M)
    R64 = np.concatenate(R_list, axis=0) if M > 1 else np.asarray(R_list[0])
    R64 = np.ascontiguousarray(R64, dtype=np.float64).reshape(-1, 3)
    Z = np.concatenate(Z_list) if M > 1 else np.asarray(Z_list[0])
    if R64.shape[0] != Z.shape[0] or int(N.sum()) != Z.shape[0]:
        raise ValueError("R_list and Z_list disagree on the number of atoms")
    return R64, Z, N


def count_arrays(R64, N, cutoff):
    """(n_edges, n_triplets) of a concatenated batch (gemnet_graph_count_host: the per-molecule distance tests without
    any index emission) — what a static-shape launch needs to pick its capacity bucket."""
    ne, nt = ctypes.c_int64(), ctypes.c_int64()
    rc = _lib.load().gemnet_graph_count_host(R64.ctypes.data, N.ctypes.data, len(N), float(cutoff), ctypes.byref(ne),
                                             ctypes.byref(nt))
    _lib.check(rc, "gemnet_graph_count_host", kernels=0)
    return int(ne.value), int(nt.value)


def count_batch(R_list, Z_list, cutoff):
    R64, _, N = concat_molecules(R_list, Z_list)
    return count_arrays(R64, N, cutoff)


class DeviceStaticBuilder:
    """Static-shape use of the device builder for triplet models: capacity-sized device arrays, three stage calls
    (count, edges, pad) that read every size from the device header, so ONE captured CUDA graph builds the padded input
    dict of `pad_batch` for any batch that fits the capacity.  Per batch the host only writes coordinates, atomic
    numbers and molecule sizes into the pinned staging buffers (`stage`); `enqueue` issues the H2D copies and the
    kernels on the current stream (this is what gets captured)."""

    def __init__(self, cap, cutoff, device="cuda", backend=None):
        be = self.be = backend or _CudaBackend(device)
        self.cap, self.cutoff = cap, float(cutoff)
        A, E, T, M = cap.atoms, cap.edges, cap.triplets, cap.molecules
        f64 = torch.float64
        self.host = {"R64": be.host_buffer(3 * A, f64), "Z": be.host_buffer(A), "n_per_mol": be.host_buffer(M),
                     "hdr": be.host_buffer(HDR_WORDS)}
        self.host_np = {k: be.np_view(v) for k, v in self.host.items()}
        for k, v in self.host_np.items():
            # Z = 1 behind the real atoms: when a batch overflows its capacity the pad kernel does not run, and the
            # evaluation that still follows inside the captured graph must only ever see valid atomic numbers
            v[...] = 1 if k == "Z" else 0
        a = self.a = {"R64": be.empty(3 * A, f64), "Z": be.empty(A), "n_per_mol": be.empty(M), "hdr": be.empty(HDR_WORDS),
                      "mol_off": be.empty(M + 1), "atom_cnt": be.empty(3 * A), "upper_off": be.empty(A + 1),
                      "in_ptr": be.empty(A + 1), "R32": be.empty(3 * A, torch.float32), "batch_seg": be.empty(A),
                      "rowptr3": be.empty(E + 1), "N": be.empty(M + 1)}
        for k in _EDGE_ARRAYS:
            a[k] = be.empty(E)
        for k in _TRIPLET_ARRAYS:
            a[k] = be.empty(T)
        d = self.desc = DeviceGraphDesc()
        d.R, d.n_per_mol, d.hdr = be.ptr(a["R64"]), be.ptr(a["n_per_mol"]), be.ptr(a["hdr"])
        d.cutoff, d.int_cutoff, d.triplets_only = self.cutoff, 0.0, 1
        d.cap_atoms, d.cap_molecules, d.cap_edges, d.cap_triplets = A, M, E, T
        d.cap_int_edges = d.cap_intm = d.cap_quads = 0
        for k in a:
            if k not in ("R64", "n_per_mol", "hdr", "Z", "N"):
                setattr(d, k, be.ptr(a[k]))
        self.sizes = (0, 0)

    def stage(self, R_list, Z_list):
        """host side of one batch: coordinates (float64), atomic numbers, atoms per molecule, (n_atoms, n_molecules)"""
        return self.stage_arrays(*concat_molecules(R_list, Z_list))

    def stage_arrays(self, R64, Z, N):
        h = self.host_np
        A, M = int(Z.shape[0]), int(N.shape[0])
        if A + BatchCapacity.PAD_ATOMS > self.cap.atoms or M > self.cap.molecules:
            raise _lib.GemnetB200Error(f"batch (atoms={A}, molecules={M}) exceeds {self.cap}")
        h["R64"][:3 * A] = R64.reshape(-1)
        h["Z"][:A] = Z
        h["n_per_mol"][:M] = N
        h["hdr"][0], h["hdr"][1] = A, M
        self.sizes = (A, M)
        return A, M

    def enqueue(self):
        be, a, d = self.be, self.a, self.desc
        for k in ("R64", "Z", "n_per_mol", "hdr"):
            be.copy_(a[k], self.host[k])
        be.run("count", d)
        be.run("edges", d)
        be.run("pad", d, be.ptr(a["Z"]), be.ptr(a["N"]))

    def inputs(self):
        a = self.a
        out = {k: a[k] for k in SERVING_KEYS if k not in ("R",)}
        out["R"] = a["R32"].reshape(-1, 3)
        return out

    def header(self):
        return self.be.host(self.a["hdr"])
