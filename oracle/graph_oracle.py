"""TEST INFRASTRUCTURE — numpy restatement of the reference HOST index builder (row H0 of SURVEY §8a).

Follows gemnet/training/data_container.py:
  156-408  DataContainer.__getitem__      (edge list, id_swap / id_undir, sort + Kidx)
  411-425  get_triplets                   (c -> a <- b, b != c)
  428-489  get_quadruplets                (c -> a - b <- d via intermediate triplets; c != b, a != d, c != d)
  520-546  repeat_blocks                  (known-answer vectors in its docstring, :527-533)
  548-565  ragged_range                   (known-answer vector :553-556)

The reference goes through scipy.sparse CSR fancy indexing; this restatement enumerates the same
sets with plain numpy in the same order (row-major over (target, source), neighbours by ascending
source atom) and applies the same `np.argsort` (default, unstable kind) so that on the same numpy
build it reproduces the reference arrays element for element (tests/test_oracle_vs_reference.py).
Cutoff tests are made in float64 on float64 coordinates (data_container.py:258-267) while the
returned `R` is float32 (:238).
"""
import numpy as np


def repeat_blocks(sizes, repeats):
    """data_container.py:520-546.  sizes=[1,3,2], repeats=[3,2,3] -> [0 0 0 1 2 3 1 2 3 4 5 4 5 4 5]."""
    sizes = np.asarray(sizes, dtype=np.int64)
    repeats = np.asarray(repeats, dtype=np.int64)
    out = np.empty(int((sizes * repeats).sum()), dtype=np.int32)
    start = 0
    o = 0
    for size, rep in zip(sizes.tolist(), repeats.tolist()):
        block = np.arange(start, start + size, dtype=np.int32)
        for _ in range(rep):
            out[o:o + size] = block
            o += size
        start += size
    return out


def ragged_range(sizes):
    """data_container.py:548-565.  sizes=[1,3,2] -> [0 0 1 2 0 1]."""
    sizes = np.asarray(sizes, dtype=np.int64)
    out = np.empty(int(sizes.sum()), dtype=np.int32)
    o = 0
    for size in sizes.tolist():
        out[o:o + size] = np.arange(size, dtype=np.int32)
        o += size
    return out


def _directed_pairs(R_list, cutoff):
    """All (target, source) pairs, i != j, D_ij <= cutoff, block-diagonal over molecules, row-major
    (data_container.py:258-262,273-274)."""
    ts, ss = [], []
    off = 0
    for R in R_list:
        R = np.asarray(R)
        D = np.linalg.norm(R[:, None, :] - R[None, :, :], axis=-1)
        adj = D <= cutoff
        np.fill_diagonal(adj, False)
        t, s = np.nonzero(adj)  # row-major: target ascending, then source ascending
        ts.append(t + off)
        ss.append(s + off)
        off += R.shape[0]
    if not ts:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), off
    return np.concatenate(ts).astype(np.int64), np.concatenate(ss).astype(np.int64), off


def _incoming_lists(idx_t, idx_s, n_atoms):
    """For every atom x: edge ids of edges (k -> x) ordered by ascending source k — what row x of the
    reference's `edge_ids` CSR holds (data_container.py:311-315)."""
    order = np.lexsort((idx_s, idx_t))  # sort by target, then source
    counts = np.bincount(idx_t, minlength=n_atoms)
    ptr = np.zeros(n_atoms + 1, dtype=np.int64)
    ptr[1:] = np.cumsum(counts)
    return order.astype(np.int32), ptr


def build_batch(R_list, Z_list, cutoff=5.0, int_cutoff=10.0, triplets_only=True):
    """Equivalent of DataContainer.__getitem__ for an explicit list of molecules (no E/F/id keys).

    R_list: list of (n_i, 3) float64 arrays; Z_list: list of (n_i,) int arrays.
    Returns a dict of numpy arrays: N, batch_seg, Z, R(float32), id_undir, id_swap, id_c, id_a,
    id3_expand_ba, id3_reduce_ca, Kidx3 [+ the id4_* / Kidx4 arrays].  Integer arrays are int32 like the
    reference's numpy stage (it casts to int64 only in convert_to_tensor, :491-494).
    """
    data = {}
    N = np.array([len(z) for z in Z_list], dtype=np.int64)
    data["N"] = N
    data["batch_seg"] = np.repeat(np.arange(len(N), dtype=np.int32), N)
    data["Z"] = np.concatenate([np.asarray(z) for z in Z_list]).astype(np.int32)
    data["R"] = np.concatenate([np.asarray(r) for r in R_list], axis=0).astype(np.float32)

    index_keys = ["id_undir", "id_swap", "id_c", "id_a", "id3_expand_ba", "id3_reduce_ca", "Kidx3"]
    if not triplets_only:
        index_keys += ["id4_int_b", "id4_int_a", "id4_reduce_ca", "id4_expand_db", "id4_reduce_cab",
                       "id4_expand_abd", "Kidx4", "id4_reduce_intm_ca", "id4_expand_intm_db",
                       "id4_reduce_intm_ab", "id4_expand_intm_ab"]

    idx_t, idx_s, n_atoms = _directed_pairs(R_list, cutoff)
    if not triplets_only:
        idx_int_t, idx_int_s, _ = _directed_pairs(R_list, int_cutoff)

    if len(idx_t) == 0:  # data_container.py:282-285
        for k in index_keys:
            data[k] = np.array([], dtype="int32")
        return data

    # undirected ordering: [upper triangle (t < s) row-major] ++ [same pairs reversed]  (:289-293)
    mask = idx_t < idx_s
    up_t, up_s = idx_t[mask], idx_s[mask]
    idx_t = np.concatenate([up_t, up_s]).astype(np.int32)
    idx_s = np.concatenate([up_s, up_t]).astype(np.int32)
    n_undir = len(up_t)
    ind = np.arange(n_undir, dtype=np.int32)
    data["id_undir"] = np.concatenate([ind, ind]).astype(np.int32)
    data["id_c"] = idx_s
    data["id_a"] = idx_t
    id_swap = np.concatenate([ind + n_undir, ind])
    data["id_swap"] = id_swap
    if not triplets_only:
        data["id4_int_a"] = idx_int_t
        data["id4_int_b"] = idx_int_s

    inc, ptr = _incoming_lists(idx_t.astype(np.int64), idx_s.astype(np.int64), n_atoms)
    deg = (ptr[1:] - ptr[:-1])

    # ---- triplets (get_triplets, :411-425): for edge e = (s -> t) every edge (k -> s); drop k == t
    src = idx_s.astype(np.int64)
    cnt = deg[src]
    red = np.repeat(np.arange(len(idx_s), dtype=np.int64), cnt)
    starts = np.repeat(ptr[src], cnt)
    within = np.arange(len(red), dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    exp = inc[starts + within].astype(np.int32)
    red = red.astype(np.int32)
    keep = idx_t[red] != idx_s[exp]
    exp, red = exp[keep], red[keep]
    red = id_swap[red]  # (:321)
    if len(red) > 0:
        order = np.argsort(red)  # default (unstable) kind, as the reference (:326)
        red, exp = red[order], exp[order]
        _, K = np.unique(red, return_counts=True)
        data["Kidx3"] = ragged_range(K)
    else:
        data["Kidx3"] = np.array([], dtype="int32")
    data["id3_expand_ba"] = exp
    data["id3_reduce_ca"] = red
    if triplets_only:
        return data

    # ---- quadruplets (get_quadruplets, :428-489)
    it = idx_int_t.astype(np.int64)
    is_ = idx_int_s.astype(np.int64)
    nN_t = deg[it].astype(np.int32)
    nN_s = deg[is_].astype(np.int32)

    def gather_lists(atoms, counts):
        counts = counts.astype(np.int64)
        st = np.repeat(ptr[atoms], counts)
        w = np.arange(int(counts.sum()), dtype=np.int64) - np.repeat(np.cumsum(counts) - counts, counts)
        return inc[st + w].astype(np.int32)

    intm_ca = gather_lists(it, nN_t)  # edges c -> a, a = target of the interaction edge
    intm_db = gather_lists(is_, nN_s)  # edges d -> b, b = source of the interaction edge
    red_cab = repeat_blocks(nN_t, nN_s)
    red_ca = intm_ca[red_cab]
    Nrep = np.repeat(nN_t, nN_s)
    exp_abd = np.repeat(np.arange(len(intm_db)), Nrep)
    exp_db = intm_db[exp_abd]
    red_intm_ab = np.repeat(np.arange(len(it)), nN_t)
    exp_intm_ab = np.repeat(np.arange(len(it)), nN_s)

    c = idx_s[red_ca]
    a = idx_t[red_ca]
    b = idx_t[exp_db]
    d = idx_s[exp_db]
    keep = (c != b) & (a != d) & (c != d)
    red_ca, exp_db, red_cab, exp_abd = red_ca[keep], exp_db[keep], red_cab[keep], exp_abd[keep]

    if len(red_ca) > 0:
        order = np.argsort(red_ca)  # (:371)
        red_ca, exp_db, red_cab, exp_abd = red_ca[order], exp_db[order], red_cab[order], exp_abd[order]
        _, K = np.unique(red_ca, return_counts=True)
        data["Kidx4"] = ragged_range(K)
    else:
        data["Kidx4"] = np.array([], dtype="int32")
    data["id4_reduce_ca"] = red_ca
    data["id4_expand_db"] = exp_db
    data["id4_reduce_cab"] = red_cab
    data["id4_expand_abd"] = exp_abd
    data["id4_reduce_intm_ca"] = intm_ca
    data["id4_expand_intm_db"] = intm_db
    data["id4_reduce_intm_ab"] = red_intm_ab
    data["id4_expand_intm_ab"] = exp_intm_ab
    return data


def to_torch_inputs(data, device="cpu"):
    """convert_to_tensor (data_container.py:491-518): int64 indices, float32 R."""
    import torch

    out = {}
    for k, v in data.items():
        if k == "R":
            out[k] = torch.tensor(np.asarray(v), dtype=torch.float32, device=device)
        else:
            out[k] = torch.tensor(np.asarray(v), dtype=torch.int64, device=device)
    return out
