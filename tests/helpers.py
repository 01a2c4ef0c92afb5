"""Shared test helpers: fixture loading and batch construction (uses oracle/ as the checker only)."""
import json
import os

import numpy as np
import torch

from oracle import graph_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def coll_subset():
    d = load_npz("coll_subset.npz")
    N = d["N"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(N)])
    return N, off, d["Z"].astype(np.int64), d["R"]


def molecules(ids):
    N, off, Z, R = coll_subset()
    return [R[off[i]:off[i + 1]] for i in ids], [Z[off[i]:off[i + 1]] for i in ids]


def oracle_batch(ids, triplets_only, cutoff=5.0, int_cutoff=10.0):
    Rl, Zl = molecules(ids)
    return graph_oracle.build_batch(Rl, Zl, cutoff, int_cutoff, triplets_only)


def model_kwargs(fixture):
    kw = json.loads(str(fixture["kwargs_json"]))
    kw["scale_file"] = None
    return kw


def canonical_segments(reduce_ids, *others):
    """Sort by (reduce, others...) lexicographically — the canonical within-segment order used to compare
    index builders whose within-segment order is unspecified (np.argsort default kind is unstable,
    data_container.py:326,371)."""
    keys = tuple(reversed((reduce_ids,) + others))
    order = np.lexsort(keys)
    return tuple(a[order] for a in (reduce_ids,) + others)
