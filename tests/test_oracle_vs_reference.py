"""CPU, build container only: pins oracle/ against the LIVE unmodified reference with the shipped pretrained
checkpoints (skipped where /root/reference does not exist, e.g. on the GPU box)."""
import json
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import gemnet_oracle as go
from oracle import graph_oracle as gr
from oracle import ref_shims

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shims.reference_available(), reason="needs /root/reference")]
REF = ref_shims.REFERENCE_ROOT


@pytest.mark.parametrize("M,n_mol", [("T", 8), ("Q", 3)])
def test_oracle_matches_live_reference_pretrained(M, n_mol):
    warnings.filterwarnings("ignore")
    GemNet, DataContainer = ref_shims.load_reference()
    kw = json.load(open(os.path.join(REF, "pretrained", f"GemNet-{M}", "model_kwargs.json")))
    kw["scale_file"] = os.path.join(REF, "pretrained", kw["scale_file"])
    sd = torch.load(os.path.join(REF, "pretrained", f"GemNet-{M}", "model.pth"), map_location="cpu")
    model = GemNet(**kw)
    model.load_state_dict(sd)
    dc = DataContainer(os.path.join(REF, "data", "coll_v1.2_test.npz"), cutoff=kw["cutoff"],
                       int_cutoff=kw["int_cutoff"], triplets_only=kw["triplets_only"])
    ids = list(range(300, 300 + n_mol))
    batch = dc[ids]
    inputs = {k: v for k, v in batch.items() if k not in ("E", "F")}
    # index builder: element-wise identical (same numpy argsort)
    raw = np.load(os.path.join(REF, "data", "coll_v1.2_test.npz"))
    off = np.concatenate([[0], np.cumsum(raw["N"])])
    mine = gr.build_batch([raw["R"][off[i]:off[i + 1]] for i in ids], [raw["Z"][off[i]:off[i + 1]] for i in ids],
                          kw["cutoff"], kw["int_cutoff"], kw["triplets_only"])
    for k, v in mine.items():
        assert np.array_equal(inputs[k].numpy(), np.asarray(v).astype(inputs[k].numpy().dtype)), k
    # float64: algebraic identity
    model.double()
    i64 = dict(inputs)
    i64["R"] = inputs["R"].double()
    E_ref, F_ref = model.predict(i64)
    E, F = go.OracleGemNet(kw, sd, dtype=torch.float64).predict(i64)
    assert (E - E_ref).abs().max() < 1e-10 * E_ref.abs().max()
    assert (F - F_ref).abs().max() < 1e-10 * F_ref.abs().max()
    # float32: within the north_star tolerance with margin
    model.float()
    E_ref32, F_ref32 = model.predict(dict(inputs))
    E32, F32 = go.OracleGemNet(kw, sd).predict(dict(inputs))
    assert (E32 - E_ref32).abs().max() < 1e-5 * E_ref32.abs().max()
    assert (F32 - F_ref32).abs().max() < 1e-5 * F_ref32.abs().max()
