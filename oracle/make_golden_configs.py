"""TEST INFRASTRUCTURE — round-2 golden fixtures for the configurations BASELINE.json names beyond configs[1]
(Q parity only at 16 molecules, nothing at 8 A).  Imports the UNMODIFIED reference through
oracle/ref_shims.py (needs /root/reference or the staged oracle/_ref) and writes small npz files under tests/golden/:

  model_Q_b64.npz         configs[2]: GemNet-Q, COLL test molecules 0..63 (batch 64), cutoff 5 / int_cutoff 10
  model_Q_c8.npz          configs[4]: GemNet-Q, cutoff 8 A, int_cutoff 10, 12 COLL molecules
  model_T_c8_clusters.npz configs[4]: GemNet-T, cutoff 8 A on synthetic 18-atom clusters (4 896 triplets each;
                          SURVEY 8d input 3: atoms uniform in a 3.5 A ball, min distance 0.9 A, seed 1234)
Each holds mol_ids / the cluster coordinates, the model kwargs, the seed of the synthetic weights and the reference's
float64 (truth) and float32 outputs.   python oracle/make_golden_configs.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gemnet_oracle as go  # noqa: E402
from oracle import graph_oracle as gr  # noqa: E402
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def clusters(n_clusters, n_atoms=18, radius=3.5, dmin=0.9, seed=1234):
    rng = np.random.default_rng(seed)
    Rl, Zl = [], []
    for _ in range(n_clusters):
        pts = []
        while len(pts) < n_atoms:
            p = rng.normal(size=3)
            p = p / np.linalg.norm(p) * radius * rng.uniform() ** (1 / 3)
            if all(np.linalg.norm(p - q) >= dmin for q in pts):
                pts.append(p)
        Rl.append(np.array(pts))
        Zl.append(rng.choice([1, 6, 8], size=n_atoms, p=[0.543, 0.330, 0.127]))
    return Rl, Zl


def run(name, kw, Rl, Zl, extra):
    GemNet, _ = ref_shims.load_reference()
    kw_ref = dict(kw, scale_file=os.path.join(ref_shims.REFERENCE_ROOT, "pretrained", "scaling_factors.json"))
    sd = go.synthetic_state_dict(kw, seed=1234)
    inputs = gr.to_torch_inputs(gr.build_batch(Rl, Zl, kw["cutoff"], kw["int_cutoff"], kw["triplets_only"]))
    model = GemNet(**kw_ref)
    model.load_state_dict(sd, strict=True)
    E32, F32 = model.predict(dict(inputs))
    model.double()
    i64 = dict(inputs)
    i64["R"] = inputs["R"].double()
    E64, F64 = model.predict(i64)
    res = dict(extra, seed=np.array(1234), kwargs_json=np.array(json.dumps({k: v for k, v in kw.items() if k != "scale_file"})),
               E32=E32.numpy(), F32=F32.numpy(), E64=E64.numpy(), F64=F64.numpy())
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **res)
    print(name, "E", E64.shape, "F", F64.shape, "ref32 vs ref64 dF",
          float(np.abs(F32.numpy() - F64.numpy()).max() / np.abs(F64.numpy()).max()), flush=True)


def main():
    sub = dict(np.load(os.path.join(OUT, "coll_subset.npz")))
    N = sub["N"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(N)])
    mols = lambda ids: ([sub["R"][off[i]:off[i + 1]] for i in ids], [sub["Z"][off[i]:off[i + 1]].astype(np.int64) for i in ids])  # noqa: E731
    kwQ = json.loads(str(np.load(os.path.join(OUT, "model_Q.npz"))["kwargs_json"]))
    kwT = json.loads(str(np.load(os.path.join(OUT, "model_T.npz"))["kwargs_json"]))
    which = sys.argv[1:] or ["T_c8", "Q_c8", "Q_b64"]
    if "T_c8" in which:
        Rl, Zl = clusters(4)
        run("model_T_c8_clusters", dict(kwT, cutoff=8.0), Rl, Zl,
            {"R": np.concatenate(Rl), "Z": np.concatenate(Zl), "N": np.array([18] * 4)})
    if "Q_c8" in which:
        ids = list(range(100, 112))
        run("model_Q_c8", dict(kwQ, cutoff=8.0), *mols(ids), {"mol_ids": np.array(ids)})
    if "Q_b64" in which:
        ids = list(range(64))
        run("model_Q_b64", kwQ, *mols(ids), {"mol_ids": np.array(ids)})


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
