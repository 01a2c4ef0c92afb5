"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference ships no tests or golden vectors, so these fixtures are the pins:

  coll_subset.npz        first 256 molecules of data/coll_v1.2_test.npz (N, Z, R float64) — inputs only
  graph_{T,Q}.npz        DataContainer.__getitem__ index arrays (data_container.py:156-408) for small batches
  ops.npz                per-layer outputs of the reference layers in float64 (+ the reference's own float32
                         result where its float32 noise matters): geometry (gemnet.py:262-451), BesselBasisLayer,
                         SphericalBasisLayer, TensorBasisLayer (basis_layers.py), EfficientInteractionDownProjection
                         and EfficientInteractionBilinear (efficient.py:41-57,159-189), AtomUpdateBlock scatter
  model_{T,Q,dT,dQ}.npz  GemNet.forward (gemnet.py:453-615) energies and forces, float32 and float64, with
                         seeded synthetic weights (oracle.gemnet_oracle.synthetic_state_dict, seed 1234) that
                         are re-creatable anywhere, so no checkpoint has to travel
  _local/…               (git-ignored) the same with the shipped pretrained checkpoints + the checkpoints
                         themselves as flat npz, for boxes that receive the working tree (gpurun)
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

from . import gemnet_oracle as go
from . import ref_shims

warnings.filterwarnings("ignore")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF = ref_shims.REFERENCE_ROOT
N_SUBSET = 256
MODEL_KW = {
    "T": dict(triplets_only=True, direct_forces=False),
    "Q": dict(triplets_only=False, direct_forces=False),
    "dT": dict(triplets_only=True, direct_forces=True),
    "dQ": dict(triplets_only=False, direct_forces=True, forces_coupled=True),
}
MODEL_BATCH = {"T": list(range(64)), "Q": list(range(16)), "dT": list(range(64, 96)), "dQ": list(range(16, 28))}


def base_kwargs():
    kw = json.load(open(os.path.join(REF, "pretrained", "GemNet-T", "model_kwargs.json")))
    kw["scale_file"] = os.path.join(REF, "pretrained", "scaling_factors.json")
    return kw


def np_inputs(batch):
    return {k: v.numpy() for k, v in batch.items() if k not in ("E", "F")}


def main():
    GemNet, DataContainer = ref_shims.load_reference()
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(os.path.join(OUT, "_local"), exist_ok=True)
    test_npz = os.path.join(REF, "data", "coll_v1.2_test.npz")

    raw = np.load(test_npz)
    n_atoms = int(raw["N"][:N_SUBSET].sum())
    np.savez_compressed(os.path.join(OUT, "coll_subset.npz"), N=raw["N"][:N_SUBSET].astype(np.int16),
                        Z=raw["Z"][:n_atoms].astype(np.int8), R=raw["R"][:n_atoms].astype(np.float64))

    dcs = {t: DataContainer(test_npz, cutoff=5.0, int_cutoff=10.0, triplets_only=t) for t in (True, False)}

    # ---- graph fixtures
    for name, trip_only, ids in (("T", True, list(range(8)) + [200, 201]), ("Q", False, [0, 1, 2, 3, 200])):
        b = np_inputs(dcs[trip_only][ids])
        out = {"mol_ids": np.array(ids)}
        for k, v in b.items():
            if k in ("R", "Z", "N"):
                continue
            out[k] = v.astype(np.int32)
        np.savez_compressed(os.path.join(OUT, f"graph_{name}.npz"), **out)

    # ---- per-op fixtures (float64 truth from the reference layers, float32 reference result alongside)
    from gemnet.model.layers.basis_layers import BesselBasisLayer, SphericalBasisLayer, TensorBasisLayer
    from gemnet.model.layers.efficient import EfficientInteractionBilinear, EfficientInteractionDownProjection

    ops = {}
    ids = [1, 4, 7]  # 8 + 6 + 4 atoms: a few thousand quadruplets
    bq = dcs[False][ids]
    R64 = bq["R"].double()  # float32-rounded coordinates, evaluated in float64
    I = {k: v for k, v in bq.items()}
    ops["mol_ids"] = np.array(ids)
    D_ca, V_ca = GemNet.calculate_interatomic_vectors(R64, I["id_c"], I["id_a"])
    D_ab, _ = GemNet.calculate_interatomic_vectors(R64, I["id4_int_b"], I["id4_int_a"])
    A3 = GemNet.calculate_angles3(R64, I["id_c"], I["id_a"], I["id3_reduce_ca"], I["id3_expand_ba"])
    Phi_cab, Phi_abd, Theta = GemNet.calculate_angles(
        R64, I["id_c"], I["id_a"], I["id4_int_b"], I["id4_int_a"], I["id4_expand_abd"], I["id4_reduce_cab"],
        I["id4_expand_intm_db"], I["id4_reduce_intm_ca"], I["id4_expand_intm_ab"], I["id4_reduce_intm_ab"])
    ops.update(D_ca=D_ca.numpy(), V_ca=V_ca.numpy(), D_ab=D_ab.numpy(), Angles3=A3.numpy(),
               Phi_cab=Phi_cab.numpy(), Phi_abd=Phi_abd.numpy(), Theta_cabd=Theta.numpy())

    torch.manual_seed(0)
    rbf_layer = BesselBasisLayer(6, cutoff=5.0)
    with torch.no_grad():
        rbf_layer.frequencies.mul_(1 + 0.03 * torch.randn(6))
    ops["frequencies"] = rbf_layer.frequencies.detach().numpy().astype(np.float32)
    ops["rbf32"] = rbf_layer(D_ca.float()).detach().numpy()
    ops["rbf64"] = rbf_layer.double()(D_ca).detach().numpy()

    cbf3 = SphericalBasisLayer(7, 6, cutoff=5.0, efficient=True)
    nE = D_ca.shape[0]
    rbf_env64, sph64 = cbf3(D_ca, A3, I["id3_reduce_ca"], I["Kidx3"])
    rbf_env32, sph32 = cbf3(D_ca.float(), A3.float(), I["id3_reduce_ca"], I["Kidx3"])
    ops["cbf3_rbf_env64"] = rbf_env64.numpy()  # (7,E,6)
    ops["cbf3_rbf_env32"] = rbf_env32.numpy()
    ops["cbf3_sph64"] = sph64[I["id3_reduce_ca"], I["Kidx3"]].numpy()  # (T,7) un-padded
    ops["cbf3_sph32"] = sph32[I["id3_reduce_ca"], I["Kidx3"]].numpy()

    cbf4 = SphericalBasisLayer(7, 6, cutoff=10.0, efficient=False)
    ops["cbf4_64"] = cbf4(D_ab, Phi_abd, I["id4_expand_intm_ab"], None).numpy()  # (M,42)
    ops["cbf4_32"] = cbf4(D_ab.float(), Phi_abd.float(), I["id4_expand_intm_ab"], None).numpy()

    sbf = TensorBasisLayer(7, 6, cutoff=5.0, efficient=True)
    rad49_64, sphq64 = sbf(D_ca, Phi_cab, Theta, I["id4_reduce_ca"], I["Kidx4"])
    rad49_32, sphq32 = sbf(D_ca.float(), Phi_cab.float(), Theta.float(), I["id4_reduce_ca"], I["Kidx4"])
    ops["sbf_rbf_env64"] = rad49_64.numpy()  # (49,E,6)
    ops["sbf_rbf_env32"] = rad49_32.numpy()
    ops["sbf_sph64"] = sphq64[I["id4_reduce_ca"], I["Kidx4"]].numpy()  # (Q,49)
    ops["sbf_sph32"] = sphq32[I["id4_reduce_ca"], I["Kidx4"]].numpy()

    # efficient bilinear, triplet sizes (C=64,I=16,O=64,S=7) and quadruplet sizes (C=32,I=32,O=32,S=49)
    g = torch.Generator().manual_seed(7)
    for tag, S_, C_, I_, O_, idr, kidx, rbf_env, sph in (
            ("trip", 7, 64, 16, 64, I["id3_reduce_ca"], I["Kidx3"], rbf_env64, sph64),
            ("quad", 49, 32, 32, 32, I["id4_reduce_ca"], I["Kidx4"], rad49_64, sphq64)):
        down = EfficientInteractionDownProjection(S_, 6, I_).double()
        bil = EfficientInteractionBilinear(C_, I_, O_).double()
        with torch.no_grad():
            down.weight.copy_(torch.randn(down.weight.shape, generator=g, dtype=torch.float64) / 6 ** 0.5)
            bil.weight.copy_(torch.randn(bil.weight.shape, generator=g, dtype=torch.float64) / (C_ * I_) ** 0.5)
        m = torch.randn(idr.shape[0], C_, generator=g, dtype=torch.float32).double()  # float32-representable
        basis = down((rbf_env, sph))
        out = bil(basis, m, idr, kidx)
        ops[f"{tag}_Wdown"] = down.weight.detach().numpy()
        ops[f"{tag}_Wbil"] = bil.weight.detach().numpy()
        ops[f"{tag}_m"] = m.numpy().astype(np.float32)
        ops[f"{tag}_rbf_W1"] = basis[0].detach().numpy()  # (E,I,S)
        ops[f"{tag}_out64"] = out.detach().numpy()
        out32 = bil.float()(tuple(t.float() for t in down.float()((rbf_env.float(), sph.float()))), m.float(), idr, kidx)
        ops[f"{tag}_out32"] = out32.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **ops)

    # ---- whole-model fixtures
    for name, over in MODEL_KW.items():
        kw = base_kwargs()
        kw.update(over)
        ids = MODEL_BATCH[name]
        inputs = {k: v for k, v in dcs[kw["triplets_only"]][ids].items() if k not in ("E", "F")}
        res = {"mol_ids": np.array(ids), "seed": np.array(1234)}
        weight_sets = {"syn": go.synthetic_state_dict(kw, seed=1234)}
        ck = os.path.join(REF, "pretrained", f"GemNet-{name}", "model.pth")
        if os.path.exists(ck):
            weight_sets["pre"] = torch.load(ck, map_location="cpu")
        local = {}
        for wname, sd in weight_sets.items():
            model = GemNet(**kw)
            print(name, wname, model.load_state_dict(sd), flush=True)
            E32, F32 = model.predict(dict(inputs))
            model.double()
            i64 = dict(inputs)
            i64["R"] = inputs["R"].double()  # float32-rounded coordinates, evaluated in float64
            E64, F64 = model.predict(i64)
            tgt = res if wname == "syn" else local
            tgt.update({f"E32": E32.numpy(), f"F32": F32.numpy(), f"E64": E64.numpy(), f"F64": F64.numpy()})
        kwj = {k: v for k, v in kw.items() if k != "scale_file"}
        res["kwargs_json"] = np.array(json.dumps(kwj))
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **res)
        if local:
            local["mol_ids"] = np.array(ids)
            local["kwargs_json"] = np.array(json.dumps(kwj))
            np.savez_compressed(os.path.join(OUT, "_local", f"model_{name}_pretrained.npz"), **local)
            uniq = {k: v.numpy() for k, v in weight_sets["pre"].items()
                    if ".linear.weight" not in k and ".seq_energy." not in k}
            np.savez(os.path.join(OUT, "_local", f"weights_{name}.npz"), **uniq)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        p = os.path.join(OUT, f)
        if os.path.isfile(p):
            print(f"  {f}: {os.path.getsize(p) / 1024:.1f} KiB")


if __name__ == "__main__":
    sys.exit(main())
