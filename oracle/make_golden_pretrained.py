"""TEST INFRASTRUCTURE — tests/golden/model_{T,Q}_pretrained.npz: energies and forces of the UNMODIFIED reference with
its SHIPPED checkpoints (pretrained/GemNet-{T,Q}/model.pth) on COLL test molecules, float32 and float64 (the float64 run
on the float32-rounded coordinates).  Only the reference's OUTPUTS are committed (a few KB); the checkpoints themselves
reach the GPU box with the staged reference (oracle/build_ref.sh -> oracle/_ref/pretrained), where
tests/test_gpu_model.py::test_pretrained_checkpoint_parity loads them into gemnet_pytorch_b200.GemNet.

Run where /root/reference exists:   python -m oracle.make_golden_pretrained
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

from . import ref_shims
from .make_golden import MODEL_BATCH, MODEL_KW, OUT, base_kwargs

warnings.filterwarnings("ignore")


def main():
    GemNet, DataContainer = ref_shims.load_reference()
    REF = ref_shims.REFERENCE_ROOT
    test_npz = os.path.join(REF, "data", "coll_v1.2_test.npz")
    for name in ("T", "Q"):
        kw = base_kwargs()
        kw.update(MODEL_KW[name])
        ids = MODEL_BATCH[name]
        dc = DataContainer(test_npz, cutoff=kw["cutoff"], int_cutoff=kw["int_cutoff"], triplets_only=kw["triplets_only"])
        inputs = {k: v for k, v in dc[ids].items() if k not in ("E", "F")}
        model = GemNet(**kw)
        print(name, model.load_state_dict(torch.load(os.path.join(REF, "pretrained", f"GemNet-{name}", "model.pth"),
                                                     map_location="cpu")), flush=True)
        E32, F32 = model.predict(dict(inputs))
        model.double()
        i64 = dict(inputs)
        i64["R"] = inputs["R"].double()
        E64, F64 = model.predict(i64)
        kwj = {k: v for k, v in kw.items() if k != "scale_file"}
        np.savez_compressed(os.path.join(OUT, f"model_{name}_pretrained.npz"), mol_ids=np.array(ids),
                            kwargs_json=np.array(json.dumps(kwj)), E32=E32.numpy(), F32=F32.numpy(), E64=E64.numpy(),
                            F64=F64.numpy())
        print(f"  model_{name}_pretrained.npz written ({len(ids)} molecules)")


if __name__ == "__main__":
    sys.exit(main())
