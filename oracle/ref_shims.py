"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference from /root/reference.

The reference (TUM-DAML/gemnet_pytorch @ a0164f7) does not import on python 3.12 / numpy 2 /
torch 2.11 without `torch_scatter`.  This module installs four shims that do not touch any
hot-path arithmetic and then exposes the reference classes.  It is used by
`oracle/make_golden.py` (fixture generation) and by the CPU tests that pin `oracle/gemnet_oracle.py`
against the reference when `/root/reference` exists (it does not exist on the GPU box).

Shims:
  1. `torch_scatter.scatter` -> index_add based sum / mean     (gemnet/model/gemnet.py:10, atom_update_block.py:2)
  2. `np.bool = bool`                                           (gemnet/training/data_container.py:261,267)
  3. `np.math = math`                                           (gemnet/model/layers/basis_utils.py:102-103,166-167)
  4. `LambdaLR(verbose=...)` kwarg dropped                      (gemnet/training/schedules.py:46)
"""
import math
import os
import sys
import types

import numpy as np
import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # oracle/build_ref.sh (travels to the GPU box)
REFERENCE_ROOT = os.environ.get("GEMNET_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/gemnet/model") else _STAGED)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gemnet", "model"))


def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and out is None
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    res = res.index_add(0, index, src)
    if reduce in ("sum", "add"):
        return res
    if reduce != "mean":
        raise NotImplementedError(reduce)
    cnt = torch.zeros(n, dtype=src.dtype, device=src.device)
    cnt = cnt.index_add(0, index, torch.ones_like(index, dtype=src.dtype)).clamp(min=1)
    return res / cnt.view((-1,) + (1,) * (src.dim() - 1))


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if not hasattr(np, "bool"):
        np.bool = bool
    if not hasattr(np, "math"):
        np.math = math
    if "torch_scatter" not in sys.modules:
        mod = types.ModuleType("torch_scatter")
        mod.scatter = _scatter
        sys.modules["torch_scatter"] = mod
    import torch.optim.lr_scheduler as lrs

    if not getattr(lrs.LambdaLR, "_gemnet_shim", False):
        _orig = lrs.LambdaLR.__init__

        def _init(self, optimizer, lr_lambda, last_epoch=-1, verbose=None):
            _orig(self, optimizer, lr_lambda, last_epoch=last_epoch)

        lrs.LambdaLR.__init__ = _init
        lrs.LambdaLR._gemnet_shim = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load_reference():
    """Returns (GemNet, DataContainer) classes of the unmodified reference."""
    install()
    from gemnet.model.gemnet import GemNet  # noqa
    from gemnet.training.data_container import DataContainer  # noqa

    return GemNet, DataContainer
