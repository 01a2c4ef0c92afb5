"""Seeded synthetic weights for benchmarks and smoke runs (there is no network for checkpoints on a GPU box).

`seed_synthetic_weights_(model, seed)` fills a GemNet in place from a CPU generator: variance 1/fan_in like the
reference's He-orthogonal init (initializers.py:27-40), but with scale factors in [0.4, 0.8] and 0.1x read-out weights so
that energies are O(1) eV, forces O(10) eV/A and float32 rounding noise stays far below the 1e-4 parity gate (a
Gaussian-weight network with unit scale factors amplifies it through the blocks); Bessel frequencies perturbed off
n*pi like the trained checkpoints.  The result is an ordinary reference-format state_dict (`model.state_dict()`)."""
import math

import torch


@torch.no_grad()
def seed_synthetic_weights_(model, seed=1234):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    for name, p in sorted(model.named_parameters(), key=lambda kv: kv[0]):  # shared tensors appear once
        shape = tuple(p.shape)
        if name == "rbf_basis.frequencies":
            n = shape[0]
            v = math.pi * torch.arange(1, n + 1, dtype=torch.float64) * (
                1 + 0.05 * (torch.rand(n, generator=g, dtype=torch.float64) - 0.5))
        elif name.endswith("scale_factor"):
            v = 0.4 + 0.4 * torch.rand(shape, generator=g, dtype=torch.float64)
        elif name.startswith("atom_emb."):
            v = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * math.sqrt(3)
        else:
            fan_in = math.prod(shape[:-1]) if len(shape) == 3 else (shape[1] if len(shape) == 2 else 1)
            v = torch.randn(shape, generator=g, dtype=torch.float64) / math.sqrt(max(fan_in, 1))
            if ".out_energy." in name or ".out_forces." in name:
                v = v * 0.1
        p.copy_(v.to(p.dtype))
    return model
