"""gemnet_pytorch_b200 — B200-native (sm_100a) implementation of GemNet's interaction-block hot path behind the
reference's `GemNet` interface.  See DESIGN.md / INTEGRATION.md."""
from .model import GemNet  # noqa: F401
from .graph import build_batch, Molecule, GraphPlan  # noqa: F401

__all__ = ["GemNet", "build_batch", "Molecule", "GraphPlan"]
