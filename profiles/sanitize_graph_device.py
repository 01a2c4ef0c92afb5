"""Small driver for compute-sanitizer over the index-construction kernels (csrc/graph_device.cu) and the index checks of
prep.cu: dynamic-shape builds of a GemNet-T and a GemNet-Q batch (checked against the host builder), a static-shape
build + pad, an overflowing static build, csr_by_key / index_range_check with out-of-range ids."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gemnet_pytorch_b200 import build_batch, graph, ops  # noqa: E402
from tests import helpers as H  # noqa: E402

dev = torch.device("cuda")
for trip, ids in ((True, list(range(100, 124))), (False, list(range(30, 36)))):
    Rl, Zl = H.molecules(ids)
    host = build_batch(Rl, Zl, 5.0, 10.0, trip, as_torch=False)
    got = graph.build_batch_device(Rl, Zl, 5.0, 10.0, trip, dev)
    for k, v in host.items():
        assert np.array_equal(got[k].cpu().numpy().reshape(v.shape), v), k
Rl, Zl = H.molecules(list(range(100, 116)))
hb = build_batch(Rl, Zl, 5.0, 10.0, True)
cap = graph.BatchCapacity.for_batches([hb], margin=1.1)
want, _ = graph.pad_batch(hb, cap, check=False)
sb = graph.DeviceStaticBuilder(cap, 5.0, device=dev)
sb.stage(Rl, Zl)
sb.enqueue()
assert sb.header()[graph.HDR_OVERFLOW] == 0
for k in graph.SERVING_KEYS:
    assert np.array_equal(sb.inputs()[k].cpu().numpy().reshape(want[k].shape), want[k].numpy()), k
small = graph.DeviceStaticBuilder(graph.BatchCapacity(cap.atoms, 64, 64, cap.molecules), 5.0, device=dev)
small.stage(Rl, Zl)
small.enqueue()
assert small.header()[graph.HDR_OVERFLOW] != 0
flag = torch.zeros(1, dtype=torch.int32, device=dev)
ops.csr_by_key(torch.tensor([2, -1, 0, 9, 2], dtype=torch.int32, device=dev), 3, flag)
ops.index_range_check([(torch.arange(5, dtype=torch.int32, device=dev), 4)], flag)
assert flag.item() == 1
torch.cuda.synchronize()
print("sanitize_graph_device: all checks passed")
