"""Host-side profile of one eval step (engine path): where the Python time goes."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from gemnet_pytorch_b200 import GemNet, build_batch  # noqa: E402
from oracle import gemnet_oracle as go  # noqa: E402

dev = torch.device("cuda")
kw = dict(bench.BASE_KW, triplets_only=True)
model = GemNet(**kw)
model.load_state_dict(go.synthetic_state_dict(kw, 1234))
model = model.to(dev).eval()
Rl, Zl = bench.synthetic_molecules(64, 1234)
inputs = {k: v.to(dev) for k, v in build_batch(Rl, Zl, 5.0, 10.0, True).items()}
for _ in range(5):
    model(inputs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    model(inputs)
t_launch = (time.perf_counter() - t0) / 20
torch.cuda.synchronize()
t_total = (time.perf_counter() - t0) / 20
print(f"host launch time per step {t_launch * 1e3:.2f} ms, wall per step {t_total * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    model(inputs)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
