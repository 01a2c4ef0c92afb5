"""clock64() stamps of the chain kernel's phases (thread 0 of CTA 0; gemnet_chain_set_profile_buffer) for an n-step
program at M rows.  python profiles/prof_chain_phases.py [M] [steps]"""
import sys

import torch

sys.path.insert(0, ".")
from gemnet_pytorch_b200 import _lib, chain, dense_ops  # noqa: E402
from gemnet_pytorch_b200.chain import ACT, CARRY_ADD, SET_CARRY, TO_A, Program, Step, Term  # noqa: E402

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8186
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator().manual_seed(0)
Ws = [(torch.randn(128, 128, generator=g) / 11).to(dev) for _ in range(n)]
pw = [dense_ops.packed_weight(w, False) for w in Ws]
x = torch.randn(M, 128, generator=g).to(dev)
bufs = [torch.empty(M, 128, device=dev) for _ in range(n + 1)]
steps = [Step(pw[0], terms=[Term(x)], flags=ACT | SET_CARRY | TO_A, dact=bufs[1])]
for j in range(1, n):
    steps.append(Step(pw[j], flags=ACT | CARRY_ADD | SET_CARRY | TO_A, dact=bufs[1 + j]))
steps[-1].out = bufs[0]
prog = Program(M, steps)
import os
_lib.check(_lib.load().gemnet_chain_set_tile_rows(int(os.environ.get('CHAIN_ROWS', '0'))), 'rows')
for _ in range(3):
    chain.run(prog)
buf = torch.zeros(256, dtype=torch.int64, device=dev)
_lib.load().gemnet_chain_set_profile_buffer(buf.data_ptr())
chain.run(prog)
torch.cuda.synchronize()
_lib.load().gemnet_chain_set_profile_buffer(None)
t = buf.cpu().tolist()
names = ["kernel start", "prologue done (barriers, TMEM alloc)", "predecessor complete (griddepcontrol.wait)"]
per_step = ["operand written (this thread)", "operand visible (__syncthreads)",
            "MMAs issued", "accumulators complete", "epilogue done (this thread)"]
labels = names + [f"step {s}: {p}" for s in range(n) for p in per_step]
t = [v for v in t if v]
print(f"M={M} steps={n}: total {t[-1] - t[0]} cycles")
for i in range(1, len(t)):
    print(f"{t[i] - t[i - 1]:8d}  {t[i] - t[0]:8d}  {labels[i] if i < len(labels) else '?'}")
