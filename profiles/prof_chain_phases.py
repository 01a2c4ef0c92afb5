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
# stamps (chain_tc.cu): kernel start, prologue done, predecessor complete, then per (step s, half hf) of the software
# pipeline: [s > 0: accumulators of (s-1, hf) complete, epilogue of (s-1, hf) done], [s < n: operand written, operand
# handed to the issuer warp (mbarrier arrive)] — thread 0 = a worker thread
halves = 2 if int(os.environ.get('CHAIN_ROWS', '0')) in (0, 64) and M > 148 * 16 else 1
labels = ["kernel start", "prologue done (barriers, TMEM alloc)", "predecessor complete (griddepcontrol.wait)"]
for s_ in range(n + 1):
    for h in range(halves):
        if s_ > 0:
            labels += [f"({s_ - 1},{h}) accumulators complete", f"({s_ - 1},{h}) epilogue done"]
        if s_ < n:
            labels += [f"({s_},{h}) operand written (+ weights in TMEM)", f"({s_},{h}) handed to the issuer warp"]
t = [v for v in t if v]
print(f"M={M} steps={n}: total {t[-1] - t[0]} cycles")
for i in range(1, len(t)):
    print(f"{t[i] - t[i - 1]:8d}  {t[i] - t[0]:8d}  {labels[i] if i < len(labels) else '?'}")
