"""CUDA-event timing of gemnet_chain_tc programs of growing depth / epilogue traffic at the bench sizes
(M = 8186 edges, 762 atoms): calibrates the per-step cost of the chain kernel.  python profiles/time_chain.py [M]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from gemnet_pytorch_b200 import chain, dense_ops  # noqa: E402
from gemnet_pytorch_b200.chain import ACT, CARRY_ADD, SET_CARRY, TO_A, Program, Step, Term  # noqa: E402

dev = torch.device("cuda:0")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8186
g = torch.Generator().manual_seed(0)
R = lambda *s: torch.randn(*s, generator=g).to(dev)  # noqa: E731
Ws = [(torch.randn(128, 128, generator=g) / 11).to(dev) for _ in range(8)]
pw = [dense_ops.packed_weight(w, False) for w in Ws]
x = R(M, 128)
bufs = [torch.empty(M, 128, device=dev) for _ in range(12)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(prog, reps=20, cold=False):
    for _ in range(3):
        chain.run(prog)
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        chain.run(prog)
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def report(name, prog):
    print(json.dumps({"program": name, "M": M, "steps": len(prog.steps), "us_warm": round(timeit(prog), 2),
                      "us_cold": round(timeit(prog, cold=True), 2)}), flush=True)


report("1 step: load, gemm, out", Program(M, [Step(pw[0], terms=[Term(x)], out=bufs[0])]))
report("1 step: + act + dact", Program(M, [Step(pw[0], terms=[Term(x)], flags=ACT, out=bufs[0], dact=bufs[1])]))
report("1 step: + act + dact + post0 + mul_early",
       Program(M, [Step(pw[0], terms=[Term(x)], flags=ACT, out=bufs[0], dact=bufs[1], post0=bufs[2], mul_early=bufs[3])]))
report("1 step: no global epilogue traffic (dot only)",
       Program(M, [Step(pw[0], terms=[Term(x)], flags=ACT, dot_w=Ws[0][0].contiguous(), dot_out=bufs[0][:, :1].contiguous())]))
for n in (2, 3, 5, 7):
    steps = [Step(pw[0], terms=[Term(x)], flags=ACT | SET_CARRY | TO_A)]
    for j in range(1, n):
        steps.append(Step(pw[j], flags=ACT | CARRY_ADD | SET_CARRY | TO_A))
    steps[-1].out = bufs[0]
    report(f"{n} steps, no dact", Program(M, steps))
    steps = [Step(pw[0], terms=[Term(x)], flags=ACT | SET_CARRY | TO_A, dact=bufs[1])]
    for j in range(1, n):
        steps.append(Step(pw[j], flags=ACT | CARRY_ADD | SET_CARRY | TO_A, dact=bufs[1 + j]))
    steps[-1].out = bufs[0]
    report(f"{n} steps, dact each", Program(M, steps))
# three loads (Post forward shape)
steps = [Step(pw[0], terms=[Term(x)], flags=ACT | SET_CARRY, dact=bufs[1]),
         Step(pw[1], terms=[Term(bufs[9])], flags=ACT | CARRY_ADD | SET_CARRY, dact=bufs[2]),
         Step(pw[2], terms=[Term(bufs[10])], flags=ACT | CARRY_ADD | SET_CARRY | TO_A, dact=bufs[3])]
for j in range(3, 7):
    steps.append(Step(pw[j], flags=ACT | CARRY_ADD | SET_CARRY | TO_A, dact=bufs[1 + j]))
steps[-1].out = bufs[0]
report("Post-forward shape: 3 loads + 4 chained, dact each", Program(M, steps))
# the unfused equivalent: 7 dense_tc launches
xs = x


def unfused():
    y = xs
    for j in range(7):
        y, _ = dense_ops.dense_tc(y, pw[j], act=1, want_dact=True)
    return y


for _ in range(3):
    unfused()
ts = []
for _ in range(20):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    unfused()
    e.record()
    e.synchronize()
    ts.append(s.elapsed_time(e) * 1e3)
ts.sort()
print(json.dumps({"program": "7 dense_tc launches (eager, PDL)", "M": M, "us_warm": round(ts[10], 2)}), flush=True)
