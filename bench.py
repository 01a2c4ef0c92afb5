#!/usr/bin/env python
"""bench.py — molecules/sec of the GemNet energy+force hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--model T|Q] [--batch B] [--impl b200|reference]

N=1 workload = BASELINE.json configs[1]: GemNet-T inference (energy + forces by -dE/dR, as the shipped checkpoints),
COLL-shaped synthetic batch of 64 molecules per GPU, emb_size 128, fp32.  One "step" = one energy+force evaluation
of one batch.  N>1 (torchrun): whole molecules sharded per rank, weak scaling, no data-path collective
(inference needs none); time = max over ranks.

JSON line keys follow the driver contract; in addition
  roofline      the triplet segmented-bilinear kernel (row H7) at the bench batch: algorithmic bytes (SURVEY §8d
                formula) / CUDA-event time of isolated launches vs MEASURED_PEAKS.json hbm_gbs
  roofline_large same kernel on a 2048-molecule batch (the batch-64 launch is latency-bound)
  roofline_chain / roofline_h7_bwd  the chain-fused dense stack kernel and the tcgen05 backward of H7
  cpu_baseline  the reference's own forward on the host cores: the UNMODIFIED reference (oracle/_ref, staged by
                oracle/build_ref.sh, kind "reference") when present, else the oracle port (kind "port")
  e2e           host molecules -> (E, F) on the host through GemNet.make_molecule_server, index construction INCLUDED
`--impl reference` times that CPU path on rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

BASE_KW = dict(num_spherical=7, num_radial=6, num_blocks=4, emb_size_atom=128, emb_size_edge=128, emb_size_trip=64,
               emb_size_quad=32, emb_size_rbf=16, emb_size_cbf=16, emb_size_sbf=32, emb_size_bil_trip=64,
               emb_size_bil_quad=32, num_before_skip=1, num_after_skip=1, num_concat=1, num_atom=2, num_targets=1,
               direct_forces=False, cutoff=5.0, int_cutoff=10.0, envelope_exponent=5, extensive=True,
               forces_coupled=False, output_init="HeOrthogonal", activation="swish", scale_file=None)


MOLECULE_KIND = ["coll"]  # "coll" (configs[1-3]) or "clusters" (configs[4]); set by --molecules


def cluster_molecules(n_mol, seed, n_atoms=18, radius=3.5, dmin=0.9):
    """BASELINE configs[4] input (SURVEY §8d input 3): 18-atom clusters, atoms uniform in a 3.5 A ball with rejection-
    sampled minimum distance 0.9 A — at cutoff 8 A every pair is an edge: 306 edges, 4 896 triplets per molecule."""
    rng = np.random.default_rng(seed)
    Rl, Zl = [], []
    for _ in range(n_mol):
        pts = []
        while len(pts) < n_atoms:
            p = rng.uniform(-radius, radius, size=3)
            if p @ p <= radius * radius and all(np.linalg.norm(p - q) >= dmin for q in pts):
                pts.append(p)
        Rl.append(np.array(pts))
        Zl.append(rng.choice([1, 6, 8], size=n_atoms, p=[0.543, 0.330, 0.127]).astype(np.int64))
    return Rl, Zl


def synthetic_molecules(n_mol, seed):
    """COLL-shaped synthetic molecules: resample (with replacement) from the 256 committed COLL test molecules and
    jitter the coordinates by N(0, 0.05 A) (SURVEY §8d input 2).  With --molecules clusters: cluster_molecules."""
    if MOLECULE_KIND[0] == "clusters":
        return cluster_molecules(n_mol, seed)
    sub = dict(np.load(os.path.join(GOLDEN, "coll_subset.npz")))
    N = sub["N"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(N)])
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, len(N), size=n_mol)
    Rl = [sub["R"][off[i]:off[i + 1]] + rng.normal(0.0, 0.05, size=(N[i], 3)) for i in ids]
    Zl = [sub["Z"][off[i]:off[i + 1]].astype(np.int64) for i in ids]
    return Rl, Zl


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for j, n in enumerate(names):
                if any(len(r) > 2 + j and r[2 + j].strip().lower() == "active" for r in rows):
                    out["reasons"].append(n)
            os.unlink(self.path)
        except Exception:
            pass
        return out


def bilinear_bytes(E, U, S, C, I, O, b=4):
    """each operand once, output once, padding excluded, int32 CSR."""
    return b * (E * C + U * S + E * I * S + C * I * O + E * O) + 4 * (U + E + 1)


def bilinear_flops(E, U, S, C, I, O):
    return 2 * U * S * C + 2 * E * I * S * C + 2 * E * C * I * O


def time_bilinear(plan, dev, flush, reps, algo=0):
    """CUDA-event time of isolated launches of the triplet segmented-bilinear kernel on this plan's graph."""
    from gemnet_pytorch_b200 import ops
    E, T = plan.n_edges, plan.n_triplets
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(E, 64, generator=g).to(dev)
    sph = torch.randn(T, 7, generator=g).to(dev)
    rw = torch.randn(E, 16, 7, generator=g).to(dev)
    W = (torch.randn(64, 16, 64, generator=g) / 32).to(dev)
    with torch.no_grad():
        for _ in range(3):
            ops.segmented_bilinear(x, sph, rw, W, plan.id3_expand, plan.rowptr3, 1.0, algo)
        times = []
        for _ in range(reps):
            flush()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.segmented_bilinear(x, sph, rw, W, plan.id3_expand, plan.rowptr3, 1.0, algo)
            e.record()
            e.synchronize()
            times.append(s.elapsed_time(e) * 1e-3)
    t = float(np.mean(times))
    return t, bilinear_bytes(E, T, 7, 64, 16, 64), bilinear_flops(E, T, 7, 64, 16, 64)


def time_dense(M, dev, flush, reps):
    """CUDA-event time of isolated launches of the fused dense layer (M x 128) -> 128 with ScaledSiLU and the
    derivative output (the shape of most layers of the step); algorithmic bytes = x + y + dact + packed weight image."""
    from gemnet_pytorch_b200 import dense_ops
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(M, 128, generator=g).to(dev)
    W = (torch.randn(128, 128, generator=g) / 11).to(dev)
    pk = dense_ops.packed_weight(W, False)
    for _ in range(3):
        dense_ops.dense_tc(x, pk, act=1, want_dact=True)
    times = []
    for _ in range(reps):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        dense_ops.dense_tc(x, pk, act=1, want_dact=True)
        e.record()
        e.synchronize()
        times.append(s.elapsed_time(e) * 1e-3)
    return float(np.mean(times)), 4 * M * 128 * 3 + 6 * 128 * 128, 2 * M * 128 * 128


def time_h7_backward(plan, dev, flush, reps):
    """CUDA-event time of the tcgen05 backward of the triplet bilinear layer (segsum -> dB GEMM on tensor cores + dA /
    g_rbf_W1 -> unit kernel); algorithmic bytes: every operand and every gradient once."""
    from gemnet_pytorch_b200 import ops
    E, T = plan.n_edges, plan.n_triplets
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(E, 64, generator=g).to(dev).requires_grad_(True)
    sph = torch.randn(T, 7, generator=g).to(dev).requires_grad_(True)
    rw = torch.randn(E, 16, 7, generator=g).to(dev).requires_grad_(True)
    W = (torch.randn(64, 16, 64, generator=g) / 32).to(dev)
    go_ = torch.randn(E, 64, generator=g).to(dev)
    times = []
    for it in range(reps + 3):
        out = ops.segmented_bilinear(x, sph, rw, W, plan.id3_expand, plan.rowptr3, 1.0, 1)
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.autograd.grad(out, [x, sph, rw], go_)
        e.record()
        e.synchronize()
        if it >= 3:
            times.append(s.elapsed_time(e) * 1e-3)
    bts = 4 * (E * 64 + E * 64 + T * 7 + E * 112 + 64 * 16 * 64) + 4 * (E * 64 + T * 7 + E * 112) + 4 * (T + E + 1)
    return float(np.mean(times)), bts


def time_chain(M, dev, flush, reps):
    """CUDA-event time of one chain-fused launch in the shape of fused.Post's forward (3 operand fills, 7 GEMM steps of
    128 -> 128 with ScaledSiLU, every step storing its activation derivative); algorithmic bytes: the three inputs, the
    seven derivative tensors, the output, the seven packed weight images."""
    from gemnet_pytorch_b200 import chain, dense_ops
    from gemnet_pytorch_b200.chain import ACT, CARRY_ADD, SET_CARRY, TO_A, Program, Step, Term
    g = torch.Generator(device="cpu").manual_seed(0)
    xs = [torch.randn(M, 128, generator=g).to(dev) for _ in range(3)]
    pw = [dense_ops.packed_weight((torch.randn(128, 128, generator=g) / 11).to(dev), False) for _ in range(7)]
    bufs = [torch.empty(M, 128, device=dev) for _ in range(8)]
    steps = [Step(pw[0], terms=[Term(xs[0])], flags=ACT | SET_CARRY, dact=bufs[1]),
             Step(pw[1], terms=[Term(xs[1])], flags=ACT | CARRY_ADD | SET_CARRY, dact=bufs[2]),
             Step(pw[2], terms=[Term(xs[2])], flags=ACT | CARRY_ADD | SET_CARRY | TO_A, dact=bufs[3])]
    for j in range(3, 7):
        steps.append(Step(pw[j], flags=ACT | CARRY_ADD | SET_CARRY | TO_A, dact=bufs[1 + j]))
    steps[-1].out = bufs[0]
    prog = Program(M, steps)
    for _ in range(3):
        chain.run(prog)
    times = []
    for _ in range(reps):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        chain.run(prog)
        e.record()
        e.synchronize()
        times.append(s.elapsed_time(e) * 1e-3)
    return float(np.mean(times)), 4 * M * 128 * (3 + 7 + 1) + 7 * 6 * 128 * 128, 7 * 2 * M * 128 * 128


def _ncu_dram_bytes(fname):
    """dram__bytes_read.sum + dram__bytes_write.sum from a committed `ncu --set full` summary (profiles/ncu_summary.py)"""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", fname)
    if not os.path.exists(path):
        return None
    tot = 0.0
    for ln in open(path):
        f = ln.split()
        if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(f[1]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[2]]
    return tot or None


def ncu_traffic_large(n_mol):
    """DRAM traffic of the H7 tcgen05 kernel on the 2048-molecule batch, from the committed `ncu --set full` capture of
    the same launch (profiles/final_run.sh); None for any other size."""
    return _ncu_dram_bytes("ncu_r01_final_segbil_fwd_tc_large.txt") if n_mol == 2048 else None


def ncu_traffic_bench_batch(args, kw, rank):
    """DRAM traffic of the H7 tcgen05 kernel on the batch this bench times by default (64 COLL-shaped molecules of seed
    1234, cutoff 5 A: 8 186 edges, 95 444 triplets), from the committed capture `python profiles/time_h7.py 64 --seed 1234`
    under `ncu --set full` (output written stays in L2: the figure is the DRAM reads); None for any other workload."""
    default = args.batch == 64 and rank == 0 and MOLECULE_KIND[0] == "coll" and kw["cutoff"] == 5.0
    return _ncu_dram_bytes("ncu_r02b_segbil_fwd_tc_b64.txt") if default else None


def pick_cpu_threads(fn):
    """The reference is a chain of small torch CPU ops: using every hardware thread of a 100+ core host makes it
    slower, not faster.  Time one call at a few thread counts (bounded) and keep the fastest."""
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best, best_t = None, None
    for n in [c for c in (8, 16, 32, 64) if c <= avail] or [avail]:
        torch.set_num_threads(n)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
        if dt > 30:
            break
    torch.set_num_threads(best)
    return best, avail


def reference_cpu_model(kw, sd):
    """(predict_fn, kind): the UNMODIFIED reference's GemNet on CPU through oracle/ref_shims.py when it is staged
    (oracle/_ref, oracle/build_ref.sh) or present (/root/reference) — kind "reference" — else the oracle port."""
    from oracle import gemnet_oracle as go
    from oracle import ref_shims
    if ref_shims.reference_available():
        RefGemNet, _ = ref_shims.load_reference()
        kw_ref = dict(kw, scale_file=os.path.join(ref_shims.REFERENCE_ROOT, "pretrained", "scaling_factors.json"))
        model = RefGemNet(**kw_ref)
        model.load_state_dict(sd, strict=True)
        model.eval()
        return (lambda inp: model.predict(dict(inp))), "reference"
    orc = go.OracleGemNet(kw, sd)
    return orc.predict, "port"


def run_reference(args, kw, rank, world):
    """CPU arm: the reference's own forward (energy + autograd forces) on the host cores, rank 0 only."""
    if rank != 0:
        return
    from oracle import gemnet_oracle as go
    from oracle import graph_oracle as gr
    Rl, Zl = synthetic_molecules(args.batch, 1234)
    t0 = time.perf_counter()
    inputs = gr.to_torch_inputs(gr.build_batch(Rl, Zl, kw["cutoff"], kw["int_cutoff"], kw["triplets_only"]))
    t_index = time.perf_counter() - t0
    predict, kind = reference_cpu_model(kw, go.synthetic_state_dict(kw, seed=1234))
    threads, avail = pick_cpu_threads(lambda: predict(inputs))
    for _ in range(args.warmup):
        predict(inputs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        predict(inputs)
    dt = time.perf_counter() - t0
    val = args.batch * args.steps / dt
    what = ("the unmodified reference GemNet.predict (oracle/_ref via oracle/ref_shims.py)" if kind == "reference"
            else "oracle/gemnet_oracle.py (torch CPU restatement of the reference)")
    line = {
        "impl": "reference", "metric": "molecules_per_sec", "value": val, "unit": "molecules/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, kw),
        "cpu_baseline": {"value": val, "unit": "molecules/s", "cores": threads, "host_cores": avail, "kind": kind,
                         "sample": f"{args.steps} energy+force evaluations of the same {args.batch}-molecule batch with "
                                   f"{what}, torch CPU fp32; index build {t_index * 1e3:.0f} ms excluded"},
        "e2e": {"value": val, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_train(args, kw, rank, local_rank, world):
    """BASELINE configs[3]: direct-force GemNet-dT training step on COLL-shaped synthetic molecules, fp32 (the kernels
    are fp32; bf16 storage is follow-up work), whole molecules per rank, ONE flat NCCL all-reduce of the gradients per
    optimiser step (gemnet_pytorch_b200.dist), loss normalised by global molecule / atom counts."""
    import torch.distributed as dist
    from gemnet_pytorch_b200 import GemNet, build_batch
    from gemnet_pytorch_b200 import dist as gdist
    from gemnet_pytorch_b200.weights import seed_synthetic_weights_
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kw = dict(kw, triplets_only=True, direct_forces=True)
    model = seed_synthetic_weights_(GemNet(**kw), seed=1234)
    model = model.to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    use_graph = not args.eager
    opt = torch.optim.AdamW(params, lr=1e-4, amsgrad=True, capturable=use_graph)
    Rl, Zl = synthetic_molecules(args.batch, 1234 + rank)
    inputs = {k: v.to(dev) for k, v in build_batch(Rl, Zl, kw["cutoff"], kw["int_cutoff"], True).items()}
    g = torch.Generator().manual_seed(7 + rank)
    E_t = torch.randn(args.batch, 1, generator=g).to(dev)
    F_t = torch.randn(inputs["R"].shape[0], 3, generator=g).to(dev)
    n_mol, n_atom = gdist.global_counts(args.batch, inputs["R"].shape[0], device=dev)
    from gemnet_pytorch_b200 import engine
    on_engine = engine.trainable(model)

    def step():
        opt.zero_grad(set_to_none=True)
        E, F = model(inputs)
        F = F[:, 0] if F.dim() == 3 else F
        loss = (E - E_t).abs().sum() / n_mol + (F - F_t).norm(dim=-1).sum() / n_atom  # MAE(E) + MAE(|F|) (trainer.py:338-343)
        loss.backward()
        gdist.allreduce_gradients_(params)     # ONE flat NCCL all-reduce (SUM) per step, trainer.py:346-347
        opt.step()
        return loss

    # eager warm-up (also what captures below needs: caches, packed weights, allocator pools) on a side stream
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(3, args.warmup)):
            loss = step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    run = step
    mode = "eager"
    if use_graph:
        try:
            graph_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_):
                static_loss = step()

            def run():
                graph_.replay()
                return static_loss
            mode = "cuda-graph replay of the whole training step (forward, loss, backward, all-reduce, AdamW)"
            for _ in range(args.warmup):
                run()
        except Exception as ex:  # keep the measurement: fall back to the eager step and say so
            mode = f"eager (graph capture failed: {type(ex).__name__}: {str(ex)[:120]})"
            run = step
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        loss = run()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) * 1e-3], dtype=torch.float64, device=dev)
    # the all-reduce on its own (same flat buffer size), CUDA events, max over ranks
    t_ar = torch.zeros(1, dtype=torch.float64, device=dev)
    if world > 1:
        flat = torch.zeros(sum(p.numel() for p in params), device=dev)
        for _ in range(5):
            dist.all_reduce(flat)
        torch.cuda.synchronize()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(20):
            dist.all_reduce(flat)
        e2.record()
        torch.cuda.synchronize()
        t_ar[0] = s2.elapsed_time(e2) * 1e-3 / 20
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_ar, op=dist.ReduceOp.MAX)
        dist.barrier()
    if rank == 0:
        t = float(t)
        n_floats = int(sum(p.numel() for p in params))
        print(json.dumps({
            "metric": "molecules_per_sec", "value": args.batch * world * args.steps / t, "unit": "molecules/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"GemNet-dT direct-force TRAINING step (fwd+bwd+flat NCCL all-reduce+AdamW), "
                                   f"COLL-shaped synthetic batch={args.batch} molecules/GPU, fp32",
                       "grad_floats_allreduced": n_floats},
            "mode": mode, "train_path": "engine: chain-fused tcgen05 kernels forward+backward, weight gradients = library "
                                        "GEMMs on the saved operands" if on_engine else "torch.autograd over the op kernels",
            "allreduce": {"floats": n_floats, "bytes": 4 * n_floats, "ms": float(t_ar) * 1e3,
                          "share_of_step": float(t_ar) / (t / args.steps) if world > 1 else 0.0,
                          "bus_GBps": (2 * (world - 1) / world * 4 * n_floats / float(t_ar) / 1e9) if world > 1 else None},
            "loss": float(loss.detach())}), flush=True)
    # a captured graph keeps NCCL work alive: leave without tearing the communicator down (torchrun reaps the ranks)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sys.stdout.flush()
    os._exit(0)


def workload_config(args, kw):
    name = "GemNet-T" if kw["triplets_only"] else "GemNet-Q"
    shape = "COLL-shaped synthetic" if MOLECULE_KIND[0] == "coll" else "synthetic 18-atom clusters (configs[4])"
    return {"workload": f"{name} energy+force inference (forces = -dE/dR), {shape} batch={args.batch} "
                        f"molecules/GPU, emb_size=128, cutoff {kw['cutoff']:g} A",
            "batch_per_gpu": args.batch, "l2": "flushed (256 MiB write) between timed iterations",
            "weights": "seeded synthetic (random-init architecture weights)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="T", choices=["T", "Q"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--large-batch", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--index-build", default="both", choices=["host", "device", "both"],
                    help="where the e2e path constructs the index arrays (row H0 on host threads / row N2 on the GPU); "
                         "'both' measures the two and reports the faster one as e2e, the other under e2e.other_path")
    ap.add_argument("--cutoff", type=float, default=5.0, help="embedding cutoff in A (BASELINE configs[4]: 8)")
    ap.add_argument("--molecules", default="coll", choices=["coll", "clusters"],
                    help="coll = COLL-shaped synthetic molecules; clusters = the 18-atom clusters of configs[4]")
    ap.add_argument("--eager", action="store_true", help="time the eager step instead of its CUDA-graph replay")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train = BASELINE configs[3]: GemNet-dT direct-force training step (fp32), molecule-sharded, "
                         "one flat NCCL gradient all-reduce per step")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kw = dict(BASE_KW, triplets_only=(args.model == "T"), cutoff=args.cutoff)
    MOLECULE_KIND[0] = args.molecules

    if args.impl == "reference":
        run_reference(args, kw, rank, world)
        return
    if args.mode == "train":
        run_train(args, kw, rank, local_rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from gemnet_pytorch_b200 import GemNet, _lib, build_batch, graph
    from gemnet_pytorch_b200.weights import seed_synthetic_weights_

    model = seed_synthetic_weights_(GemNet(**kw), seed=1234)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}  # (CPU copy, for the cpu_baseline leg)
    model = model.to(dev).eval()

    Rl, Zl = synthetic_molecules(args.batch, 1234 + rank)
    build_batch(Rl[:2], Zl[:2], kw["cutoff"], kw["int_cutoff"], kw["triplets_only"])  # loads the library
    t0 = time.perf_counter()
    host_inputs = build_batch(Rl, Zl, kw["cutoff"], kw["int_cutoff"], kw["triplets_only"], pin_memory=True)
    t_index = time.perf_counter() - t0
    inputs = {k: v.to(dev) for k, v in host_inputs.items()}
    plan = graph.get_plan(inputs, kw["triplets_only"])

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.zero_()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value): the whole energy+force step replayed as one CUDA graph
    for _ in range(args.warmup):
        model(inputs)
    k0 = _lib.KERNEL_LAUNCHES[0]
    model(inputs)
    launches_per_step = _lib.KERNEL_LAUNCHES[0] - k0  # our kernels per step (the graph replays exactly these)
    use_graph = not args.eager
    step = model.make_graphed(inputs) if use_graph else (lambda: model(inputs))
    for _ in range(args.warmup):
        step()

    def timed(fn):
        evs = []
        for _ in range(args.steps):
            flush()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in evs) * 1e-3

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_dev = timed(step)
    barrier()
    clocks = sampler.stop()
    if use_graph:
        for _ in range(3):
            model(inputs)  # re-warm the eager allocator pool after the capture
    t_eager = timed(lambda: model(inputs)) if use_graph else t_dev
    launches = launches_per_step * args.steps

    # ---------------- end-to-end timing through the public call a user makes (GemNet.make_molecule_server):
    #   host molecules (R, Z lists) -> index construction (row H0, native builder, worker threads) -> padding into pinned
    #   staging -> ONE graph launch (H2D, index plan, E+F, D2H) -> (E, F) on the host.
    # Eight DIFFERENT batches are cycled (no per-topology state carries over).  Timed with the HOST clock between
    # synchronisations (perf_counter); the L2 flush kernel of every step runs INSIDE the timed region (conservative).
    #   e2e.value       pipelined throughput: stream() overlaps the index construction of the next batches with the GPU
    #   e2e.latency_ms  one batch at a time, index construction inline (what an MD step pays)
    mol_batches = [(Rl, Zl)] + [synthetic_molecules(args.batch, 1234 + rank + 1000 * j) for j in range(1, 8)]
    max_atoms = max(sum(len(z) for z in Zb) for _, Zb in mol_batches) + 64

    def flushed(batches):
        for b in batches:
            flush()
            yield b

    def measure_server(srv):
        """(seconds for args.steps batches through stream(), median latency of single calls, H2D / D2H bytes per step)"""
        for Rb, Zb in mol_batches:           # warm-up: captures the graph of every capacity bucket these batches need
            srv(Rb, Zb)
        for _ in srv.stream(mol_batches[:4]):
            pass
        h2d_, d2h_ = srv.h2d_bytes, srv.d2h_bytes
        barrier()
        t0 = time.perf_counter()
        for E_h, F_h in srv.stream(flushed(mol_batches[j % 8] for j in range(args.steps))):
            pass
        torch.cuda.synchronize()
        t_stream = time.perf_counter() - t0
        if srv.pool is None:
            d2h_ = E_h.numel() * 4 + F_h.numel() * 4
            if srv.index_build == "host":
                hb = srv.build(Rl, Zl)
                h2d_ = sum(v.numel() * v.element_size() for v in hb.values())
            else:  # coordinates (float64), atomic numbers, molecule sizes, the size header
                h2d_ = sum(len(z) for z in Zl) * (24 + 8) + len(Zl) * (4 + 8) + 64
                d2h_ += 2 * 64
        lat_ = []
        for j in range(min(args.steps, 10)):
            flush()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            srv(*mol_batches[j % 8])
            lat_.append(time.perf_counter() - t0)
        return t_stream, float(np.median(lat_)), int(h2d_), int(d2h_)

    server = model.make_molecule_server(args.batch, max_atoms, workers=3, index_build="host")
    t_e2e, t_lat, h2d, d2h = measure_server(server)
    e2e_paths = {"host": (t_e2e, t_lat, h2d, d2h)}
    if args.index_build in ("device", "both"):
        server_dev = model.make_molecule_server(args.batch, max_atoms, index_build="device")
        e2e_paths["device"] = measure_server(server_dev)
    chosen = args.index_build if args.index_build != "both" else min(e2e_paths, key=lambda k: e2e_paths[k][0])
    t_e2e, t_lat, h2d, d2h = e2e_paths[chosen]
    # the reference-format host batch (indices already built) through the serving pool alone, and through eager predict
    host_batches = [server.build(*mb) for mb in mol_batches[:4]]
    t_serv = t_e2e_eager = None
    if server.pool is not None:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for j in range(args.steps):
            flush()
            server.pool(host_batches[j % 4])
        t_serv = time.perf_counter() - t0
        for j in range(4):
            model.predict({k: v.to(dev, non_blocking=True) for k, v in host_batches[j % 4].items()})
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for j in range(args.steps):
            flush()
            model.predict({k: v.to(dev, non_blocking=True) for k, v in host_batches[j % 4].items()})
        t_e2e_eager = time.perf_counter() - t0

    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = tt.tolist()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, tc_peak, peak_kind = peaks()
    mols = args.batch * world * args.steps
    line = {
        "metric": "molecules_per_sec", "value": mols / t_dev, "unit": "molecules/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, kw), "device": torch.cuda.get_device_name(dev),
        "clocks": clocks,
        "e2e": {"value": mols / t_e2e, "unit": "molecules/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e / args.steps * 1e3,
                "latency_ms": t_lat * 1e3, "host_index_build_ms": t_index * 1e3,
                "index_construction": ("INCLUDED (row H0, native builder on 3 worker threads, pipelined by stream())"
                                       if chosen == "host" else
                                       "INCLUDED (row N2, on the GPU: only float64 coordinates, atomic numbers and molecule "
                                       "sizes cross PCIe; the host counts edges / triplets to pick the capacity bucket)"),
                "index_build": chosen,
                "other_path": {k: {"ms_per_step": v[0] / args.steps * 1e3, "latency_ms": v[1] * 1e3,
                                   "h2d_bytes_per_step": v[2], "d2h_bytes_per_step": v[3]}
                               for k, v in e2e_paths.items() if k != chosen},
                "clock": "host perf_counter between synchronisations; the per-step L2 flush kernel is inside the region",
                "path": (("GemNet.make_molecule_server(index_build='host'): host (R, Z) lists -> H0 -> static-shape CUDA "
                          "graph pool (H2D, index plan, E+F, D2H)" if chosen == "host" else
                          "GemNet.make_molecule_server(index_build='device'): host (R, Z) lists -> static-shape CUDA graph "
                          "pool (H2D of coordinates, index construction kernels + pad, index plan, E+F, D2H)")
                         + f", capacity buckets (edges, triplets) = {sorted(k[:2] for k in (server if chosen == 'host' else server_dev).pool.runners)}")
                if server.pool is not None else
                ("GemNet.make_molecule_server: H0 -> GemNet.predict (eager)" if chosen == "host" else
                 "GemNet.make_molecule_server(index_build='device'): coordinates -> index kernels -> GemNet.predict (eager)"),
                "serving_pool_only_ms_per_step": None if t_serv is None else t_serv / args.steps * 1e3,
                "eager_ms_per_step": None if t_e2e_eager is None else t_e2e_eager / args.steps * 1e3},
        "gpu_launches": int(launches),
        "gpu_launches_per_step": launches_per_step,
        "mode": "cuda-graph replay of the eager step" if use_graph else "eager",
        "eager": {"value": mols / max(t_eager, 1e-12) if world == 1 else None, "ms_per_step": t_eager / args.steps * 1e3},
        "graph": {"atoms": plan.n_atoms, "edges": plan.n_edges, "triplets": plan.n_triplets,
                  "quadruplets": getattr(plan, "n_quads", 0)},
    }

    # ---------------- roofline of the dominant kernel (H7, triplet segmented bilinear)
    t_k, bts, fl = time_bilinear(plan, dev, flush, 30, algo=1)
    t_k0, _, _ = time_bilinear(plan, dev, flush, 30, algo=0)
    line["roofline"] = {"kernel": "segbil_fwd_tc_kernel (H7 triplets, tcgen05 bf16x3 + weight pack kernel)",
                        "bound": "hbm", "achieved": bts / t_k / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bts / t_k / 1e9 / hbm_peak, "traffic": ncu_traffic_bench_batch(args, kw, rank),
                        "peak_kind": peak_kind,
                        "bytes_per_launch": bts, "us_per_launch": t_k * 1e6, "tflops": fl / t_k / 1e12,
                        "tensor_frac_of_bf16_peak": 3 * 2 * plan.n_edges * 1024 * 64 / t_k / 1e12 / tc_peak,
                        "bytes_per_triplet": bts / max(plan.n_triplets, 1),
                        "fp32_cuda_core_kernel_us": t_k0 * 1e6}
    # the kernel with the largest share of the step's GPU time (57 %, profiles/launches_*_summary.txt): fused dense layer
    t_d, b_d, f_d = time_dense(plan.n_edges, dev, flush, 30)
    line["roofline_dense"] = {"kernel": "dense_tc_kernel (fused Dense + ScaledSiLU + derivative, tcgen05 bf16x3x3)",
                              "rows": plan.n_edges, "bound": "hbm", "achieved": b_d / t_d / 1e9, "peak": hbm_peak,
                              "unit": "GB/s", "frac": b_d / t_d / 1e9 / hbm_peak, "us_per_launch": t_d * 1e6,
                              "bytes_per_launch": b_d, "fp32_tflops": f_d / t_d / 1e12, "traffic": None}
    t_c, b_c, f_c = time_chain(plan.n_edges, dev, flush, 20)
    line["roofline_chain"] = {"kernel": "chain_tc_kernel (7 fused dense layers: 3 operand fills, ScaledSiLU + derivative "
                                        "stores, skip connections; tcgen05 bf16x3x3, transposed tiles)",
                              "rows": plan.n_edges, "bound": "hbm", "achieved": b_c / t_c / 1e9, "peak": hbm_peak,
                              "unit": "GB/s", "frac": b_c / t_c / 1e9 / hbm_peak, "us_per_launch": t_c * 1e6,
                              "bytes_per_launch": b_c, "fp32_tflops": f_c / t_c / 1e12, "traffic": None,
                              "note": "latency-bound at batch 64 (128 tiles, 7 dependent GEMM steps per tile)"}
    t_b, b_b = time_h7_backward(plan, dev, flush, 20)
    line["roofline_h7_bwd"] = {"kernel": "segsum + segbil_bwd_edges_tc_kernel (dB = g W'^T on tcgen05) + segbil_bwd_units",
                               "bound": "hbm", "achieved": b_b / t_b / 1e9, "peak": hbm_peak, "unit": "GB/s",
                               "frac": b_b / t_b / 1e9 / hbm_peak, "us_per_call": t_b * 1e6, "bytes_per_call": b_b,
                               "traffic": None}
    if args.large_batch and world == 1:
        Rb, Zb = synthetic_molecules(args.large_batch, 4321)
        big = {k: v.to(dev) for k, v in build_batch(Rb, Zb, kw["cutoff"], kw["int_cutoff"], True).items()}
        bplan = graph.GraphPlan(big, True)
        t_k, bts, fl = time_bilinear(bplan, dev, flush, 10, algo=1)
        t_k0, _, _ = time_bilinear(bplan, dev, flush, 5, algo=0)
        line["roofline_large"] = {"molecules": args.large_batch, "edges": bplan.n_edges, "triplets": bplan.n_triplets,
                                  "bound": "hbm", "achieved": bts / t_k / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": bts / t_k / 1e9 / hbm_peak, "us_per_launch": t_k * 1e6,
                                  "tflops": fl / t_k / 1e12,
                                  "tensor_frac_of_bf16_peak": 3 * 2 * bplan.n_edges * 1024 * 64 / t_k / 1e12 / tc_peak,
                                  "fp32_cuda_core_kernel_us": t_k0 * 1e6,
                                  "bytes_per_launch": bts, "traffic": ncu_traffic_large(args.large_batch)}
        t_c, b_c, f_c = time_chain(bplan.n_edges, dev, flush, 5)
        line["roofline_chain_large"] = {"rows": bplan.n_edges, "bound": "hbm", "achieved": b_c / t_c / 1e9, "peak": hbm_peak,
                                        "unit": "GB/s", "frac": b_c / t_c / 1e9 / hbm_peak, "us_per_launch": t_c * 1e6,
                                        "bytes_per_launch": b_c, "fp32_tflops": f_c / t_c / 1e12,
                                        "tensor_frac_of_bf16_peak": 6 * f_c / t_c / 1e12 / tc_peak}
        t_b, b_b = time_h7_backward(bplan, dev, flush, 5)
        line["roofline_h7_bwd_large"] = {"bound": "hbm", "achieved": b_b / t_b / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                         "frac": b_b / t_b / 1e9 / hbm_peak, "us_per_call": t_b * 1e6, "bytes_per_call": b_b}
        t_d, b_d, f_d = time_dense(bplan.n_edges, dev, flush, 10)
        line["roofline_dense_large"] = {"rows": bplan.n_edges, "bound": "hbm", "achieved": b_d / t_d / 1e9,
                                        "peak": hbm_peak, "unit": "GB/s", "frac": b_d / t_d / 1e9 / hbm_peak,
                                        "us_per_launch": t_d * 1e6, "bytes_per_launch": b_d,
                                        "fp32_tflops": f_d / t_d / 1e12}
        del big, bplan

    # ---------------- CPU baseline beside it: the reference's own forward on the host cores (bounded sample)
    if world == 1 and not args.no_cpu_baseline:
        from oracle import graph_oracle as gr  # the checker / CPU baseline: the only use of oracle/ in this arm
        cpu_inputs = gr.to_torch_inputs(gr.build_batch(Rl, Zl, kw["cutoff"], kw["int_cutoff"], kw["triplets_only"]))
        predict, kind = reference_cpu_model(kw, sd)
        threads, avail = pick_cpu_threads(lambda: predict(cpu_inputs))
        n_rep, t0 = 0, time.perf_counter()
        while n_rep < 5 or (time.perf_counter() - t0 < 10 and n_rep < 40):
            E_c, F_c = predict(cpu_inputs)
            n_rep += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": args.batch * n_rep / dt, "unit": "molecules/s",
                                "cores": threads, "host_cores": avail, "kind": kind,
                                "sample": f"{n_rep} energy+force evaluations of the same {args.batch}-molecule batch with "
                                          + ("the unmodified reference GemNet.predict (oracle/_ref, torch CPU fp32)"
                                             if kind == "reference" else
                                             "oracle/gemnet_oracle.py (torch CPU fp32 restatement of the reference)")}
        # parity of the timed GPU result against the CPU result of the same batch
        E_g, F_g = model.predict(inputs)
        line["parity_vs_cpu"] = {"rel_dE": float((E_g - E_c).abs().max() / E_c.abs().max()),
                                 "rel_dF": float((F_g - F_c).abs().max() / F_c.abs().max())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
