"""CPU, world_size 2, gloo: the N>1 host logic — molecule sharding and the single flat gradient all-reduce with
global-count loss normalisation reproduce the single-process gradient exactly."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gemnet_pytorch_b200 import dist as gdist


def test_shard_molecules_covers_and_balances():
    rng = np.random.default_rng(0)
    n = rng.integers(2, 27, size=257)
    for world in (1, 2, 4, 8):
        parts = gdist.shard_molecules(n, world, power=3)
        allidx = sorted(i for p in parts for i in p)
        assert allidx == list(range(len(n)))
        loads = [float((n[p].astype(float) ** 3).sum()) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.05


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 8), torch.nn.SiLU(), torch.nn.Linear(8, 1))
    g = torch.Generator().manual_seed(1)
    n_atoms = torch.randint(2, 9, (12,), generator=g)               # 12 "molecules"
    feats = [torch.randn(int(k), 5, generator=g) for k in n_atoms]
    tgtE = torch.randn(12, generator=g)
    parts = gdist.shard_molecules(n_atoms.numpy(), world, power=3)
    mine = parts[rank]
    nmol, natom = gdist.global_counts(len(mine), int(sum(int(n_atoms[i]) for i in mine)))
    assert (nmol, natom) == (12, int(n_atoms.sum()))
    # energy term: mean over molecules; "force" term: mean over atoms  (trainer.py:280-290,338-343)
    loss = 0.0
    for i in mine:
        per_atom = model(feats[i]).squeeze(-1)
        loss = loss + (per_atom.sum() - tgtE[i]).abs() / nmol + per_atom.abs().sum() / natom
    loss.backward()
    n = gdist.allreduce_gradients_(model.parameters())
    assert n == sum(p.numel() for p in model.parameters())
    if rank == 0:
        ref = torch.nn.Sequential(torch.nn.Linear(5, 8), torch.nn.SiLU(), torch.nn.Linear(8, 1))
        ref.load_state_dict(model.state_dict())
        tot = 0.0
        for i in range(12):
            per_atom = ref(feats[i]).squeeze(-1)
            tot = tot + (per_atom.sum() - tgtE[i]).abs() / 12 + per_atom.abs().sum() / int(n_atoms.sum())
        tot.backward()
        err = max((a.grad - b.grad).abs().max().item() for a, b in zip(model.parameters(), ref.parameters()))
        ret.put(err)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) < 1e-6
