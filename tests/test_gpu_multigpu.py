"""Two GPUs, NCCL: molecule-sharded data-parallel TRAINING of GemNet-dT on the engine (configs[3]) — the gradients
after the single flat all-reduce (gemnet_pytorch_b200.dist, inserted where trainer.py:346-347 would have it) equal the
single-GPU gradients of the whole batch (sum / global-count loss weighting).  Skipped on a one-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multigpu.py -m gpu` (log: profiles/pytest_multigpu_r02.log)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _loss(E, F, E_t, F_t, n_mol, n_atom):
    F = F[:, 0] if F.dim() == 3 else F
    return (E - E_t).abs().sum() / n_mol + (F - F_t).norm(dim=-1).sum() / n_atom  # trainer.py:338-343 (MAE / MAE)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gemnet_pytorch_b200 import GemNet, build_batch, engine
    from gemnet_pytorch_b200 import dist as gdist
    from oracle import gemnet_oracle as go
    fx = H.load_npz("model_dT.npz")
    kw = H.model_kwargs(fx)
    sd = go.synthetic_state_dict(kw, seed=17)

    def make():
        m = GemNet(**kw)
        m.load_state_dict(sd, strict=True)
        return m.to(dev).train()

    ids = list(range(40, 72))
    N, _, _, _ = H.coll_subset()
    parts = gdist.shard_molecules(N[ids], world, power=3)
    mine = [ids[i] for i in parts[rank]]
    g = torch.Generator().manual_seed(3)
    E_all = torch.randn(len(ids), 1, generator=g)
    F_all = {i: torch.randn(int(N[i]), 3, generator=g) for i in ids}

    def batch(mol_ids):
        inp = {k: v.to(dev) for k, v in build_batch(*H.molecules(mol_ids), kw["cutoff"], kw["int_cutoff"], True).items()}
        E_t = torch.stack([E_all[ids.index(i)] for i in mol_ids]).to(dev)
        F_t = torch.cat([F_all[i] for i in mol_ids]).to(dev)
        return inp, E_t, F_t

    model = make()
    assert engine.trainable(model)
    inp, E_t, F_t = batch(mine)
    n_mol, n_atom = gdist.global_counts(len(mine), inp["R"].shape[0], device=dev)
    assert n_mol == len(ids)
    E, F = model(inp)
    _loss(E, F, E_t, F_t, n_mol, n_atom).backward()
    params = [p for p in model.parameters() if p.requires_grad]
    n = gdist.allreduce_gradients_(params)
    if rank == 0:
        ref = make()
        inp, E_t, F_t = batch(ids)
        E, F = ref(inp)
        _loss(E, F, E_t, F_t, len(ids), inp["R"].shape[0]).backward()
        worst = 0.0
        for (name, p), q in zip(model.named_parameters(), ref.parameters()):
            if p.grad is None or q.grad is None:
                assert p.grad is None and q.grad is None, name
                continue
            worst = max(worst, float((p.grad - q.grad).abs().max() / q.grad.abs().max().clamp(min=1e-12)))
        ret.put((worst, n))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradients_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    worst, n = ret.get(timeout=10)
    print(f"DP (2 x B200, NCCL flat all-reduce of {n} floats) vs single GPU: worst relative gradient difference {worst:.2e}")
    assert n > 2_000_000
    assert worst < 2e-4   # float32 summation order (per-rank partial sums + force atomics)
