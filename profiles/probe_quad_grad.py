"""How accurate is the quadruplet geometry backward (R -> cbf4 angular part, Y_lm) where the dihedral is well
conditioned?  Masks the upstream gradient of Y_lm to quadruplets whose two in-plane projections are at least `smin` of
the bond length (sin of both polar angles >= smin) and compares with the float64 oracle.  Feeds the tolerance of
tests/test_gpu_ops.py::test_quadruplet_geometry_and_ylm."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import basis_oracle as bo  # noqa: E402  (checker only)
from oracle import gemnet_oracle as go  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_gpu_ops import _quad_plan, batch, kernel_grads, oracle_grads, relerr  # noqa: E402

from gemnet_pytorch_b200 import ops  # noqa: E402

dev = torch.device("cuda")
fx = H.load_npz("ops.npz")
cpu, gpu = batch(fx["mol_ids"].tolist(), False, dev)
plan = _quad_plan(gpu)
R = gpu["R"]
sph_abd, bd_proj = ops.quad_intermediate(R, plan.abd_center, plan.abd_first, plan.abd_second, 7, True)
cs_cab, ac_proj = ops.quad_intermediate(R, plan.cab_center, plan.cab_first, plan.cab_second, 7, False)
ylm = ops.quad_ylm(cs_cab, ac_proj, bd_proj, plan.id4_reduce_cab, plan.id4_expand_abd, 7)
torch.manual_seed(0)
g1, g2 = torch.randn_like(sph_abd), torch.randn_like(ylm)
pc, pa, th = go.quad_angles(cpu["R"].double(), cpu["id_c"], cpu["id_a"], cpu["id4_int_b"], cpu["id4_int_a"],
                            cpu["id4_expand_abd"], cpu["id4_reduce_cab"], cpu["id4_expand_intm_db"],
                            cpu["id4_reduce_intm_ca"], cpu["id4_expand_intm_ab"], cpu["id4_reduce_intm_ab"])
# polar angles of the two planes' generating triplets, per quadruplet
sin_cab = torch.sin(pc)
sin_abd = torch.sin(pa)[cpu["id4_expand_abd"]]


def kern(Rr):
    s, bd = ops.quad_intermediate(Rr, plan.abd_center, plan.abd_first, plan.abd_second, 7, True)
    cs, ac = ops.quad_intermediate(Rr, plan.cab_center, plan.cab_first, plan.cab_second, 7, False)
    return s, ops.quad_ylm(cs, ac, bd, plan.id4_reduce_cab, plan.id4_expand_abd, 7)


def oracle(Rr):
    pc_, pa_, th_ = go.quad_angles(Rr, cpu["id_c"], cpu["id_a"], cpu["id4_int_b"], cpu["id4_int_a"], cpu["id4_expand_abd"],
                                   cpu["id4_reduce_cab"], cpu["id4_expand_intm_db"], cpu["id4_reduce_intm_ca"],
                                   cpu["id4_expand_intm_ab"], cpu["id4_reduce_intm_ab"])
    return bo.legendre_y_l0(pa_, 7), bo.real_sph_harm_all(pc_, th_, 7)


for smin in (0.0, 0.1, 0.3, 0.5):
    keep = ((sin_cab.abs() >= smin) & (sin_abd.abs() >= smin)).float()
    g2m = g2 * keep.to(dev)[:, None]
    (gR,) = kernel_grads(kern, [R], (g1, g2m))
    (gR0,) = oracle_grads(oracle, [cpu["R"]], (g1, g2m))
    print(f"smin={smin}: kept {int(keep.sum())}/{keep.numel()} quadruplets, relerr(gR) = {relerr(gR, gR0):.3e}")
