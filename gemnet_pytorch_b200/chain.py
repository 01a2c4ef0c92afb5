"""Chain-fused dense stacks on tcgen05 (include/gemnet_b200.h `gemnet_chain_tc`): host-side program builder.

A `Program` is the list of steps the kernel walks for every 128-row tile — operand fill, GEMM with a packed weight
image, pointwise epilogue (see the header for the exact per-element semantics).  The builders in fused.py describe
ResidualLayer / TripletInteraction / InteractionBlock / AtomUpdateBlock / OutputBlock / EdgeEmbedding stacks
(base_layers.py:61-89, interaction_block.py:186-233,385-421,517-566,653-696, atom_update_block.py:55-72,157-193,
embedding_block.py:70-74) and their input-gradients with it.

`run()` launches the kernel.  `RUNNER[0]` is the launch function: the CUDA library by default; the CPU tests replace it
with a torch interpreter of the same program (tests/chain_emulator.py) to check the programs' algebra without a GPU —
the product never does (no CPU fallback: a CPU tensor reaching `run` raises).
"""
import ctypes

import torch

from . import _lib

MAX_STEPS = 12
MAX_TERMS = 3
ACT, CARRY_ADD, SET_CARRY, TO_A, A_FROM_CARRY = 1, 2, 4, 8, 16

_P = ctypes.c_void_p


class CTerm(ctypes.Structure):
    _fields_ = [("src", _P), ("idx", _P), ("mul", _P), ("scale", ctypes.c_float), ("kind", ctypes.c_int32)]


class CStep(ctypes.Structure):
    _fields_ = [("n_terms", ctypes.c_int32), ("load_set_carry", ctypes.c_int32), ("term", CTerm * MAX_TERMS),
                ("amul", _P), ("Wp", _P), ("K", ctypes.c_int32), ("N", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("pre0", _P), ("pre0_idx", _P), ("pre1", _P), ("pre1_idx", _P),
                ("mul_early", _P), ("mul_late", _P), ("post0", _P), ("post1", _P),
                ("s_acc", ctypes.c_float), ("s_carry", ctypes.c_float), ("s_post0", ctypes.c_float),
                ("s_post1", ctypes.c_float), ("s_store", ctypes.c_float), ("s_set_carry", ctypes.c_float),
                ("s_a", ctypes.c_float), ("s_pad", ctypes.c_float),
                ("aux", _P), ("dact", _P), ("out", _P), ("smul", _P), ("out_idx", _P), ("next_amul", _P),
                ("dot_w", _P), ("dot_out", _P), ("dot_in", _P), ("save_x", _P)]


class CProgram(ctypes.Structure):
    _fields_ = [("M", ctypes.c_int64), ("n_steps", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("step", CStep * MAX_STEPS)]


class Term:
    """One summand of an operand fill: scale * src[idx] (* mul), or (kind 1) the outer product scale * src[r] * mul[c]."""
    __slots__ = ("src", "idx", "mul", "scale", "kind")

    def __init__(self, src, scale=1.0, idx=None, mul=None, kind=0):
        self.src, self.idx, self.mul, self.scale, self.kind = src, idx, mul, float(scale), kind


class Step:
    """One GEMM step.  `weight` is a packed image record of dense_ops.packed_weight (list [buf, N, K, ...])."""
    _TENSORS = ("amul", "pre0", "pre0_idx", "pre1", "pre1_idx", "mul_early", "mul_late", "post0", "post1", "aux",
                "dact", "out", "smul", "out_idx", "next_amul", "dot_w", "dot_out", "dot_in", "save_x")

    def __init__(self, weight, terms=(), load_set_carry=False, flags=0, s_acc=1.0, s_carry=1.0, s_post0=1.0,
                 s_post1=1.0, s_store=1.0, s_set_carry=1.0, s_a=1.0, **tensors):
        self.weight, self.terms, self.load_set_carry, self.flags = weight, list(terms), bool(load_set_carry), int(flags)
        self.s_acc, self.s_carry, self.s_post0, self.s_post1 = float(s_acc), float(s_carry), float(s_post0), float(s_post1)
        self.s_store, self.s_set_carry, self.s_a = float(s_store), float(s_set_carry), float(s_a)
        for k in self._TENSORS:
            setattr(self, k, tensors.pop(k, None))
        if tensors:
            raise TypeError(f"unknown step fields {sorted(tensors)}")
        self.N, self.K = weight[1], weight[2]


class Program:
    def __init__(self, M, steps):
        self.M, self.steps = int(M), list(steps)
        if not 1 <= len(self.steps) <= MAX_STEPS:
            raise _lib.GemnetB200Error(f"chain program needs 1..{MAX_STEPS} steps")


def supported(*dims):
    """K / N sizes the chain kernel takes (multiples of 8 in [8, 128])."""
    return all(8 <= d <= 128 and d % 8 == 0 for d in dims)


def _ptr(t, dtype, what):
    if t is None:
        return None
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype is dtype and t.is_contiguous()):
        raise _lib.GemnetB200Error(f"chain: {what} must be a contiguous CUDA {dtype} tensor (no CPU fallback)")
    return t.data_ptr()


def _launch_cuda(prog, settled):
    c = CProgram()
    c.M, c.n_steps = prog.M, len(prog.steps)
    for i, st in enumerate(prog.steps):
        cs = c.step[i]
        if len(st.terms) > MAX_TERMS:
            raise _lib.GemnetB200Error("chain: too many operand terms")
        cs.n_terms, cs.load_set_carry = len(st.terms), int(st.load_set_carry)
        for j, tm in enumerate(st.terms):
            ct = cs.term[j]
            ct.src, ct.mul = _ptr(tm.src, torch.float32, "term.src"), _ptr(tm.mul, torch.float32, "term.mul")
            ct.idx, ct.scale, ct.kind = _ptr(tm.idx, torch.int32, "term.idx"), tm.scale, tm.kind
        cs.Wp, cs.K, cs.N, cs.flags = st.weight[0].data_ptr(), st.K, st.N, st.flags
        cs.s_acc, cs.s_carry, cs.s_post0, cs.s_post1 = st.s_acc, st.s_carry, st.s_post0, st.s_post1
        cs.s_store, cs.s_set_carry, cs.s_a = st.s_store, st.s_set_carry, st.s_a
        for k in Step._TENSORS:
            setattr(cs, k, _ptr(getattr(st, k), torch.int32 if k.endswith("_idx") else torch.float32, k))
    rc = _lib.load().gemnet_chain_tc(ctypes.byref(c), 1 if settled else 0,
                                     torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    _lib.check(rc, "gemnet_chain_tc")


RUNNER = [_launch_cuda]


def run(prog):
    """Launch one chain program on the current stream."""
    settled = True
    for st in prog.steps:
        if st.weight[4]:  # packed by the most recent launch: the image must not be prefetched early
            st.weight[4], settled = False, False
    RUNNER[0](prog, settled)
