"""torch.autograd.Function wrappers around the C-ABI kernels (include/gemnet_b200.h).

Each function mirrors one reference layer call (cited in the header) and has a hand-written CUDA
backward, so forces -dE/dR (gemnet/model/gemnet.py:598-613) come out of torch.autograd through the
kernels.  Backwards are first-order only (`once_differentiable`): training with direct_forces=False
needs a double backward and is not supported yet.

There is no fallback: tensors must be CUDA float32 / int32 and the library must be built.
"""
import numpy as np
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from . import basis_consts

_check = _lib.check


def _L():
    return _lib.load()


def _stream():
    """Raw cudaStream_t of torch's current stream on the current device.  torch.cuda.current_stream() costs ~20 us of
    Python per call (device-index resolution through os.environ); the two raw C getters are < 1 us — this is on the
    path of every kernel launch."""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _p(t):
    return None if t is None else t.data_ptr()


def _f32(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float32, "gemnet_b200 ops need CUDA float32 tensors (no CPU fallback)"
    return t if t.is_contiguous() else t.contiguous()


def _i32(t):
    assert t.is_cuda and t.dtype == torch.int32 and t.is_contiguous(), "index tensors must be contiguous CUDA int32"
    return t


# ----------------------------------------------------------------------------- index preparation
def csr_from_sorted(sorted_ids, n_rows):
    """(rowptr int32[n_rows+1], flag int32[1])"""
    ids = _i32(sorted_ids)
    rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=ids.device)
    flag = torch.zeros(1, dtype=torch.int32, device=ids.device)
    _check(_L().gemnet_csr_from_sorted(_p(ids), ids.numel(), n_rows, _p(rowptr), _p(flag), _stream()),
           "gemnet_csr_from_sorted")
    return rowptr, flag


def csr_by_key(ids, n_rows, flag=None):
    """(rowptr int32[n_rows+1], perm int32[n]) — stable counting sort of positions by id.  `flag` (int32[1], zeroed by
    the caller) is raised by an id outside [0, n_rows) (such entries are skipped, not written out of bounds)."""
    ids = _i32(ids)
    rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=ids.device)
    perm = torch.empty(ids.numel(), dtype=torch.int32, device=ids.device)
    ws = torch.empty(n_rows + 1 + ids.numel(), dtype=torch.int32, device=ids.device)
    _check(_L().gemnet_csr_by_key(_p(ids), ids.numel(), n_rows, _p(rowptr), _p(perm), _p(ws), _p(flag), _stream()),
           "gemnet_csr_by_key", kernels=5)
    return rowptr, perm


def index_range_check(lists, flag):
    """lists: [(int32 CUDA tensor, bound), ...] (at most 12); raises `flag` (int32[1]) if an id lies outside [0, bound)."""
    import ctypes
    lists = [(_i32(t), b) for t, b in lists if t.numel()]
    if not lists:
        return
    n = len(lists)
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t, _ in lists])
    sizes = (ctypes.c_int64 * n)(*[t.numel() for t, _ in lists])
    bounds = (ctypes.c_int64 * n)(*[int(b) for _, b in lists])
    _check(_L().gemnet_index_range_check(ptrs, sizes, bounds, n, _p(flag), _stream()), "gemnet_index_range_check")


# ----------------------------------------------------------------------------- H1
class _EdgeGeometry(Function):
    @staticmethod
    def forward(ctx, R, id_c, id_a):
        R = _f32(R)
        E = id_c.numel()
        D = torch.empty(E, dtype=torch.float32, device=R.device)
        V = torch.empty(E, 3, dtype=torch.float32, device=R.device)
        _check(_L().gemnet_edge_geometry_fwd(_p(R), _p(_i32(id_c)), _p(_i32(id_a)), E, _p(D), _p(V), _stream()),
               "gemnet_edge_geometry_fwd")
        ctx.save_for_backward(D, V, id_c, id_a)
        ctx.n_atoms = R.shape[0]
        return D, V

    @staticmethod
    @once_differentiable
    def backward(ctx, gD, gV):
        D, V, id_c, id_a = ctx.saved_tensors
        gR = torch.zeros(ctx.n_atoms, 3, dtype=torch.float32, device=D.device)
        gD, gV = _f32(gD), _f32(gV)
        _check(_L().gemnet_edge_geometry_bwd(_p(D), _p(V), _p(id_c), _p(id_a), D.numel(), _p(gD), _p(gV), _p(gR),
                                             _stream()), "gemnet_edge_geometry_bwd")
        return gR, None, None


def edge_geometry(R, id_c, id_a):
    """GemNet.calculate_interatomic_vectors (gemnet.py:262-286) -> D (E,), V (E,3)."""
    return _EdgeGeometry.apply(R, id_c, id_a)


# ----------------------------------------------------------------------------- H3
class _BesselRbf(Function):
    @staticmethod
    def forward(ctx, D, freq, cutoff, p):
        D, freq = _f32(D), _f32(freq)
        nR = freq.numel()
        out = torch.empty(D.numel(), nR, dtype=torch.float32, device=D.device)
        _check(_L().gemnet_bessel_rbf_fwd(_p(D), _p(freq), D.numel(), nR, cutoff, p, _p(out), _stream()),
               "gemnet_bessel_rbf_fwd")
        ctx.save_for_backward(D, freq)
        ctx.cfg = (cutoff, p)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        D, freq = ctx.saved_tensors
        cutoff, p = ctx.cfg
        g = _f32(g)
        gD = torch.empty_like(D)
        gf = torch.zeros_like(freq) if ctx.needs_input_grad[1] else None
        _check(_L().gemnet_bessel_rbf_bwd(_p(D), _p(freq), D.numel(), freq.numel(), cutoff, p, _p(g), _p(gD), _p(gf),
                                          _stream()), "gemnet_bessel_rbf_bwd")
        return gD, gf, None, None


def bessel_rbf(D, freq, cutoff, envelope_p):
    """BesselBasisLayer.forward (basis_layers.py:45-49)."""
    return _BesselRbf.apply(D, freq, float(cutoff), int(envelope_p))


# ----------------------------------------------------------------------------- H4/H5 radial
_tables = {}


def radial_tables(num_spherical, num_radial, device):
    key = (num_spherical, num_radial, str(device))
    if key not in _tables:
        z, n = basis_consts.bessel_tables(num_spherical, num_radial)
        _tables[key] = (torch.tensor(z, dtype=torch.float64, device=device).contiguous(),
                        torch.tensor(n, dtype=torch.float64, device=device).contiguous())
    return _tables[key]


_small = {}


def const_tensor(name, builder, device, dtype):
    key = (name, str(device))
    if key not in _small:
        _small[key] = torch.tensor(np.asarray(builder()), dtype=dtype, device=device).contiguous()
    return _small[key]


class _SphBesselRadial(Function):
    @staticmethod
    def forward(ctx, D, S, nR, cutoff, p):
        D = _f32(D)
        z, n = radial_tables(S, nR, D.device)
        out = torch.empty(D.numel(), S, nR, dtype=torch.float32, device=D.device)
        _check(_L().gemnet_sph_bessel_radial_fwd(_p(D), D.numel(), S, nR, _p(z), _p(n), cutoff, p, _p(out),
                                                 _stream()), "gemnet_sph_bessel_radial_fwd")
        ctx.save_for_backward(D)
        ctx.cfg = (S, nR, cutoff, p)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (D,) = ctx.saved_tensors
        S, nR, cutoff, p = ctx.cfg
        z, n = radial_tables(S, nR, D.device)
        g = _f32(g)
        gD = torch.empty_like(D)
        _check(_L().gemnet_sph_bessel_radial_bwd(_p(D), D.numel(), S, nR, _p(z), _p(n), cutoff, p, _p(g), _p(gD),
                                                 _stream()), "gemnet_sph_bessel_radial_bwd")
        return gD, None, None, None, None


def sph_bessel_radial(D, num_spherical, num_radial, cutoff, envelope_p):
    """rbf_env of SphericalBasisLayer / TensorBasisLayer (basis_layers.py:121-128), layout (E, S, R)."""
    return _SphBesselRadial.apply(D, int(num_spherical), int(num_radial), float(cutoff), int(envelope_p))


# ----------------------------------------------------------------------------- H2 + H4 angular (triplets)
class _TripletSph(Function):
    @staticmethod
    def forward(ctx, R, id_c, id_a, red, exp, S):
        R = _f32(R)
        T = red.numel()
        pref = const_tensor(f"pref0_{S}", lambda: basis_consts.legendre_prefactors(S), R.device, torch.float32)
        sph = torch.empty(T, S, dtype=torch.float32, device=R.device)
        _check(_L().gemnet_triplet_sph_fwd(_p(R), _p(_i32(id_c)), _p(_i32(id_a)), _p(_i32(red)), _p(_i32(exp)), T, S,
                                           _p(pref), _p(sph), None, _stream()), "gemnet_triplet_sph_fwd")
        ctx.save_for_backward(R, id_c, id_a, red, exp, pref)
        ctx.S = S
        return sph

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        R, id_c, id_a, red, exp, pref = ctx.saved_tensors
        g = _f32(g)
        gR = torch.zeros_like(R)
        _check(_L().gemnet_triplet_sph_bwd(_p(R), _p(id_c), _p(id_a), _p(red), _p(exp), red.numel(), ctx.S, _p(pref),
                                           _p(g), _p(gR), _stream()), "gemnet_triplet_sph_bwd")
        return gR, None, None, None, None, None


def triplet_sph(R, id_c, id_a, id3_reduce_ca, id3_expand_ba, num_spherical):
    """calculate_angles3 (gemnet.py:421-451) + Y_l0 of SphericalBasisLayer (basis_layers.py:130-131): (T, S), unpadded."""
    return _TripletSph.apply(R, id_c, id_a, id3_reduce_ca, id3_expand_ba, int(num_spherical))


def triplet_angles(R, id_c, id_a, id3_reduce_ca, id3_expand_ba, num_spherical=1):
    """Row H2 on its own (no autograd): the angle the reference feeds to its basis layer."""
    R = _f32(R)
    T = id3_reduce_ca.numel()
    S = int(num_spherical)
    pref = const_tensor(f"pref0_{S}", lambda: basis_consts.legendre_prefactors(S), R.device, torch.float32)
    sph = torch.empty(T, S, dtype=torch.float32, device=R.device)
    ang = torch.empty(T, dtype=torch.float32, device=R.device)
    _check(_L().gemnet_triplet_sph_fwd(_p(R), _p(_i32(id_c)), _p(_i32(id_a)), _p(_i32(id3_reduce_ca)),
                                       _p(_i32(id3_expand_ba)), T, S, _p(pref), _p(sph), _p(ang), _stream()),
           "gemnet_triplet_sph_fwd")
    return ang


# ----------------------------------------------------------------------------- H2 + H4/H5 angular (quadruplets)
class _QuadIntm(Function):
    """mode 'sph': returns (sph (M,S), proj (M,3)); mode 'cossin': returns (cossin (M,2), proj (M,3))."""

    @staticmethod
    def forward(ctx, R, center, first, second, S, want_sph):
        R = _f32(R)
        M = center.numel()
        pref = const_tensor(f"pref0_{S}", lambda: basis_consts.legendre_prefactors(S), R.device, torch.float32)
        sph = torch.empty(M, S, dtype=torch.float32, device=R.device) if want_sph else None
        cs = None if want_sph else torch.empty(M, 2, dtype=torch.float32, device=R.device)
        proj = torch.empty(M, 3, dtype=torch.float32, device=R.device)
        _check(_L().gemnet_quad_intm_fwd(_p(R), _p(_i32(center)), _p(_i32(first)), _p(_i32(second)), M, S, _p(pref),
                                         _p(sph), _p(cs), _p(proj), _stream()), "gemnet_quad_intm_fwd")
        ctx.save_for_backward(R, center, first, second, pref)
        ctx.cfg = (S, want_sph)
        return (sph if want_sph else cs), proj

    @staticmethod
    @once_differentiable
    def backward(ctx, g_first, g_proj):
        R, center, first, second, pref = ctx.saved_tensors
        S, want_sph = ctx.cfg
        g_first, g_proj = _f32(g_first), _f32(g_proj)
        gR = torch.zeros_like(R)
        _check(_L().gemnet_quad_intm_bwd(_p(R), _p(center), _p(first), _p(second), center.numel(), S, _p(pref),
                                         _p(g_first) if want_sph else None, None if want_sph else _p(g_first),
                                         _p(g_proj), _p(gR), _stream()), "gemnet_quad_intm_bwd")
        return gR, None, None, None, None, None


def quad_intermediate(R, center, first, second, num_spherical, want_sph):
    return _QuadIntm.apply(R, center, first, second, int(num_spherical), bool(want_sph))


class _QuadYlm(Function):
    @staticmethod
    def forward(ctx, cossin, ac_proj, bd_proj, red_cab, exp_abd, S):
        cossin, ac_proj, bd_proj = _f32(cossin), _f32(ac_proj), _f32(bd_proj)
        Q = red_cab.numel()
        pref = const_tensor(f"preflm_{S}", lambda: basis_consts.real_sph_harm_prefactors(S), cossin.device,
                            torch.float32)
        ylm = torch.empty(Q, S * S, dtype=torch.float32, device=cossin.device)
        _check(_L().gemnet_quad_ylm_fwd(_p(cossin), _p(ac_proj), _p(bd_proj), _p(_i32(red_cab)), _p(_i32(exp_abd)), Q,
                                        S, _p(pref), _p(ylm), None, _stream()), "gemnet_quad_ylm_fwd")
        ctx.save_for_backward(cossin, ac_proj, bd_proj, red_cab, exp_abd, pref)
        ctx.S = S
        return ylm

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        cossin, ac_proj, bd_proj, red_cab, exp_abd, pref = ctx.saved_tensors
        g = _f32(g)
        g_cs, g_ac, g_bd = torch.zeros_like(cossin), torch.zeros_like(ac_proj), torch.zeros_like(bd_proj)
        _check(_L().gemnet_quad_ylm_bwd(_p(cossin), _p(ac_proj), _p(bd_proj), _p(red_cab), _p(exp_abd),
                                        red_cab.numel(), ctx.S, _p(pref), _p(g), _p(g_cs), _p(g_ac), _p(g_bd),
                                        _stream()), "gemnet_quad_ylm_bwd")
        return g_cs, g_ac, g_bd, None, None, None


def quad_ylm(cossin_cab, ac_proj, bd_proj, id4_reduce_cab, id4_expand_abd, num_spherical):
    """Real Y_lm(Phi_cab, Theta_cabd) of TensorBasisLayer (basis_layers.py:269-270): (Q, S^2), unpadded."""
    return _QuadYlm.apply(cossin_cab, ac_proj, bd_proj, id4_reduce_cab, id4_expand_abd, int(num_spherical))


def quad_angles(cossin_cab, ac_proj, bd_proj, id4_reduce_cab, id4_expand_abd, num_spherical):
    """Row H2 for quadruplets on its own (no autograd): (Q,2) = (Phi_cab, Theta_cabd)."""
    S = int(num_spherical)
    Q = id4_reduce_cab.numel()
    pref = const_tensor(f"preflm_{S}", lambda: basis_consts.real_sph_harm_prefactors(S), cossin_cab.device,
                        torch.float32)
    ylm = torch.empty(Q, S * S, dtype=torch.float32, device=cossin_cab.device)
    tp = torch.empty(Q, 2, dtype=torch.float32, device=cossin_cab.device)
    _check(_L().gemnet_quad_ylm_fwd(_p(_f32(cossin_cab)), _p(_f32(ac_proj)), _p(_f32(bd_proj)),
                                    _p(_i32(id4_reduce_cab)), _p(_i32(id4_expand_abd)), Q, S, _p(pref), _p(ylm),
                                    _p(tp), _stream()), "gemnet_quad_ylm_fwd")
    return tp


# ----------------------------------------------------------------------------- H6
class _DownProj(Function):
    @staticmethod
    def forward(ctx, rad, Wd, l_of_s):
        rad, Wd = _f32(rad), _f32(Wd)
        E, L, nR = rad.shape
        S, _, I = Wd.shape
        out = torch.empty(E, I, S, dtype=torch.float32, device=rad.device)
        _check(_L().gemnet_downproj_fwd(_p(rad), _p(Wd), _p(_i32(l_of_s)), E, L, nR, S, I, _p(out), _stream()),
               "gemnet_downproj_fwd")
        ctx.save_for_backward(rad, Wd, l_of_s)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        rad, Wd, l_of_s = ctx.saved_tensors
        g = _f32(g)
        E, L, nR = rad.shape
        S, _, I = Wd.shape
        g_rad = torch.empty_like(rad) if ctx.needs_input_grad[0] else None
        g_Wd = torch.zeros_like(Wd) if ctx.needs_input_grad[1] else None
        _check(_L().gemnet_downproj_bwd(_p(rad), _p(Wd), _p(l_of_s), E, L, nR, S, I, _p(g), _p(g_rad), _p(g_Wd),
                                        _stream()), "gemnet_downproj_bwd",
               kernels=int(g_rad is not None) + int(g_Wd is not None))
        return g_rad, g_Wd, None


def down_projection(rad, Wd, l_of_s):
    """EfficientInteractionDownProjection.forward (efficient.py:41-57): (E,L,R),(S,R,I) -> rbf_W1 (E,I,S)."""
    return _DownProj.apply(rad, Wd, l_of_s)


# ----------------------------------------------------------------------------- H7
def _segbilinear_weight_grad(x, expand, rowptr, sph, rbf_W1, g_out_scaled, O=None):
    """dL/dW (C,I,O) — training only: A (E,S,C) by the CSR segment-sum kernel (no (U,S,C) intermediate: that would be
    8 GB for a GemNet-dQ batch), then two library contractions."""
    E, I, S = rbf_W1.shape
    C = x.shape[1]
    A = torch.empty(E, S, C, dtype=torch.float32, device=x.device)
    _check(_L().gemnet_segbilinear_segsum(_p(x), _p(_i32(expand)), _p(_i32(rowptr)), _p(sph), E, S, C, I,
                                          g_out_scaled.shape[1], _p(A), _stream()), "gemnet_segbilinear_segsum")
    B = torch.bmm(rbf_W1, A)                                   # (E, I, C) = (E,I,S) @ (E,S,C)
    return torch.mm(B.transpose(1, 2).reshape(E, C * I).t(), g_out_scaled).view(C, I, -1)


_segbil_packs = {}


def segbil_packed(W, backward):
    """bf16 hi/lo image of the bilinear weight for the tcgen05 kernels (gemnet_segbilinear_pack), packed once per
    weight version (see dense_ops.packed_weight for the `.data` caveat; dense_ops.invalidate() clears this cache too)."""
    C, I, O = W.shape
    key = (W.data_ptr(), (C, I, O), bool(backward))
    hit = _segbil_packs.get(key)
    if hit is not None and hit[2] == W._version:
        return hit[0]
    S = 7
    n = _L().gemnet_segbilinear_pack_bytes(S, C, I, O, int(backward))
    buf = hit[0] if hit is not None else torch.empty(n, dtype=torch.uint8, device=W.device)
    _check(_L().gemnet_segbilinear_pack(_p(_f32(W)), S, C, I, O, int(backward), _p(buf), _stream()),
           "gemnet_segbilinear_pack")
    _segbil_packs[key] = (buf, W, W._version)
    return buf


def _tc_shape(S, C, I, O):
    return (S, C, I, O) == (7, 64, 16, 64)


class _SegBilinear(Function):
    @staticmethod
    def forward(ctx, x, sph, rbf_W1, W, expand, rowptr, scale, algo):
        x, sph, rbf_W1, W = _f32(x), _f32(sph), _f32(rbf_W1), _f32(W)
        E, I, S = rbf_W1.shape
        C, _, O = W.shape
        U = expand.numel()
        out = torch.empty(E, O, dtype=torch.float32, device=x.device)
        if algo and _tc_shape(S, C, I, O):
            # tcgen05 kernel with the weight image packed once per weight version (algo 3 = TC | PREPACKED)
            _check(_L().gemnet_segbilinear_fwd(_p(x), _p(_i32(expand)), _p(_i32(rowptr)), _p(sph), _p(rbf_W1),
                                               _p(segbil_packed(W, False)), scale, E, U, S, C, I, O, 3, _p(out), None,
                                               _stream()), "gemnet_segbilinear_fwd")
        else:
            algo = 0
            nws = _L().gemnet_segbilinear_workspace_bytes(E, S, C, I, O, algo, 0)
            keep = nws > 0 and (any(ctx.needs_input_grad[:3]) if hasattr(ctx, "needs_input_grad") else False)
            if keep:  # split pipeline + a backward will follow: keep the segment sums for it (workspace of bwd size)
                nws = _L().gemnet_segbilinear_workspace_bytes(E, S, C, I, O, algo, 1)
            ws = torch.empty(nws, dtype=torch.uint8, device=x.device) if nws > 0 else None
            _check(_L().gemnet_segbilinear_fwd(_p(x), _p(_i32(expand)), _p(_i32(rowptr)), _p(sph), _p(rbf_W1), _p(W),
                                               scale, E, U, S, C, I, O, algo, _p(out), _p(ws), _stream()),
                   "gemnet_segbilinear_fwd", kernels=1 if ws is None else 2)
            ctx.ws_saved = ws if keep else None
        ctx.save_for_backward(x, sph, rbf_W1, W, expand, rowptr)
        ctx.cfg = (scale, algo)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, sph, rbf_W1, W, expand, rowptr = ctx.saved_tensors
        scale, algo = ctx.cfg
        g = _f32(g)
        E, I, S = rbf_W1.shape
        C, _, O = W.shape
        U = expand.numel()
        g_x = torch.zeros_like(x) if ctx.needs_input_grad[0] else None
        g_sph = torch.empty_like(sph) if ctx.needs_input_grad[1] else None
        g_rw = torch.empty_like(rbf_W1) if ctx.needs_input_grad[2] else None
        if g_x is not None or g_sph is not None or g_rw is not None:
            tc = bool(algo) and _tc_shape(S, C, I, O) and BWD_TC[0]
            a = 3 if tc else 0
            nws = _L().gemnet_segbilinear_workspace_bytes(E, S, C, I, O, a, 1)
            ws = getattr(ctx, "ws_saved", None)
            if ws is not None and not tc and ws.numel() >= nws:
                a |= 4  # GEMNET_SEGBIL_A_SAVED: the forward's segment sums are in the workspace head
            else:
                ws = torch.empty(nws, dtype=torch.uint8, device=x.device) if nws > 0 else None
            Wk = segbil_packed(W, True) if tc else W
            _check(_L().gemnet_segbilinear_bwd(_p(x), _p(expand), _p(rowptr), _p(sph), _p(rbf_W1), _p(Wk), scale, E, U,
                                               S, C, I, O, a, _p(g), _p(g_x), _p(g_sph), _p(g_rw), None, _p(ws),
                                               _stream()), "gemnet_segbilinear_bwd", kernels=3 if ws is not None else 1)
        g_W = _segbilinear_weight_grad(x, expand, rowptr, sph, rbf_W1, g * scale) if ctx.needs_input_grad[3] else None
        return g_x, g_sph, g_rw, g_W, None, None, None, None


import os as _os

BWD_TC = [_os.environ.get("GEMNET_B200_H7_BWD_TC", "1") != "0"]  # dB = g W'^T of the triplet shape on tcgen05


def segmented_bilinear(x, sph, rbf_W1, W, expand, rowptr, scale=1.0, algo=0):
    """EfficientInteractionBilinear.forward (efficient.py:159-189) fused with the source-edge gather
    (interaction_block.py:678/562) and the following ScalingFactor (interaction_block.py:682/555).

    x (n_x,C) un-gathered table; expand (U,) row of x per unit; rowptr (E+1,) CSR over units by target edge;
    sph (U,S) un-padded; rbf_W1 (E,I,S); W (C,I,O)  ->  (E,O)."""
    return _SegBilinear.apply(x, sph, rbf_W1, W, expand, rowptr, float(scale), int(algo))


# ----------------------------------------------------------------------------- H10 / H11
class _EdgeToAtom(Function):
    @staticmethod
    def forward(ctx, m, r, id_a, perm, rowptr, scale, had_scale):
        m, r = _f32(m), _f32(r)
        E, F = m.shape
        A = rowptr.numel() - 1
        out = torch.empty(A, F, dtype=torch.float32, device=m.device)
        had = torch.empty_like(m) if had_scale is not None else None
        _check(_L().gemnet_edge_to_atom_fwd(_p(m), _p(r), _p(_i32(perm)), _p(_i32(rowptr)), A, E, F, scale, _p(out),
                                            0.0 if had_scale is None else had_scale, _p(had), _stream()),
               "gemnet_edge_to_atom_fwd", kernels=1 if had is None else 2)
        ctx.save_for_backward(m, r, id_a)
        ctx.cfg = (scale, had_scale)
        if had is None:
            return out
        return out, had

    @staticmethod
    @once_differentiable
    def backward(ctx, g_out, g_had=None):
        m, r, id_a = ctx.saved_tensors
        scale, had_scale = ctx.cfg
        g_out = _f32(g_out)
        g_had = _f32(g_had) if had_scale is not None else None
        E, F = m.shape
        g_m = torch.empty_like(m) if ctx.needs_input_grad[0] else None
        g_r = torch.empty_like(r) if ctx.needs_input_grad[1] else None
        _check(_L().gemnet_edge_to_atom_bwd(_p(m), _p(r), _p(_i32(id_a)), E, F, scale, _p(g_out),
                                            0.0 if had_scale is None else had_scale, _p(g_had), _p(g_m), _p(g_r),
                                            _stream()), "gemnet_edge_to_atom_bwd")
        return g_m, g_r, None, None, None, None, None


def edge_to_atom(m, r, id_a, perm, rowptr, scale=1.0, had_scale=None):
    """scale * scatter_add(m * r, id_a) (atom_update_block.py:64-68,166-173) [+ had_scale * m * r (:181-182)]."""
    return _EdgeToAtom.apply(m, r, id_a, perm, rowptr, float(scale), None if had_scale is None else float(had_scale))


# ----------------------------------------------------------------------------- H12
class _MoleculeSum(Function):
    @staticmethod
    def forward(ctx, E_atom, batch_seg, mol_rowptr, mean):
        E_atom = _f32(E_atom)
        nMol = mol_rowptr.numel() - 1
        nT = E_atom.shape[1]
        out = torch.empty(nMol, nT, dtype=torch.float32, device=E_atom.device)
        _check(_L().gemnet_molecule_sum_fwd(_p(E_atom), _p(_i32(mol_rowptr)), nMol, nT, int(mean), _p(out), _stream()),
               "gemnet_molecule_sum_fwd")
        ctx.save_for_backward(batch_seg, mol_rowptr)
        ctx.cfg = (E_atom.shape[0], nT, int(mean))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        batch_seg, mol_rowptr = ctx.saved_tensors
        nA, nT, mean = ctx.cfg
        g = _f32(g)
        gE = torch.empty(nA, nT, dtype=torch.float32, device=g.device)
        _check(_L().gemnet_molecule_sum_bwd(_p(g), _p(_i32(batch_seg)), _p(mol_rowptr), nA, nT, mean, _p(gE),
                                            _stream()), "gemnet_molecule_sum_bwd")
        return gE, None, None, None


def molecule_sum(E_atom, batch_seg, mol_rowptr, mean=False):
    """scatter(E_a, batch_seg, reduce="add"|"mean") (gemnet.py:578-584)."""
    return _MoleculeSum.apply(E_atom, batch_seg, mol_rowptr, bool(mean))


class _ForceReadout(Function):
    @staticmethod
    def forward(ctx, F_ca, V, id_a, perm, rowptr):
        F_ca, V = _f32(F_ca), _f32(V)
        A = rowptr.numel() - 1
        nT = F_ca.shape[1]
        out = torch.empty(A, nT, 3, dtype=torch.float32, device=F_ca.device)
        _check(_L().gemnet_force_readout_fwd(_p(F_ca), _p(V), _p(_i32(perm)), _p(_i32(rowptr)), A, nT, _p(out),
                                             _stream()), "gemnet_force_readout_fwd")
        ctx.save_for_backward(F_ca, V, id_a)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        F_ca, V, id_a = ctx.saved_tensors
        g = _f32(g)
        gF = torch.empty_like(F_ca) if ctx.needs_input_grad[0] else None
        gV = torch.empty_like(V) if ctx.needs_input_grad[1] else None
        _check(_L().gemnet_force_readout_bwd(_p(F_ca), _p(V), _p(_i32(id_a)), F_ca.shape[0], F_ca.shape[1], _p(g),
                                             _p(gF), _p(gV), _stream()), "gemnet_force_readout_bwd")
        return gF, gV, None, None, None


def force_readout(F_ca, V, id_a, perm, rowptr):
    """F_j = scatter_add(F_ca[:, :, None] * V_ca[:, None, :], id_a) (gemnet.py:594-597): (A, T, 3)."""
    return _ForceReadout.apply(F_ca, V, id_a, perm, rowptr)


# ----------------------------------------------------------------------------- N1: fused pointwise + gathers
def _n4(t):
    return t.numel()  # any length: the kernels vectorise the first n - n % 4 elements and finish the tail scalar


class _ScaledSiLU(Function):
    @staticmethod
    def forward(ctx, z):
        z = _f32(z)
        out = torch.empty_like(z)
        _check(_L().gemnet_scaled_silu_fwd(_p(z), None, 1.0, _n4(z), 0, _p(out), _stream()), "gemnet_scaled_silu_fwd")
        ctx.save_for_backward(z)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (z,) = ctx.saved_tensors
        g = _f32(g)
        gz = torch.empty_like(z)
        _check(_L().gemnet_scaled_silu_bwd(_p(z), _p(g), 1.0, z.numel(), _p(gz), _stream()), "gemnet_scaled_silu_bwd")
        return gz


def scaled_silu(z):
    """ScaledSiLU (base_layers.py:51-58): silu(z) / 0.6 in one kernel."""
    return _ScaledSiLU.apply(z)


class _ResidualTail(Function):
    @staticmethod
    def forward(ctx, z, x, s):
        z, x = _f32(z), _f32(x)
        out = torch.empty_like(z)
        _check(_L().gemnet_scaled_silu_fwd(_p(z), _p(x), s, _n4(z), 1, _p(out), _stream()), "gemnet_scaled_silu_fwd")
        ctx.save_for_backward(z)
        ctx.s = s
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (z,) = ctx.saved_tensors
        g = _f32(g)
        gz = torch.empty_like(z)
        _check(_L().gemnet_scaled_silu_bwd(_p(z), _p(g), ctx.s, z.numel(), _p(gz), _stream()), "gemnet_scaled_silu_bwd")
        gx = None
        if ctx.needs_input_grad[1]:
            gx = torch.empty_like(z)
            _check(_L().gemnet_add_scale(_p(g), None, None, ctx.s, z.numel(), _p(gx), _stream()), "gemnet_add_scale")
        return gz, gx, None


def residual_tail(z, x, s):
    """(x + silu(z)/0.6) * s — the last activation + skip + 1/sqrt(2) of ResidualLayer (base_layers.py:86-89)."""
    return _ResidualTail.apply(z, x, float(s))


class _AddScale(Function):
    @staticmethod
    def forward(ctx, s, a, b, c):
        a, b, c = _f32(a), _f32(b), _f32(c)
        out = torch.empty_like(a)
        _check(_L().gemnet_add_scale(_p(a), _p(b), _p(c), s, _n4(a), _p(out), _stream()), "gemnet_add_scale")
        ctx.s = s
        ctx.has = (b is not None, c is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _f32(g)
        gs = torch.empty_like(g)
        _check(_L().gemnet_add_scale(_p(g), None, None, ctx.s, g.numel(), _p(gs), _stream()), "gemnet_add_scale")
        return None, gs, gs if ctx.has[0] else None, gs if ctx.has[1] else None


def lincomb(terms):
    """sum_i s_i * t_i over a list of (tensor, scale) with equal shapes (fp32, contiguous): ceil((n-1)/2) launches."""
    terms = [(_f32(t), float(s)) for t, s in terms]
    while True:
        (a, sa), rest = terms[0], terms[1:3]
        (b, sb) = rest[0] if len(rest) > 0 else (None, 0.0)
        (c, sc) = rest[1] if len(rest) > 1 else (None, 0.0)
        out = torch.empty_like(a)
        _check(_L().gemnet_lincomb3(_p(a), sa, _p(b), sb, _p(c), sc, a.numel(), _p(out), _stream()), "gemnet_lincomb3")
        terms = [(out, 1.0)] + terms[3:]
        if len(terms) == 1:
            return out


def add_scale(s, a, b=None, c=None):
    """(a [+ b [+ c]]) * s — skip connections with their 1/sqrt(2), 1/sqrt(3) (interaction_block.py:202-233)."""
    return _AddScale.apply(float(s), a, b, c)


class _MulScale(Function):
    @staticmethod
    def forward(ctx, a, b, s):
        a, b = _f32(a), _f32(b)
        out = torch.empty_like(a)
        _check(_L().gemnet_mul_scale_fwd(_p(a), _p(b), s, _n4(a), _p(out), _stream()), "gemnet_mul_scale_fwd")
        ctx.save_for_backward(a, b)
        ctx.s = s
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = _f32(g)
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        if ga is not None or gb is not None:
            _check(_L().gemnet_mul_scale_bwd(_p(a), _p(b), _p(g), ctx.s, a.numel(), _p(ga), _p(gb), _stream()),
                   "gemnet_mul_scale_bwd")
        return ga, gb, None


def mul_scale(a, b, s=1.0):
    """a * b * s — Hadamard product with its ScalingFactor (interaction_block.py:661-664; scaling.py:170-174)."""
    return _MulScale.apply(a, b, float(s))


class _GatherRows(Function):
    @staticmethod
    def forward(ctx, x, idx, perm, rowptr):
        x = _f32(x)
        n, F = idx.numel(), x.shape[1]
        out = torch.empty(n, F, dtype=torch.float32, device=x.device)
        _check(_L().gemnet_gather_rows(_p(x), _p(_i32(idx)), n, F, _p(out), _stream()), "gemnet_gather_rows")
        ctx.save_for_backward(perm, rowptr)
        ctx.shape = (x.shape[0], F)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        perm, rowptr = ctx.saved_tensors
        g = _f32(g)
        n_rows, F = ctx.shape
        gx = torch.empty(n_rows, F, dtype=torch.float32, device=g.device)
        _check(_L().gemnet_segment_sum_rows(_p(g), _p(_i32(perm)), _p(_i32(rowptr)), n_rows, F, _p(gx), _stream()),
               "gemnet_segment_sum_rows")
        return gx, None, None, None


def gather_rows(x, idx, perm, rowptr):
    """x[idx] with a CSR (perm, rowptr = csr_by_key(idx, x.shape[0])) so that the backward is a deterministic
    segment sum instead of ATen's sort + atomics index_put (embedding_block.py:70-71, interaction_block.py:693)."""
    return _GatherRows.apply(x, idx, perm, rowptr)


# ----------------------------------------------------------------------------- helpers of the fused engine (fused.py)
def segment_sum_rows(src, perm, rowptr, n_rows):
    """out[r,:] = sum_{k in [rowptr[r], rowptr[r+1])} src[perm[k],:] — the backward of a row gather, no atomics."""
    src = _f32(src)
    F = src.shape[1]
    out = torch.empty(n_rows, F, dtype=torch.float32, device=src.device)
    _check(_L().gemnet_segment_sum_rows(_p(src), _p(_i32(perm)), _p(_i32(rowptr)), n_rows, F, _p(out), _stream()),
           "gemnet_segment_sum_rows")
    return out


def edge_to_atom_grad_r(m, id_a, g_out, s_out, g_had, s_had):
    """g_r[e,:] = (s_out * g_out[id_a[e],:] + s_had * g_had[e,:]) * m[e,:]  — the rbf-side gradient of the edge->atom
    aggregation (atom_update_block.py:64-68) on its own (the edge side stays lazy, fused.EdgeToAtomLazy)."""
    m = _f32(m)
    E, F = m.shape
    g_r = torch.empty_like(m)
    _check(_L().gemnet_edge_to_atom_bwd(_p(m), _p(m), _p(_i32(id_a)), E, F, float(s_out), _p(_f32(g_out)),
                                        float(s_had), _p(_f32(g_had)), None, _p(g_r), _stream()),
           "gemnet_edge_to_atom_bwd")
    return g_r


def materialize_term(g, idx, mul, scale):
    """scale * g[idx] * mul as a tensor (idx / mul optional): a lazy gradient summand for a consumer that is not a
    fused chain."""
    g = _f32(g)
    if idx is not None and mul is not None:
        mul = _f32(mul)
        out = torch.empty_like(mul)
        _check(_L().gemnet_edge_to_atom_bwd(_p(mul), _p(mul), _p(_i32(idx)), mul.shape[0], mul.shape[1], float(scale),
                                            _p(g), 0.0, None, _p(out), None, _stream()), "gemnet_edge_to_atom_bwd")
        return out
    if mul is not None:
        mul = _f32(mul)
        out = torch.empty_like(mul)
        _check(_L().gemnet_mul_scale_fwd(_p(g), _p(mul), float(scale), _n4(g), _p(out), _stream()), "gemnet_mul_scale_fwd")
        return out
    if idx is not None:
        n, F = idx.numel(), g.shape[1]
        out = torch.empty(n, F, dtype=torch.float32, device=g.device)
        _check(_L().gemnet_gather_rows(_p(g), _p(_i32(idx)), n, F, _p(out), _stream()), "gemnet_gather_rows")
        return out if scale == 1.0 else lincomb([(out, scale)])
    return g if scale == 1.0 else lincomb([(g, scale)])


def _ptr_array(tensors):
    import ctypes
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def rbf_fan_fwd(rbf, W1s, W2s, groups):
    """out (J, E, H): out[j] = (rbf W1s[groups[j]]^T) W2s[j]^T  (gemnet_rbf_fan_fwd)."""
    import ctypes
    rbf = _f32(rbf)
    E, R = rbf.shape
    J, H = len(W2s), W2s[0].shape[0]
    W1s, W2s = [_f32(w) for w in W1s], [_f32(w) for w in W2s]
    out = torch.empty(J, E, H, dtype=torch.float32, device=rbf.device)
    _check(_L().gemnet_rbf_fan_fwd(_p(rbf), E, R, _ptr_array(W1s), len(W1s), _ptr_array(W2s),
                                   (ctypes.c_int32 * J)(*groups), J, H, _p(out), _stream()), "gemnet_rbf_fan_fwd")
    return out


def rbf_fan_bwd(E, R, W1s, W2s, groups, g_outs):
    """g_rbf (E, R) = sum_j (g_outs[j] W2s[j]) W1s[groups[j]]; g_outs[j] may be None (gemnet_rbf_fan_bwd)."""
    import ctypes
    J, H = len(W2s), W2s[0].shape[0]
    W1s, W2s = [_f32(w) for w in W1s], [_f32(w) for w in W2s]
    g_outs = [_f32(g) for g in g_outs]
    dev = W2s[0].device
    ws = torch.empty(J * E * 16, dtype=torch.float32, device=dev)
    g_rbf = torch.empty(E, R, dtype=torch.float32, device=dev)
    _check(_L().gemnet_rbf_fan_bwd(E, R, _ptr_array(W1s), len(W1s), _ptr_array(W2s), (ctypes.c_int32 * J)(*groups), J, H,
                                   _ptr_array(g_outs), _p(ws), _p(g_rbf), _stream()), "gemnet_rbf_fan_bwd", kernels=2)
    return g_rbf


def weight_grad(gz, a):
    """dW (N, K) = gz^T (N x rows) @ a (rows x K): the weight gradient of a dense layer from the two operands the chain
    kernels saved (training only).  A plain GEMM with the row count as reduction dimension: library call."""
    return torch.mm(gz.t(), a)
