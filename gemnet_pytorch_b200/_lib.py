"""ctypes binding of the C-ABI library (include/gemnet_b200.h).  No CPU fallback: importing the ops
without the built library, or calling them without a CUDA device, raises."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libgemnet_b200.so")

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I32 = ctypes.c_int32
_F = ctypes.c_float
_D = ctypes.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "gemnet_abi_version": [],
    "gemnet_last_error": [],
    "gemnet_csr_from_sorted": [_P, _I64, _I64, _P, _P, _P],
    "gemnet_csr_by_key": [_P, _I64, _I64, _P, _P, _P, _P, _P],
    "gemnet_index_range_check": [_P, _P, _P, _I32, _P, _P],
    "gemnet_edge_geometry_fwd": [_P, _P, _P, _I64, _P, _P, _P],
    "gemnet_edge_geometry_bwd": [_P, _P, _P, _P, _I64, _P, _P, _P, _P],
    "gemnet_bessel_rbf_fwd": [_P, _P, _I64, _I32, _F, _I32, _P, _P],
    "gemnet_bessel_rbf_bwd": [_P, _P, _I64, _I32, _F, _I32, _P, _P, _P, _P],
    "gemnet_sph_bessel_radial_fwd": [_P, _I64, _I32, _I32, _P, _P, _F, _I32, _P, _P],
    "gemnet_sph_bessel_radial_bwd": [_P, _I64, _I32, _I32, _P, _P, _F, _I32, _P, _P, _P],
    "gemnet_triplet_sph_fwd": [_P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P],
    "gemnet_triplet_sph_bwd": [_P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P],
    "gemnet_quad_intm_fwd": [_P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P],
    "gemnet_quad_intm_bwd": [_P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P, _P],
    "gemnet_quad_ylm_fwd": [_P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P],
    "gemnet_quad_ylm_bwd": [_P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P, _P, _P],
    "gemnet_downproj_fwd": [_P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P],
    "gemnet_downproj_bwd": [_P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P, _P],
    "gemnet_segbilinear_workspace_bytes": [_I64, _I32, _I32, _I32, _I32, _I32, _I32],
    "gemnet_segbilinear_pack_bytes": [_I32, _I32, _I32, _I32, _I32],
    "gemnet_segbilinear_pack": [_P, _I32, _I32, _I32, _I32, _I32, _P, _P],
    "gemnet_segbilinear_segsum": [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P],
    "gemnet_segbilinear_fwd": [_P, _P, _P, _P, _P, _P, _F, _I64, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P, _P],
    "gemnet_segbilinear_bwd": [_P, _P, _P, _P, _P, _P, _F, _I64, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P, _P,
                               _P, _P, _P, _P],
    "gemnet_edge_to_atom_fwd": [_P, _P, _P, _P, _I64, _I64, _I32, _F, _P, _F, _P, _P],
    "gemnet_edge_to_atom_bwd": [_P, _P, _P, _I64, _I32, _F, _P, _F, _P, _P, _P, _P],
    "gemnet_molecule_sum_fwd": [_P, _P, _I64, _I32, _I32, _P, _P],
    "gemnet_molecule_sum_bwd": [_P, _P, _P, _I64, _I32, _I32, _P, _P],
    "gemnet_force_readout_fwd": [_P, _P, _P, _P, _I64, _I32, _P, _P],
    "gemnet_force_readout_bwd": [_P, _P, _P, _I64, _I32, _P, _P, _P, _P],
    "gemnet_scaled_silu_fwd": [_P, _P, _F, _I64, _I32, _P, _P],
    "gemnet_scaled_silu_bwd": [_P, _P, _F, _I64, _P, _P],
    "gemnet_add_scale": [_P, _P, _P, _F, _I64, _P, _P],
    "gemnet_lincomb3": [_P, _F, _P, _F, _P, _F, _I64, _P, _P],
    "gemnet_mul_scale_fwd": [_P, _P, _F, _I64, _P, _P],
    "gemnet_mul_scale_bwd": [_P, _P, _P, _F, _I64, _P, _P, _P],
    "gemnet_gather_rows": [_P, _P, _I64, _I32, _P, _P],
    "gemnet_segment_sum_rows": [_P, _P, _P, _I64, _I32, _P, _P],
    "gemnet_dense_pack_floats": [_I32, _I32],
    "gemnet_dense_pack_weight": [_P, _I64, _I64, _I32, _I32, _P, _P],
    "gemnet_dense_pack_many": [_P, _I32, _P],
    "gemnet_dense_tc": [_P, _P, _I64, _I32, _I32, _P, _I32, _I32, _F, _P, _F, _P, _P, _I32, _P],
    "gemnet_rbf_fan_fwd": [_P, _I64, _I32, _P, _I32, _P, _P, _I32, _I32, _P, _P],
    "gemnet_rbf_fan_bwd": [_I64, _I32, _P, _I32, _P, _P, _I32, _I32, _P, _P, _P, _P],
    "gemnet_chain_tc": [_P, _I32, _P],
    "gemnet_chain_set_profile_buffer": [_P],
    "gemnet_chain_set_tile_rows": [_I32],
    "gemnet_graph_build_host": [_P, _P, _I64, _D, _D, _I32, ctypes.POINTER(_P)],
    "gemnet_graph_size": [_P, ctypes.c_char_p],
    "gemnet_graph_copy_host": [_P, ctypes.c_char_p, _P],
    "gemnet_graph_copy_host_i64": [_P, ctypes.c_char_p, _P],
    "gemnet_graph_free": [_P],
    "gemnet_pad_batch_host": [_P] * 10 + [_I64] * 8 + [_P] * 10,
    "gemnet_graph_device_count": [_P, _P],
    "gemnet_graph_device_edges": [_P, _P],
    "gemnet_graph_device_quads": [_P, _P],
    "gemnet_graph_device_pad": [_P, _P, _P, _P],
    "gemnet_graph_count_host": [_P, _P, _I64, _D, _P, _P],
}
_RESTYPES = {
    "gemnet_last_error": ctypes.c_char_p,
    "gemnet_segbilinear_workspace_bytes": _I64,
    "gemnet_segbilinear_pack_bytes": _I64,
    "gemnet_dense_pack_floats": _I64,
    "gemnet_graph_size": _I64,
    "gemnet_graph_free": None,
    "gemnet_chain_set_profile_buffer": None,
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class GemnetB200Error(RuntimeError):
    pass


def load():
    """Load libgemnet_b200.so (built in-tree by gemnet_pytorch_b200._build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GemnetB200Error(
            f"{LIB_PATH} not found: build it with `python -m gemnet_pytorch_b200._build` "
            "(there is no CPU / PyTorch fallback for the GemNet hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    if lib.gemnet_abi_version() != 1:
        raise GemnetB200Error("libgemnet_b200.so ABI version mismatch")
    _lib = lib
    return lib


# number of CUDA kernels launched through the C ABI since import (bench.py reports the per-step count)
KERNEL_LAUNCHES = [0]


def check(rc, what, kernels=1):
    KERNEL_LAUNCHES[0] += kernels
    if rc != 0:
        msg = load().gemnet_last_error()
        raise GemnetB200Error(f"{what} failed: {msg.decode() if msg else 'unknown error'}")
