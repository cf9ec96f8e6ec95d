"""ctypes binding of libfpv_b200.so (the C ABI declared in include/fpv_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfpv_b200.so")


class SmplxModelStruct(ctypes.Structure):
    """Mirror of fpv_smplx_model_t."""
    _fields_ = [
        ("num_verts", c_int32), ("num_extra", c_int32), ("ell_width", c_int32), ("reserved", c_int32),
        ("basis_kn", c_void_p),
        ("basis_nk_hi", c_void_p), ("basis_nk_lo", c_void_p),
        ("basis_kn_hi", c_void_p), ("basis_kn_lo", c_void_p),
        ("j_template", c_void_p), ("j_shapedirs", c_void_p), ("parents", c_void_p),
        ("hand_comps", c_void_p), ("pose_mean", c_void_p),
        ("ell_joint", c_void_p), ("ell_weight", c_void_p),
        ("csr_ptr", c_void_p), ("csr_vert", c_void_p), ("csr_weight", c_void_p),
        ("extra_vertex_ids", c_void_p),
    ]


class VposerModelStruct(ctypes.Structure):
    """Mirror of fpv_vposer_model."""
    _fields_ = [
        ("w1", c_void_p), ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p), ("w3", c_void_p), ("b3", c_void_p),
        ("w1t", c_void_p), ("w2t", c_void_p), ("w3t", c_void_p),
        ("latent", c_int), ("hidden", c_int), ("joints", c_int),
    ]


# name -> (restype, argtypes); every symbol include/fpv_b200.h declares
SIGNATURES = {
    "fpv_last_error": (c_char_p, []),
    "fpv_abi_version": (c_int, []),
    "fpv_device_query": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "fpv_launch_count": (ctypes.c_ulonglong, []),
    "fpv_profile_enable": (c_int, [c_int]),
    "fpv_profile_count": (c_int, []),
    "fpv_profile_get": (c_int, [c_int, ctypes.c_char_p, POINTER(c_float), POINTER(ctypes.c_double),
                                POINTER(ctypes.c_double)]),
    "fpv_fp32_probe": (c_int, [POINTER(ctypes.c_double), c_void_p]),
    "fpv_nn_planes_bytes": (c_size_t, [c_int64, c_int64]),
    "fpv_nn_pack_planes": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "fpv_nn_search_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "fpv_nn_search": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64,
                              c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fpv_nn_unpack_keys": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "fpv_nn_set_tuning": (c_int, [c_int, c_int, c_int]),
    "fpv_nn_set_engine": (c_int, [c_int, c_int]),
    "fpv_nn_sphere_set_chunking": (c_int, [c_int]),
    "fpv_nn_tc_debug": (c_int, [c_void_p]),
    "fpv_nn_culled_tile": (c_int, [c_int]),
    "fpv_nn_tile_boxes_floats": (c_size_t, [c_int64, c_int]),
    "fpv_nn_tile_boxes": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "fpv_nn_culled_search": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                     c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                     c_void_p]),
    "fpv_nn_culled_search_keys": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                          c_int64, c_void_p, POINTER(c_void_p), c_int, c_void_p, c_int64, c_void_p,
                                          c_void_p, c_void_p, c_int, c_void_p]),
    "fpv_nn_sphere_fused_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "fpv_fix_shift_for": (c_int, [c_float, c_int64]),
    "fpv_nn_sphere_fused": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                    c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                    c_void_p]),
    "fpv_scene2body_grad": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "fpv_adam_tick": (c_int, [c_void_p, c_void_p]),
    "fpv_adam_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                                c_void_p, c_void_p]),
    "fpv_p2p_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "fpv_p2p_free": (c_int, [c_void_p]),
    "fpv_p2p_export": (c_int, [c_void_p, POINTER(ctypes.c_ubyte)]),
    "fpv_p2p_open": (c_int, [POINTER(ctypes.c_ubyte), POINTER(c_void_p)]),
    "fpv_p2p_close": (c_int, [c_void_p]),
    "fpv_p2p_barrier": (c_int, [POINTER(c_void_p), c_int, c_int, c_void_p, c_void_p, ctypes.c_double, c_void_p]),
    "fpv_p2p_min_unpack": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int,
                                   c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "fpv_p2p_push": (c_int, [c_void_p, c_int64, POINTER(c_void_p), c_int, c_void_p, c_int64, c_void_p]),
    "fpv_p2p_gather": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, POINTER(c_int64), POINTER(c_int64), c_void_p,
                               c_void_p]),
    "fpv_p2p_sum": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "fpv_morton_keys": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fpv_nn_gather_pack": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fpv_nn_sphere_table_floats": (c_size_t, [c_int64, c_int]),
    "fpv_nn_sphere_table": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "fpv_nn_sphere_search": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_void_p, c_int, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p,
                                     c_void_p]),
    "fpv_chamfer_fwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "fpv_chamfer_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "fpv_chamfer_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int, c_int]),
    "fpv_chamfer_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fpv_chamfer_bwd_bcast": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int,
                                      c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fpv_reduce_workspace_bytes": (c_size_t, [c_int64]),
    "fpv_robust_mean_fwd": (c_int, [c_void_p, c_int64, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fpv_robust_mean_bwd": (c_int, [c_void_p, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "fpv_tdiff_l1_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fpv_tdiff_l1_bwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fpv_transform_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "fpv_transform_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "fpv_transform_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "fpv_split_tf32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "fpv_tc_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "fpv_tc_gemm_3xtf32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                   c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "fpv_smplx_saved_bytes": (c_size_t, [POINTER(SmplxModelStruct), c_int64]),
    "fpv_smplx_workspace_bytes": (c_size_t, [POINTER(SmplxModelStruct), c_int64]),
    "fpv_smplx_fwd": (c_int, [POINTER(SmplxModelStruct), c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "fpv_smplx_bwd": (c_int, [POINTER(SmplxModelStruct), c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_size_t, c_void_p]),
    "fpv_rot6d_to_aa_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "fpv_rot6d_to_aa_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "fpv_aa_to_rot6d": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "fpv_vposer_saved_floats": (c_size_t, [POINTER(VposerModelStruct), c_int64]),
    "fpv_vposer_decode_fwd": (c_int, [POINTER(VposerModelStruct), c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "fpv_vposer_decode_bwd": (c_int, [POINTER(VposerModelStruct), c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "fpv_dct_prior_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "fpv_dct_prior_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "fpv_dct_prior_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python 4dcapture-fpv_b200/build.py` "
                "(or __graft_entry__.build()). This package has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if L.fpv_abi_version() != 1:
            raise RuntimeError("libfpv_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().fpv_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what or 'fpv call'} failed (status {rc}): {msg}")


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors) -> None:
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("4dcapture-fpv_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("expected a CUDA tensor (this package has no CPU path); got device " + str(t.device))


def workspace(nbytes: int, device):
    import torch
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
