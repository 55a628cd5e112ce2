"""ctypes binding of libvittf_b200.so (C ABI declared in include/vittf.h).

There is NO fallback: if the shared library is missing, or a call fails, a
``VittfError`` is raised.  The product path never routes through PyTorch eager ops
or the CPU oracle for the hot kernels.
"""
import ctypes as C
import os
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("VITTF_LIB", _HERE / "libvittf_b200.so"))


class VittfError(RuntimeError):
    pass


class VitConfig(C.Structure):
    _fields_ = [("embed_dim", C.c_int), ("depth", C.c_int), ("num_heads", C.c_int), ("patch", C.c_int),
                ("mlp_hidden", C.c_int)]


class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b",
                                          "fc1_w", "fc1_b", "fc2_w", "fc2_b", "qkv_colsum", "fc1_colsum")]


class LnFold(C.Structure):
    """vittf_ln_fold (include/vittf.h): LayerNorm folded into the GEMM epilogues."""
    _fields_ = [("colsum", C.c_void_p), ("stats", C.c_void_p), ("eps", C.c_float),
                ("xt", C.c_void_p), ("stats_out", C.c_void_p), ("m_pad", C.c_int64)]


class BlsParams(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("D", C.c_int), ("sigma_spatial", C.c_double), ("lam", C.c_double),
                ("A_diag_min", C.c_double), ("cg_tol", C.c_double), ("cg_maxiter", C.c_int), ("luma_bins", C.c_int)]


U8, F16, BF16, F32, F64 = 0, 1, 2, 3, 4
DTYPE_CODE = {torch.uint8: U8, torch.float16: F16, torch.bfloat16: BF16, torch.float32: F32, torch.float64: F64}
EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32, EPI_QKV_SPLIT, EPI_KFEAT_F16, EPI_BIAS_RESID_LN = range(6)
LN_SLOTS = 8            # VITTF_LN_SLOTS
SIM_NS, SIM_REFNTF, SIM_LEGACY, SIM_CLAMP_MEAN = 0, 1, 2, 3

_p, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
_SIGNATURES = {
    "vittf_last_error": (C.c_char_p, []),
    "vittf_version": (_i, []),
    "vittf_device_arch": (_i, [C.POINTER(_i)]),
    "vittf_launch_count": (_i64, []),
    "vittf_launch_count_reset": (None, []),
    "vittf_vit_timing_enable": (_i, [_p, _i]),
    "vittf_vit_timing_read": (_i, [_p, C.POINTER(_d), C.POINTER(_i64)]),
    "vittf_minmax": (_i, [_p, _i64, _i, _p, _p]),
    "vittf_vit_create": (_i, [C.POINTER(_p), C.POINTER(VitConfig), C.POINTER(BlockWeights), _p, _p, _i, _i]),
    "vittf_vit_destroy": (None, [_p]),
    "vittf_vit_workspace_bytes": (_i64, [_p, _i, _i]),
    "vittf_vit_k_features": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _i64, _p]),
    "vittf_pool_axis": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    "vittf_accumulate_gathered_f16": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "vittf_accumulate_f16": (_i, [_p, _p, _i64, _p]),
    "vittf_gemm_bf16": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "vittf_attention": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "vittf_attention_workspace_bytes": (_i64, [_i, _i, _i]),
    "vittf_attention_prescaled": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _i64, _p]),
    "vittf_layernorm": (_i, [_p, _p, _p, _p, _i64, _i, _p]),
    "vittf_gemm_ln_slots": (_i, [_i]),
    "vittf_gemm_bf16_ln": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "vittf_ln_prepare": (_i, [_p, _p, _p, _p, _i64, _i64, _i, _p]),
    "vittf_patch_embed": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "vittf_sample_prototypes": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _i, _p, _p]),
    "vittf_sim_lowres_layout": (_i, [_i, _i, _i, _i, _i, _p]),
    "vittf_sim_lowres": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _i, _i, _p]),
    "vittf_sim_upsample": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _i, _p, _p]),
    "vittf_class_max": (_i, [_p, _i, _i64, _p, _p]),
    "vittf_quantize_maps_u8": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _i, _i, _i, _p, _p]),
    "vittf_labels": (_i, [_p, _i, _i, _i64, _p, _i, _p, _p]),
    "vittf_bls_workspace_bytes": (_i64, [C.POINTER(BlsParams), _i]),
    "vittf_bls_solve": (_i, [C.POINTER(BlsParams), _p, _p, _p, _p, _i, _p, _p, _p, _i64, _p]),
    "vittf_sobel_confidence": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "vittf_binary_erosion": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "vittf_topk_voxels": (_i, [_p, _i, _i64, _i, _p, _p, _p]),
    "vittf_mean_pairwise_distance": (_i, [_p, _i, _i, _i, _p, _p]),
    "vittf_confusion_matrix": (_i, [_p, _p, _i64, _i, _p, _p, _p]),
    "vittf_bls_grid_cells": (_i64, [C.POINTER(BlsParams)]),
    "vittf_bls_grid_workspace_bytes": (_i64, [C.POINTER(BlsParams), _i]),
    "vittf_bls_sobel_slab": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "vittf_bls_splat_slab": (_i, [C.POINTER(BlsParams), _p, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "vittf_bls_grid_solve": (_i, [C.POINTER(BlsParams), _i, _p, _p, _p, _p, _i64, _p]),
    "vittf_bls_slice_slab": (_i, [C.POINTER(BlsParams), _p, _p, _p, _i, _i, _i, _p, _p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Loads the shared library once; raises VittfError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VittfError(f"{LIB_PATH} not found: build it with vittf_b200/csrc/build.sh "
                         "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().vittf_last_error().decode(errors="replace")
        raise VittfError(f"{what} failed with status {status}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VittfError("vittf_b200 kernels need CUDA tensors; there is no CPU fallback")
        if t is not None and not t.is_contiguous():
            raise VittfError("vittf_b200 kernels need contiguous tensors")
