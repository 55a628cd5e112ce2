"""Tensor-level wrappers around the C ABI (one function per exported kernel entry point).

PyTorch is used for device memory and streams only; every computation below happens in
libvittf_b200.so.  All functions raise ``VittfError`` on failure -- no fallbacks.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import DTYPE_CODE, check, load, ptr, require_cuda, stream_ptr

AXIS_INDEX = {"x": 0, "y": 1, "z": 2}


def _on_device(fn):
    """Runs the wrapped entry point with the CUDA device of its first CUDA tensor argument current: kernels launch on
    the current device, so a tensor on another GPU (e.g. compute_qkv(dev='cuda:1')) must switch to it first."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def tok_pad_of(tokens):
    return (tokens + 127) // 128 * 128


def device_arch():
    out = C.c_int(0)
    check(load().vittf_device_arch(C.byref(out)), "vittf_device_arch")
    return out.value


@_on_device
def minmax(vol):
    require_cuda(vol)
    if vol.data_ptr() % 16:                 # a contiguous view with a storage offset (vol[1:]): the kernel reads 16-byte words
        vol = vol.clone()
    out = torch.empty(2, dtype=torch.float32, device=vol.device)
    check(load().vittf_minmax(ptr(vol), vol.numel(), DTYPE_CODE[vol.dtype], ptr(out), stream_ptr(vol.device)), "vittf_minmax")
    return out


@_on_device
def gemm_bf16(a, w, bias, epi, out=None, out2=None, tokens=0, tok_pad=0):
    """a (M,K) bf16, w (N,K) bf16, bias (N) fp32; see include/vittf.h for the epilogues."""
    require_cuda(a, w, bias, out, out2)
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        if epi in (_lib.EPI_BIAS_BF16, _lib.EPI_BIAS_GELU_BF16):
            out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
        elif epi == _lib.EPI_QKV_SPLIT:
            d = N // 3
            out = torch.empty(M, 2 * d, dtype=torch.bfloat16, device=a.device)
            out2 = torch.zeros((M // tokens) * d, tok_pad, dtype=torch.bfloat16, device=a.device)
        elif epi == _lib.EPI_KFEAT_F16:
            out = torch.empty((M // tokens) * (tokens - 1), N, dtype=torch.float16, device=a.device)
        else:
            raise _lib.VittfError("residual epilogue needs the fp32 stream passed as out=")
    check(load().vittf_gemm_bf16(ptr(a), ptr(w), ptr(bias), ptr(out), ptr(out2), M, N, K, epi, tokens, tok_pad,
                                 stream_ptr(a.device)), "vittf_gemm_bf16")
    return (out, out2) if epi == _lib.EPI_QKV_SPLIT else out


def ln_slots(n):
    """Partial-sum slots the residual-stream epilogue (EPI_BIAS_RESID_LN) writes per row for an N-column stream."""
    return int(load().vittf_gemm_ln_slots(int(n)))


def m_pad_of(rows):
    return (rows + 255) // 256 * 256


def xt_to_rows(xt, rows, D):
    """Row-tiled stream xt[m_pad/32][D/4][32][4] -> row-major (rows, D) (test / debugging helper, torch glue)."""
    return xt.view(-1, D // 4, 32, 4).permute(0, 2, 1, 3).reshape(-1, D)[:rows]


@_on_device
def ln_prepare(x):
    """Row-major fp32 stream (rows, D) -> (xt row-tiled fp32, xb raw bf16 copy, stats (m_pad, LN_SLOTS, 2))."""
    require_cuda(x)
    rows, D = x.shape
    mp = m_pad_of(rows)
    xt = torch.zeros(mp * D, dtype=torch.float32, device=x.device)
    xb = torch.empty(rows, D, dtype=torch.bfloat16, device=x.device)
    stats = torch.zeros(mp, _lib.LN_SLOTS, 2, dtype=torch.float32, device=x.device)
    check(load().vittf_ln_prepare(ptr(x), ptr(xt), ptr(xb), ptr(stats), rows, mp, D, stream_ptr(x.device)), "vittf_ln_prepare")
    return xt, xb, stats


@_on_device
def gemm_bf16_ln(a, w, bias, epi, colsum=None, stats=None, xt=None, out=None, out2=None, tokens=0, tok_pad=0, eps=1e-6,
                 stats_out=None):
    """vittf_gemm_bf16_ln: consumer (colsum + stats: LayerNorm of the A operand's fp32 source applied in the epilogue) and /
    or producer (epi = EPI_BIAS_RESID_LN: xt += a w^T + bias in the row-tiled stream; returns (xb, stats_out))."""
    require_cuda(a, w, bias, colsum, stats, xt, out, out2)
    M, K = a.shape
    N = w.shape[0]
    mp = m_pad_of(M)
    fold = _lib.LnFold()
    fold.m_pad = mp
    fold.eps = eps
    if colsum is not None:
        fold.colsum = colsum.data_ptr()
        fold.stats = stats.data_ptr()
    if epi == _lib.EPI_BIAS_RESID_LN:
        stats_out = torch.zeros(mp, _lib.LN_SLOTS, 2, dtype=torch.float32, device=a.device) if stats_out is None else stats_out
        fold.xt = xt.data_ptr()
        fold.stats_out = stats_out.data_ptr()
        out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device) if out is None else out
    elif out is None:
        if epi in (_lib.EPI_BIAS_BF16, _lib.EPI_BIAS_GELU_BF16):
            out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
        elif epi == _lib.EPI_QKV_SPLIT:
            d = N // 3
            out = torch.empty(M, 2 * d, dtype=torch.bfloat16, device=a.device)
            out2 = torch.zeros((M // tokens) * d, tok_pad, dtype=torch.bfloat16, device=a.device)
        elif epi == _lib.EPI_KFEAT_F16:
            out = torch.empty((M // tokens) * (tokens - 1), N, dtype=torch.float16, device=a.device)
        else:
            raise _lib.VittfError("vittf_gemm_bf16_ln: unsupported epilogue")
    import ctypes as C
    check(load().vittf_gemm_bf16_ln(ptr(a), ptr(w), ptr(bias), ptr(out), ptr(out2), M, N, K, epi, tokens, tok_pad, C.byref(fold),
                                    stream_ptr(a.device)), "vittf_gemm_bf16_ln")
    if epi == _lib.EPI_BIAS_RESID_LN:
        return out, stats_out
    return (out, out2) if epi == _lib.EPI_QKV_SPLIT else out


@_on_device
def attention(qk, vt, batch, tokens, heads, tok_pad):
    require_cuda(qk, vt)
    out = torch.empty(batch * tokens, heads * 64, dtype=torch.bfloat16, device=qk.device)
    check(load().vittf_attention(ptr(qk), ptr(vt), ptr(out), batch, tokens, tok_pad, heads, stream_ptr(qk.device)),
          "vittf_attention")
    return out


@_on_device
def attention_prescaled(qk, vt, batch, tokens, heads, tok_pad, return_flags=False):
    """softmax(q k^T) v for q already multiplied by hd^-0.5 * log2(e): max-free first pass + safe pass over flagged CTAs."""
    require_cuda(qk, vt)
    out = torch.empty(batch * tokens, heads * 64, dtype=torch.bfloat16, device=qk.device)
    need = load().vittf_attention_workspace_bytes(batch, tokens, heads)
    ws = torch.empty(need // 4, dtype=torch.int32, device=qk.device)
    check(load().vittf_attention_prescaled(ptr(qk), ptr(vt), ptr(out), batch, tokens, tok_pad, heads, ptr(ws), need,
                                           stream_ptr(qk.device)), "vittf_attention_prescaled")
    return (out, ws) if return_flags else out


@_on_device
def layernorm(x, w, b):
    require_cuda(x, w, b)
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    rows = x.numel() // x.shape[-1]
    check(load().vittf_layernorm(ptr(x), ptr(w), ptr(b), ptr(y), rows, x.shape[-1], stream_ptr(x.device)), "vittf_layernorm")
    return y


@_on_device
def patch_embed(vol, axis, s0, s1, im0, im1, patch, mm, patch_w, patch_b, pos):
    require_cuda(vol, mm, patch_w, patch_b, pos)
    D = patch_b.numel()
    tokens = 1 + (im0 // patch) * (im1 // patch)
    out = torch.empty(s1 - s0, tokens, D, dtype=torch.float32, device=vol.device)
    X, Y, Z = vol.shape
    check(load().vittf_patch_embed(ptr(vol), DTYPE_CODE[vol.dtype], X, Y, Z, AXIS_INDEX[axis], s0, s1, im0, im1, patch, D,
                                   ptr(mm), ptr(patch_w), ptr(patch_b), ptr(pos), ptr(out), stream_ptr(vol.device)),
          "vittf_patch_embed")
    return out


@_on_device
def pool_axis(k, f0, f1, axis, n_out, out=None, accumulate=False, total_slices=None, slice0=0, slabs=None, compact=False):
    """k (S, f0*f1, D) fp16 -> (D, ., ., .) fp16 in the reference layout (infer.py:203).
    Sharded use: k holds global slices [slice0, slice0+S) of `total_slices`; only output slabs
    `slabs=(o0, o1)` are written -- into the full-size array, or with compact=True into this rank's
    block whose slab axis has extent o1 - o0 (what the multi-GPU all-gather sends)."""
    require_cuda(k, out)
    n_local, T, D = k.shape
    S = n_local if total_slices is None else total_slices
    o0, o1 = (0, n_out) if slabs is None else slabs
    assert T == f0 * f1
    ext = o1 - o0 if compact else n_out
    shape = {"z": (D, f0, f1, ext), "y": (D, f0, ext, f1), "x": (D, ext, f0, f1)}[axis]
    if out is None:
        assert not accumulate
        out = torch.empty(shape, dtype=torch.float16, device=k.device)
    assert tuple(out.shape) == shape and out.dtype == torch.float16
    check(load().vittf_pool_axis(ptr(k), S, slice0, n_local, f0, f1, D, AXIS_INDEX[axis], n_out, o0, o1, ptr(out),
                                 int(accumulate), int(compact), stream_ptr(k.device)), "vittf_pool_axis")
    return out


@_on_device
def accumulate_gathered(out, staging, axis, accumulate):
    """out fp16 (D,fX,fY,fZ) = (or +=, in fp16) the un-permuted all-gather `staging` (world, D, e0, e1, e2) of one axis."""
    require_cuda(out, staging)
    D, fX, fY, fZ = out.shape
    world = staging.shape[0]
    assert out.dtype == staging.dtype == torch.float16 and staging.numel() == out.numel()
    check(load().vittf_accumulate_gathered_f16(ptr(out), ptr(staging), world, D, fX, fY, fZ, AXIS_INDEX[axis], int(accumulate),
                                               stream_ptr(out.device)), "vittf_accumulate_gathered_f16")
    return out


@_on_device
def accumulate_f16(out, inp):
    """out = fp16(out + inp) in place (infer.py:332)."""
    require_cuda(out, inp)
    assert out.dtype == torch.float16 and inp.dtype == torch.float16 and out.shape == inp.shape
    check(load().vittf_accumulate_f16(ptr(out), ptr(inp), out.numel(), stream_ptr(out.device)), "vittf_accumulate_f16")
    return out


@_on_device
def sample_prototypes(feats, rel, mode):
    require_cuda(feats, rel)
    F, w, h, d = feats.shape
    A = rel.shape[0]
    out = torch.empty(A, F, dtype=torch.float32, device=feats.device)
    check(load().vittf_sample_prototypes(ptr(feats), DTYPE_CODE[feats.dtype], F, w, h, d, ptr(rel), A,
                                         {"nearest": 0, "bilinear": 1}[mode], ptr(out), stream_ptr(feats.device)),
          "vittf_sample_prototypes")
    return out


@_on_device
def sim_lowres(feats, protos, want_gram=True, voxel_major=False, x_planes=None):
    """Pass 1: dots fp32 (A, n_lr) -- or (n_lr, A4), A4 = A rounded up to 4, with voxel_major=True where the fused
    tensor-core pass is available (else the flag is ignored) -- and the 14 Gram planes.  x_planes=(xa, xb) restricts
    the evaluation to those low-res x planes (slab sharding).  Returns (dots, gram, layout)."""
    require_cuda(feats, protos)
    F, w, h, d = feats.shape
    A = protos.shape[0]
    n = w * h * d
    layout = 0
    if voxel_major and want_gram:
        layout = int(load().vittf_sim_lowres_layout(DTYPE_CODE[feats.dtype], F, w, h, d, ptr(feats)))
    dots = torch.empty((n, (A + 3) // 4 * 4) if layout else (A, n), dtype=torch.float32, device=feats.device)
    gram = torch.empty(14, n, dtype=torch.float32, device=feats.device) if want_gram else None
    xa, xb = (0, w) if x_planes is None else x_planes
    check(load().vittf_sim_lowres(ptr(feats), DTYPE_CODE[feats.dtype], F, w, h, d, ptr(protos), A, ptr(dots), ptr(gram), layout,
                                  int(xa), int(xb), stream_ptr(feats.device)), "vittf_sim_lowres")
    return dots, gram, layout


@_on_device
def sim_upsample(dots, gram, lr_shape, class_offsets, out_shape, mode, threshold=0.25, exponent=2.0, z0=0, z1=None, out=None,
                 layout=0, n_protos=None, x0=0, x1=None):
    require_cuda(dots, gram, class_offsets, out)
    w, h, d = lr_shape
    W, H, D = out_shape
    z1 = D if z1 is None else z1
    x1 = W if x1 is None else x1
    C_ = class_offsets.numel() - 1
    A = n_protos if n_protos is not None else (dots.shape[0] if layout == 0 else dots.shape[1])
    if out is None:
        out = torch.empty(C_, x1 - x0, H, z1 - z0, dtype=torch.float32, device=dots.device)
    check(load().vittf_sim_upsample(ptr(dots), ptr(gram), w, h, d, A, ptr(class_offsets), C_, W, H, D, x0, x1, z0, z1,
                                    mode, float(threshold), float(exponent), int(layout), ptr(out), stream_ptr(dots.device)),
          "vittf_sim_upsample")
    return out


@_on_device
def class_max(sims):
    require_cuda(sims)
    C_ = sims.shape[0]
    out = torch.empty(C_, dtype=torch.float32, device=sims.device)
    check(load().vittf_class_max(ptr(sims), C_, sims[0].numel(), ptr(out), stream_ptr(sims.device)), "vittf_class_max")
    return out


def nearest_src(o, n_in, n_out):
    """Source index of F.interpolate(mode='nearest'): min(floor(o * in/out), in - 1) in fp32 like ATen."""
    import numpy as np
    return min(int(np.floor(np.float32(o) * (np.float32(n_in) / np.float32(n_out)))), n_in - 1)


@_on_device
def quantize_maps_u8(sims, cmax, out_shape, depth=None, z0=0):
    """predict_ntf.py:95-100: sims fp32 (C, W, H, zs) = z-slab [z0, z0+zs) of a (W, H, depth) grid, cmax fp32 (C) the
    GLOBAL per-class maxima -> (uint8 (C, Wo, Ho, zo1-zo0), (zo0, zo1)): the wrapped uint8 quantisation, nearest-resized
    to out_shape = (Wo, Ho, Do); [zo0, zo1) are the output planes whose source plane lies in the slab."""
    require_cuda(sims, cmax)
    C_, W, H, zs = sims.shape
    D = zs if depth is None else depth
    Wo, Ho, Do = out_shape
    planes = [o for o in range(Do) if z0 <= nearest_src(o, D, Do) < z0 + zs]
    zo0, zo1 = (planes[0], planes[-1] + 1) if planes else (0, 0)
    out = torch.empty(C_, Wo, Ho, zo1 - zo0, dtype=torch.uint8, device=sims.device)
    check(load().vittf_quantize_maps_u8(ptr(sims), C_, W, H, D, z0, z0 + zs, ptr(cmax), Wo, Ho, Do, zo0, zo1, ptr(out),
                                        stream_ptr(sims.device)), "vittf_quantize_maps_u8")
    return out, (zo0, zo1)


@_on_device
def labels(sims, thresholds_u8=None, mode=0):
    require_cuda(sims, thresholds_u8)
    C_ = sims.shape[0]
    out = torch.empty(sims.shape[1:], dtype=torch.uint8, device=sims.device)
    check(load().vittf_labels(ptr(sims), DTYPE_CODE[sims.dtype], C_, out.numel(), ptr(thresholds_u8), mode, ptr(out),
                              stream_ptr(sims.device)), "vittf_labels")
    return out


@_on_device
def sobel_confidence(r_u8):
    require_cuda(r_u8)
    W, H, D = r_u8.shape
    out = torch.empty(W, H, D, dtype=torch.float32, device=r_u8.device)
    scratch = torch.empty(1, dtype=torch.float32, device=r_u8.device)
    check(load().vittf_sobel_confidence(ptr(r_u8), W, H, D, ptr(out), ptr(scratch), stream_ptr(r_u8.device)),
          "vittf_sobel_confidence")
    return out


@_on_device
def bls_solve(t, r_u8, conf, luma_lut, sigma_spatial, lam, diag_min, cg_tol, cg_maxiter, luma_bins):
    """t (nrhs,W,H,D) fp32, r_u8 (W,H,D) uint8, conf (W,H,D) fp32 or None -> (out fp32 (nrhs,W,H,D), iters int32)."""
    require_cuda(t, r_u8, conf, luma_lut)
    nrhs, W, H, D = t.shape
    prm = _lib.BlsParams(W, H, D, float(sigma_spatial), float(lam), float(diag_min), float(cg_tol), int(cg_maxiter),
                         int(luma_bins))
    need = load().vittf_bls_workspace_bytes(C.byref(prm), nrhs)
    if need < 0:
        raise _lib.VittfError("vittf_bls_workspace_bytes: bad parameters")
    ws = torch.empty(need, dtype=torch.uint8, device=t.device)
    out = torch.empty_like(t)
    iters = torch.zeros(nrhs, dtype=torch.int32, device=t.device)
    check(load().vittf_bls_solve(C.byref(prm), ptr(t), ptr(r_u8), ptr(conf), ptr(luma_lut), nrhs, ptr(out), ptr(iters),
                                 ptr(ws), need, stream_ptr(t.device)), "vittf_bls_solve")
    return out, iters


@_on_device
def bls_solve_sharded(t_slab, r_u8, conf_slab, luma_lut, sigma_spatial, lam, diag_min, cg_tol, cg_maxiter, luma_bins, z0, z1,
                      all_reduce_max=None, all_reduce_sum=None):
    """The solver in stages over the z-slab [z0, z1) of this rank: t_slab (nrhs,W,H,z1-z0) fp32,
    r_u8 the FULL (W,H,D) reference, conf_slab (W,H,z1-z0) or None (Sobel).  `all_reduce_max(tensor)` /
    `all_reduce_sum(tensor)` are the two exchanges (in place; None = single rank).  The grid stage (bistochastisation
    + PCG) runs replicated.  Returns (out slab fp32 (nrhs,W,H,z1-z0), iters int32 (nrhs))."""
    require_cuda(t_slab, r_u8, conf_slab, luma_lut)
    nrhs = t_slab.shape[0]
    W, H, D = r_u8.shape
    zs = z1 - z0
    if tuple(t_slab.shape[1:]) != (W, H, zs):
        raise ValueError(f"t_slab must be (nrhs,{W},{H},{zs}), got {tuple(t_slab.shape)}")
    dev = t_slab.device
    prm = _lib.BlsParams(W, H, D, float(sigma_spatial), float(lam), float(diag_min), float(cg_tol), int(cg_maxiter),
                         int(luma_bins))
    lib = load()
    ncell = lib.vittf_bls_grid_cells(C.byref(prm))
    if ncell < 0:
        raise _lib.VittfError("vittf_bls_grid_cells: bad parameters")
    st = stream_ptr(dev)
    acc = torch.zeros((2 + nrhs) * ncell, dtype=torch.float64, device=dev)
    if conf_slab is None:
        c_raw = torch.empty(W, H, zs, dtype=torch.float32, device=dev)
        c_max = torch.zeros(1, dtype=torch.float32, device=dev)
        check(lib.vittf_bls_sobel_slab(ptr(r_u8), W, H, D, z0, z1, ptr(c_raw), ptr(c_max), st), "vittf_bls_sobel_slab")
        if all_reduce_max is not None:
            all_reduce_max(c_max)
        check(lib.vittf_bls_splat_slab(C.byref(prm), ptr(t_slab), ptr(r_u8), ptr(c_raw), ptr(c_max), ptr(luma_lut), nrhs, z0, z1,
                                       ptr(acc), st), "vittf_bls_splat_slab")
    else:
        check(lib.vittf_bls_splat_slab(C.byref(prm), ptr(t_slab), ptr(r_u8), ptr(conf_slab), None, ptr(luma_lut), nrhs, z0, z1,
                                       ptr(acc), st), "vittf_bls_splat_slab")
    if all_reduce_sum is not None:
        all_reduce_sum(acc)
    y = torch.empty(nrhs * ncell, dtype=torch.float64, device=dev)
    iters = torch.zeros(nrhs, dtype=torch.int32, device=dev)
    need = lib.vittf_bls_grid_workspace_bytes(C.byref(prm), nrhs)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    check(lib.vittf_bls_grid_solve(C.byref(prm), nrhs, ptr(acc), ptr(y), ptr(iters), ptr(ws), need, st), "vittf_bls_grid_solve")
    out = torch.empty(nrhs, W, H, zs, dtype=torch.float32, device=dev)
    check(lib.vittf_bls_slice_slab(C.byref(prm), ptr(r_u8), ptr(luma_lut), ptr(y), nrhs, z0, z1, ptr(out), st), "vittf_bls_slice_slab")
    return out, iters


@_on_device
def binary_erosion(mask_u8, connectivity):
    """scipy.ndimage.binary_erosion(mask, generate_binary_structure(3, connectivity)) (border_value 0) on the device:
    mask uint8 (W,H,D) -> uint8 (W,H,D)."""
    require_cuda(mask_u8)
    if mask_u8.dtype != torch.uint8 or mask_u8.dim() != 3:
        raise TypeError("binary_erosion expects a 3-D uint8 mask")
    W, H, D = mask_u8.shape
    out = torch.empty_like(mask_u8)
    check(load().vittf_binary_erosion(ptr(mask_u8), W, H, D, int(connectivity), ptr(out), stream_ptr(mask_u8.device)),
          "vittf_binary_erosion")
    return out


@_on_device
def topk_voxels(maps, K):
    """maps fp32 (n_maps, n) CUDA -> (int64 (n_maps, K) flat indices, fp32 (n_maps) thresholds) with the tie rule of
    infer.py:92-93 (first K voxels in index order with value >= K-th largest value)."""
    require_cuda(maps)
    n_maps, n = maps.shape
    idx = torch.empty(n_maps, K, dtype=torch.int64, device=maps.device)
    thr = torch.empty(n_maps, dtype=torch.float32, device=maps.device)
    check(load().vittf_topk_voxels(ptr(maps), n_maps, n, int(K), ptr(idx), ptr(thr), stream_ptr(maps.device)), "vittf_topk_voxels")
    return idx, thr


@_on_device
def mean_pairwise_distance(feats, measure):
    """feats fp32 (N, F) CUDA -> fp32 (N): 1 - mean cosine similarity ('cosine') or mean Euclidean distance ('euclidean')."""
    require_cuda(feats)
    code = {"cosine": 0, "euclidean": 1}.get(measure)
    if code is None:
        raise ValueError(f'Unknown measure: {measure}')
    N, F_ = feats.shape
    out = torch.empty(N, dtype=torch.float32, device=feats.device)
    check(load().vittf_mean_pairwise_distance(ptr(feats), N, F_, code, ptr(out), stream_ptr(feats.device)),
          "vittf_mean_pairwise_distance")
    return out


@_on_device
def confusion_matrix(truth_u8, pred_u8, K):
    """(K, K) int64 table of (true, predicted) label pairs of two uint8 CUDA tensors of equal size (rows = true labels,
    as sklearn.metrics.confusion_matrix); labels >= K raise."""
    require_cuda(truth_u8)
    require_cuda(pred_u8)
    if truth_u8.dtype != torch.uint8 or pred_u8.dtype != torch.uint8 or truth_u8.numel() != pred_u8.numel():
        raise TypeError("confusion_matrix expects two uint8 tensors of the same size")
    t, p = truth_u8.contiguous(), pred_u8.contiguous()
    out = torch.empty(K * K, dtype=torch.int64, device=t.device)
    bad = torch.empty(1, dtype=torch.int32, device=t.device)
    check(load().vittf_confusion_matrix(ptr(t), ptr(p), t.numel(), int(K), ptr(out), ptr(bad), stream_ptr(t.device)),
          "vittf_confusion_matrix")
    if int(bad.item()) != 0:
        raise ValueError(f"confusion_matrix: {int(bad.item())} voxels carry labels >= K={K}")
    return out.view(K, K)
