"""Drop-in for /root/reference/bilateral_solver3d.py: ``apply_bilateral_solver3d``, ``crop_pad``,
``write_crop_into`` with the reference's signatures, parameter dictionaries and defaults; the
solver itself (grid build, bistochastisation, PCG, slice) runs in libvittf_b200.so.

Scope: grey reference volumes (r[0] == r[1] == r[2]), which is what the hot path always passes
(predict_ntf.py:92 ``cvol.expand(3, -1, -1, -1)``); colour references raise NotImplementedError.
"""
import numpy as np
import torch

from . import ops

__all__ = ['apply_bilateral_solver3d', 'crop_pad', 'write_crop_into']

_RGB2Y = np.array([[0.299, 0.587, 0.114], [-0.168736, -0.331264, 0.5], [0.5, -0.418688, -0.081312]])

grid_params_default = {'sigma_luma': 4, 'sigma_chroma': 4, 'sigma_spatial': 24}          # :156-160
bs_params_default = {'lam': 256, 'A_diag_min': 1e-5, 'cg_tol': 1e-5, 'cg_maxiter': 25}   # :162-167


def luma_lut(sigma_luma):
    """Luma bin of every grey level with the reference's exact float64 expression (:20,46) -- it is
    NOT v // sigma, so the table is built on the host and handed to the kernels."""
    v = np.arange(256, dtype=np.uint8)
    y = np.tensordot(np.stack([v, v, v], -1)[None, None], _RGB2Y, ([3], [1]))[0, 0, :, 0] + 0.0
    return (y / sigma_luma).astype(int).astype(np.int32)


def _cuda(t, dev):
    return t.to(dev).contiguous()


def solve_many(t, r0_u8, c=None, grid_params={}, bs_params={}):
    """Native solve of several targets over one grey reference.  t (n,W,H,D) float CUDA, r0_u8 (W,H,D)
    uint8 CUDA, c (W,H,D) float CUDA or None (Sobel confidence).  Returns (fp32 (n,W,H,D) CUDA, iters)."""
    gp = {**grid_params_default, **grid_params}
    bs = {**bs_params_default, **bs_params}
    lut = luma_lut(gp['sigma_luma'])
    lut_dev = torch.from_numpy(lut).to(t.device)
    return ops.bls_solve(t.float().contiguous(), r0_u8.contiguous(), None if c is None else c.float().contiguous(), lut_dev,
                         gp['sigma_spatial'], bs['lam'], bs['A_diag_min'], bs['cg_tol'], bs['cg_maxiter'], int(lut.max()) + 1)


def solve_many_sharded(t_slab, r0_u8, z_range, c_slab=None, grid_params={}, bs_params={}, group=None):
    """Multi-GPU form of `solve_many`: this rank holds the z-slab `z_range` of the targets
    (n,W,H,z1-z0) and of the optional confidence, and the full grey reference (W,H,D).  Pixel passes (Sobel, splat,
    slice) are rank-local; ONE all-reduce(max) of the Sobel maximum and ONE all-reduce(sum) of the splatted grid
    vectors (NCCL via torch.distributed) replace the halo exchange; the small grid problem is solved replicated.
    Returns (fp32 slab (n,W,H,z1-z0) CUDA, iters)."""
    import torch.distributed as dist
    gp = {**grid_params_default, **grid_params}
    bs = {**bs_params_default, **bs_params}
    lut = luma_lut(gp['sigma_luma'])
    lut_dev = torch.from_numpy(lut).to(t_slab.device)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    red_max = (lambda x: dist.all_reduce(x, op=dist.ReduceOp.MAX, group=group)) if multi else None
    red_sum = (lambda x: dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)) if multi else None
    z0, z1 = z_range
    return ops.bls_solve_sharded(t_slab.float().contiguous(), r0_u8.contiguous(),
                                 None if c_slab is None else c_slab.float().contiguous(), lut_dev, gp['sigma_spatial'],
                                 bs['lam'], bs['A_diag_min'], bs['cg_tol'], bs['cg_maxiter'], int(lut.max()) + 1, z0, z1,
                                 red_max, red_sum)


def apply_bilateral_solver3d(t, r, c=None, grid_params={}, bs_params={}):
    """bilateral_solver3d.py:211-245.  t (1,W,H,D) float in [0,1], r (3,W,H,D) uint8, c optional
    (1,W,H,D) -> float32 (W,H,D) on the CPU (like the reference)."""
    if not torch.cuda.is_available():
        raise RuntimeError("vittf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = t.device if t.is_cuda else torch.device('cuda', torch.cuda.current_device())
    r = torch.as_tensor(r)
    if r.dtype != torch.uint8:
        raise TypeError("reference volume r must be uint8 in [0,255]")
    if not (r.stride(0) == 0 or (torch.equal(r[0], r[1]) and torch.equal(r[0], r[2]))):
        raise NotImplementedError("vittf_b200 bilateral solver supports grey references only (r[0]==r[1]==r[2])")
    shape = tuple(t.shape[-3:])
    tt = _cuda(torch.as_tensor(t).reshape((1,) + shape), dev)
    cc = None if c is None else _cuda(torch.as_tensor(c).reshape(shape), dev)
    out, _ = solve_many(tt, _cuda(r[0], dev), cc, grid_params, bs_params)
    return out[0].cpu()


def crop_pad(sim, thresh=0.1, pad=0):
    """:183-204 -- bounding box of `sim > thresh` (first element if a list), padded, clamped."""
    others = sim if isinstance(sim, list) else [sim]
    first = others[0]
    nz = torch.nonzero(first > thresh)
    mi = torch.clamp(nz.min(dim=0).values[-3:] - pad, 0, None)
    ma = torch.minimum(nz.max(dim=0).values[-3:] + pad + 1, torch.tensor(first.shape[-3:], device=nz.device))
    cut = [s[..., mi[0]:ma[0], mi[1]:ma[1], mi[2]:ma[2]] for s in others]
    return (cut, (mi, ma)) if len(others) > 1 else (cut[0], (mi, ma))


def write_crop_into(uncropped, crop, mima):
    """:206-209"""
    mi, ma = mima
    uncropped[..., mi[0]:ma[0], mi[1]:ma[1], mi[2]:ma[2]] = crop
    return uncropped
