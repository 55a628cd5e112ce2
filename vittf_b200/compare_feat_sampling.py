"""Drop-in for the annotation samplers of /root/reference/compare_feat_sampling.py:13-33 (used by
predict_ntf.py:183-195 to turn label volumes into annotation point sets).

Same names, arguments and return values (LongTensor (n, 3) of voxel indices, on the CPU like the reference).
The volume-sized work -- the two erosions of `sample_surface` and the index extraction -- runs on the GPU
(libvittf_b200 `vittf_binary_erosion`, then `nonzero` as device glue); the draw itself is the reference's
`torch.multinomial` over uniform weights on the CPU generator, so a seeded run selects the same voxels as the
reference does.
"""
import numpy as np
import torch

from . import ops

ONE = torch.ones(1)                                                   # compare_feat_sampling.py:11


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("vittf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _mask_u8(vol):
    t = torch.as_tensor(np.ascontiguousarray(vol) if isinstance(vol, np.ndarray) else vol)
    if t.dim() != 3:
        raise ValueError(f"sampler expects a 3-D label mask, got shape {tuple(t.shape)}")
    return (t != 0).to(_device(), torch.uint8).contiguous()


def sample_uniform(vol, n_samples, thin_to_reasonable=False):
    """:13-17 -- n_samples distinct foreground voxels, uniformly."""
    idxs = _mask_u8(vol).nonzero().cpu()
    while thin_to_reasonable and idxs.size(0) > int(2 ** 24):
        idxs = idxs[::2]
    return idxs[torch.multinomial(ONE.expand(idxs.size(0)), n_samples)]


def surface_voxels(vol, dist_from_surface=4):
    """:20-25 -- outer = erode(vol, structure(3, dist)), inner = erode(outer, structure(3, 1)); surface = inner XOR outer."""
    m = _mask_u8(vol)
    outer = ops.binary_erosion(m, dist_from_surface)
    inner = ops.binary_erosion(outer, 1)
    return (inner ^ outer).nonzero().cpu()


def sample_surface(vol, n_samples, dist_from_surface=4):
    """:19-31 -- up to n_samples voxels of the eroded surface shell (all of them, with the reference's message, if fewer)."""
    surface_idxs = surface_voxels(vol, dist_from_surface)
    if surface_idxs.size(0) > n_samples:
        return surface_idxs[torch.multinomial(ONE.expand(surface_idxs.size(0)), n_samples)]
    print(f'Full surface only has {surface_idxs.size(0)} voxels (< n_samples={n_samples}).')
    return surface_idxs


def sample_both(vol, n_samples, dist_from_surface=4, thin_to_reasonable=False):
    """:33-34"""
    return torch.cat([sample_uniform(vol, n_samples // 2, thin_to_reasonable=thin_to_reasonable),
                      sample_surface(vol, n_samples // 2, dist_from_surface=dist_from_surface)])
