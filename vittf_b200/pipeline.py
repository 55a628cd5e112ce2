"""Whole hot path for one volume: raw voxels -> 3-axis ViT K-feature volume -> prototype similarity
(north-star order: trilinear up-sampling of the features, L2 normalisation, dot, clamp/pow, per-class
max) -> label volume.  Used by bench.py, __graft_entry__.smoke() and the end-to-end tests; every
arithmetic step is a libvittf_b200 kernel.
"""
import torch
import torch.nn.functional as F

from . import dist, infer, ops
from .bilateral_solver3d import solve_many_sharded
from .similarity import class_offsets, rel_coords, similarity_maps


def prototypes(feats, annotations, vol_shape, mode="bilinear", normalize=True):
    """Prototype vectors of the annotated voxels (infer.py:48-72 lookup), (A, F) fp32 on the device."""
    pts = torch.cat(list(annotations.values()))
    rel = rel_coords(pts, vol_shape, feats.device)
    protos = ops.sample_prototypes(feats, rel, mode)
    return F.normalize(protos, dim=-1).contiguous() if normalize else protos


def volume_to_similarity(vol_dev, model, annotations, patch=8, fos=64, batch_size=8, out_shape=None, exponent=2.0,
                         rank=0, world=1, group=None, want_labels=True):
    """Returns (feature volume fp16 (D,f,f,f), similarity maps fp32 (C,W,H,z1-z0), labels uint8 (W,H,z1-z0)
    or None, (z0, z1)).  With world > 1 the maps/labels cover this rank's z-slab."""
    feats = infer.feature_volume(vol_dev, model, patch, fos, batch_size, dev=vol_dev.device, rank=rank, world=world,
                                 group=group)
    vol_shape = tuple(vol_dev.shape[-3:])
    out_shape = vol_shape if out_shape is None else tuple(out_shape)
    protos = prototypes(feats, annotations, vol_shape)
    offs = class_offsets(annotations, feats.device)
    zr = dist.z_range(out_shape[2], world, rank)
    sims = similarity_maps(feats, protos, offs, out_shape, mode="ns", exponent=exponent, z_range=zr)
    labels = ops.labels(sims, None, mode=1) if want_labels else None
    return feats, sims, labels, zr


def distribute_volume(vol_host, dev, rank=0, world=1, group=None):
    """Host -> devices: the replicated raw volume every rank slices along all three axes.  With several ranks each one
    copies 1/world of the (pinned) bytes over its own PCIe link and ONE all-gather over NVLink assembles the copies --
    instead of `world` full host reads (broadcast once)."""
    import torch.distributed as td
    n = vol_host.numel()
    if world == 1 or n % world or not (td.is_available() and td.is_initialized()):
        return vol_host.to(dev, non_blocking=True)
    flat = vol_host.reshape(-1)
    part = flat[rank * (n // world):(rank + 1) * (n // world)].to(dev, non_blocking=True)
    full = torch.empty(n, dtype=vol_host.dtype, device=dev)
    td.all_gather_into_tensor(full, part, group=group)
    return full.view(vol_host.shape)


def quantized_maps(sims_slab, z_range, depth, group=None):
    """The uint8 maps compute_similarities hands back (predict_ntf.py:95-100: 0.99*max quantisation with the
    reference's wrap, nearest-resized to half the grid) for this rank's z-slab of fp32 maps (C,W,H,z1-z0).  The class
    maxima are global: one all-reduce(max) of C floats when several ranks hold slabs.
    Returns (uint8 (C, W//2, H//2, planes of this slab), (zo0, zo1))."""
    import torch.distributed as td
    cmax = ops.class_max(sims_slab)
    if td.is_available() and td.is_initialized() and td.get_world_size(group) > 1:
        td.all_reduce(cmax, op=td.ReduceOp.MAX, group=group)
    C_, W, H, _ = sims_slab.shape
    return ops.quantize_maps_u8(sims_slab, cmax, (W // 2, H // 2, depth // 2), depth=depth, z0=z_range[0])


def refine_similarity(sims_slab, ref_u8, z_range, grid_params=None, bs_params=None, group=None):
    """bilateral_solver3d refinement of per-class maps (configs[4]): sims_slab fp32 (C,W,H,z1-z0) = this rank's
    z-slab, ref_u8 the full grey uint8 volume (W,H,D) at the maps' resolution.  All classes share the reference and
    are solved together; with several ranks the pixel passes are slab-local and the grid vectors are all-reduced
    (bilateral_solver3d.solve_many_sharded).  Returns the refined slab fp32 (C,W,H,z1-z0)."""
    gp = {'sigma_spatial': 7, 'sigma_chroma': 5, 'sigma_luma': 5} if grid_params is None else grid_params   # predict_ntf.py:75-79
    out, _ = solve_many_sharded(sims_slab, ref_u8, z_range, None, gp, bs_params or {}, group=group)
    return out
