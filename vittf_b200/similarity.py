"""Prototype-similarity stage on the device (SURVEY.md §0.2 modes NS / REF-NTF / LEGACY).

Two native passes (include/vittf.h): ``vittf_sim_lowres`` reads the feature volume once and produces
per-voxel prototype dots (+ the 14 Gram scalars that give |interp(f)|^2), ``vittf_sim_upsample``
evaluates every output voxel from them.  The up-sampled feature volume is never materialised.
"""
import os

import torch

from . import _lib, ops

TC_MAX_PROTOS = 8
FORCE_TC = os.environ.get("VITTF_SIM_UP_TC") is not None          # A/B switch: tcgen05 up-sampling kernel for every prototype count

MODES = {"ns": _lib.SIM_NS, "refntf": _lib.SIM_REFNTF, "legacy": _lib.SIM_LEGACY, "clamp_mean": _lib.SIM_CLAMP_MEAN}


def class_offsets(annotations, device):
    sizes = [int(v.shape[0]) for v in annotations.values()]
    off = [0]
    for s in sizes:
        off.append(off[-1] + s)
    return torch.tensor(off, dtype=torch.int32, device=device)


def rel_coords(abs_coords, vol_shape, device):
    """predict_ntf.py:56: voxel index in volume space -> [-1, 1] (X,Y,Z order)."""
    ext = torch.tensor([list(vol_shape[-3:])], dtype=torch.float32, device=device)
    return ((abs_coords.to(device).float() + 0.5) / ext * 2.0 - 1.0).contiguous()


def lowres_x_planes(lr_w, out_w, x_range):
    """Low-res x planes [xa, xb) that an output x-slab reads (trilinear footprint: its cells +- one plane)."""
    if x_range is None or out_w % lr_w:
        return None
    u = out_w // lr_w
    x0, x1 = x_range
    c_lo, c_hi = (x0 - u // 2) // u, (x1 - 1 - u // 2) // u            # floor division: cells -1 .. lr_w - 1
    return max(0, c_lo), min(lr_w, c_hi + 2)


def similarity_maps(feats, protos, offsets, out_shape=None, mode="ns", exponent=2.0, threshold=0.25, z_range=None,
                    lowres=None, x_range=None):
    """feats (F,w,h,d) fp16|fp32 CUDA; protos (A,F) fp32 CUDA (already normalised for ns/legacy);
    offsets int32 (C+1) CUDA -> fp32 (C, x1-x0, H, z1-z0): the output slab `x_range` x `z_range` (default: everything).
    With an x-slab pass 1 only evaluates the low-res planes under it.  `lowres` lets a caller reuse pass 1."""
    lr = tuple(feats.shape[1:])
    out_shape = lr if out_shape is None else tuple(out_shape)
    m = MODES[mode]
    if m != _lib.SIM_NS and out_shape != lr:
        raise ValueError("refntf/legacy/clamp_mean similarities are defined at feature resolution")
    if lowres is None:
        # voxel-major dots where the tcgen05 up-sampling kernel will read them: NS mode, one integer factor 2 / 4 / 8, and
        # few prototypes (measured on B200, profiles/r2_sim_kernels.md: its cost per prototype is ~3x that of the warp-level
        # mma.sync kernel -- every prototype is a TMEM round trip -- while its fixed cost per output block is lower)
        u = out_shape[0] // lr[0] if lr[0] else 0
        tc = (m == _lib.SIM_NS and u in (2, 4, 8) and all(o == u * i for o, i in zip(out_shape, lr)) and out_shape[2] % 4 == 0
              and (protos.shape[0] <= TC_MAX_PROTOS or u == 2 or FORCE_TC))
        lowres = ops.sim_lowres(feats, protos, want_gram=(m not in (_lib.SIM_REFNTF, _lib.SIM_CLAMP_MEAN)), voxel_major=tc,
                                x_planes=lowres_x_planes(lr[0], out_shape[0], x_range))
    dots, gram, layout = lowres
    z0, z1 = (0, out_shape[2]) if z_range is None else z_range
    x0, x1 = (0, out_shape[0]) if x_range is None else x_range
    return ops.sim_upsample(dots, gram, lr, offsets, out_shape, m, threshold, exponent, z0, z1, layout=layout, n_protos=protos.shape[0],
                            x0=x0, x1=x1)
