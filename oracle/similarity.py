"""CPU restatement of the reference's similarity stage (stage 2).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Three semantics exist in the reference tree; each is restated:

  REF-NTF  /root/reference/predict_ntf.py:24-101   raw dot at feature resolution,
           ``where(sim>=.25, sim, 0)**2.5``, MEAN over a class's annotations,
           uint8 quantisation with the wrap hazard, NN resize to in_dims//2.
  LEGACY   /root/reference/old/cluster_dino.py:306-345   F.normalize(dim=0),
           nearest prototypes, ``clamp(0,1)**e``, MAX over annotations, argmax.
  NS       the north-star composition of the same torch ops: trilinear
           up-sampling of the features (F.interpolate, align_corners=False, as in
           predict_ntf.py:87), F.normalize(dim=0) (cluster_dino.py:307), einsum
           (predict_ntf.py:65), clamp/pow + max (cluster_dino.py:318,322).

Prototype lookup follows /root/reference/infer.py:48-72 (grid_sample, zero
padding, align_corners=False; coordinates flipped X,Y,Z -> z,y,x).

Pinned by tests/golden/sim_*.npz (outputs of the reference's own functions).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F


def rel_coords(abs_coords, vol_shape):
    """predict_ntf.py:56 -- voxel index (volume space) -> [-1, 1]."""
    ext = torch.tensor([list(vol_shape[-3:])], dtype=torch.float32)
    return (abs_coords.float() + 0.5) / ext * 2.0 - 1.0


def sample_prototypes(feats, rel, mode="bilinear"):
    """infer.py:48-72.  feats (F,w,h,d), rel (A,3) in X,Y,Z order -> (A,F)."""
    grid = rel.flip(-1).to(feats.dtype).view(1, 1, 1, -1, 3)
    out = F.grid_sample(feats[None], grid, mode=mode, align_corners=False)  # (1,F,1,1,A)
    return out[0, :, 0, 0].t().contiguous()


def _split(annotations):
    off = 0
    for name, pts in annotations.items():
        yield name, off, off + pts.size(0)
        off += pts.size(0)


def quantize_u8(sim):
    """predict_ntf.py:95-99 verbatim arithmetic (values above 255 WRAP)."""
    quant = 0.99 * sim.max()
    return (255.0 / quant * sim).to(torch.uint8)


@torch.no_grad()
def ref_ntf_float(volume_shape, feats, annotations):
    """predict_ntf.py:53-72: per-class float similarity at feature resolution."""
    pts = torch.cat(list(annotations.values()))
    protos = sample_prototypes(feats, rel_coords(pts.to(feats.dtype), volume_shape), "bilinear")
    out = OrderedDict()
    single_big = len(annotations) == 1 and pts.size(0) > 1024           # predict_ntf.py:62-63
    if single_big:
        sims = torch.einsum("fwhd,af->whd", feats, protos)[None, None] / protos.size(0)
    else:
        sims = torch.einsum("fwhd,af->awhd", feats, protos)[None]        # (1,A,w,h,d)
    for name, a0, a1 in _split(annotations):
        s = sims[:, a0:a1] if not single_big else sims
        s = torch.where(s >= 0.25, s, torch.zeros(1, dtype=s.dtype)) ** 2.5
        out[name] = s.mean(dim=1)[0]
    return out


@torch.no_grad()
def ref_ntf(volume_shape, feats, annotations):
    """predict_ntf.py:24-101 with bilateral_solver=False -> uint8 (W//2,H//2,D//2)."""
    half = tuple(d // 2 for d in volume_shape[-3:])
    out = OrderedDict()
    for name, sim in ref_ntf_float(volume_shape, feats, annotations).items():
        q = quantize_u8(sim)
        out[name] = F.interpolate(q[None, None], half, mode="nearest")[0, 0]
    return out


@torch.no_grad()
def legacy(feats, annotations, volume_shape, exponent=2.0):
    """old/cluster_dino.py:306-322: per-class max of clamp(cos,0,1)**e at feature
    resolution; prototypes taken with nearest sampling from the normalised volume."""
    fn = F.normalize(feats.float(), dim=0)
    pts = torch.cat(list(annotations.values()))
    protos = sample_prototypes(fn, rel_coords(pts, volume_shape), "nearest")
    sims = torch.einsum("fwhd,af->awhd", fn, protos).clamp(0, 1) ** exponent
    return OrderedDict((n, sims[a0:a1].max(dim=0).values.clamp(0, 1)) for n, a0, a1 in _split(annotations))


@torch.no_grad()
def ns_composite(feats, protos, class_offsets, out_shape, exponent=2.0, slab=16, z_range=None):
    """North-star order: up-sample FEATURES -> normalise -> dot -> clamp/pow -> class max.

    feats (F,w,h,d) any float dtype; protos (A,F) fp32 (normalised here, as the
    legacy path normalises its prototypes); class_offsets list of C+1 ints.
    Evaluated in z-slabs of the OUTPUT so the up-sampled features never exceed
    host memory.  Returns fp32 (C, W, H, z1-z0) for the output z-range (default: all of D)."""
    f32 = feats.float()
    pn = F.normalize(protos.float(), dim=-1)
    n_cls = len(class_offsets) - 1
    W, H, D = out_shape
    zlo, zhi = (0, D) if z_range is None else z_range
    out = torch.empty((n_cls, W, H, zhi - zlo), dtype=torch.float32)
    w, h, d = f32.shape[1:]
    # F.interpolate over a z-slab must reproduce the full-volume index rule, so the slab
    # is cut on the OUTPUT grid and the source window is chosen to contain every tap.
    scale = d / D
    for z0 in range(zlo, zhi, slab):
        z1 = min(zhi, z0 + slab)
        src = [max((z + 0.5) * scale - 0.5, 0.0) for z in range(z0, z1)]
        lo = int(src[0])
        hi = min(int(src[-1]) + 1, d - 1)
        up = _interp_slab(f32[..., lo:hi + 1], (W, H), src, lo)        # (F,W,H,z1-z0)
        up = F.normalize(up, dim=0)
        s = torch.einsum("fwhd,af->awhd", up, pn).clamp(0, 1) ** exponent
        for c in range(n_cls):
            out[c, :, :, z0 - zlo:z1 - zlo] = s[class_offsets[c]:class_offsets[c + 1]].max(dim=0).values
    return out


def _interp_slab(f, wh, src_z, lo):
    """Trilinear (align_corners=False) up-sampling of ``f`` (F,w,h,dz) to (F,W,H,len(src_z)):
    in-plane by F.interpolate(bilinear) -- the same separable index rule ATen's
    upsample_trilinear3d uses -- then an explicit lerp along z at positions ``src_z``."""
    fz = f.permute(0, 3, 1, 2)                                          # (F,dz,w,h)
    up = F.interpolate(fz, size=wh, mode="bilinear", align_corners=False)  # (F,dz,W,H)
    i0 = torch.tensor([int(s) - lo for s in src_z])
    i1 = torch.clamp(i0 + 1, max=up.size(1) - 1)
    t = torch.tensor([s - int(s) for s in src_z], dtype=torch.float32).view(1, -1, 1, 1)
    out = up[:, i0] * (1 - t) + up[:, i1] * t
    return out.permute(0, 2, 3, 1)


@torch.no_grad()
def ns_composite_direct(feats, protos, class_offsets, out_shape, exponent=2.0):
    """Un-slabbed NS oracle made of the literal torch calls (small inputs only)."""
    up = F.interpolate(feats[None].float(), size=tuple(out_shape), mode="trilinear", align_corners=False)[0]
    up = F.normalize(up, dim=0)
    s = torch.einsum("fwhd,af->awhd", up, F.normalize(protos.float(), dim=-1)).clamp(0, 1) ** exponent
    return torch.stack([s[class_offsets[c]:class_offsets[c + 1]].max(dim=0).values
                        for c in range(len(class_offsets) - 1)])


def ns_at_voxels(feats, protos, class_offsets, out_shape, voxels, exponent=2.0):
    """NS similarity of selected OUTPUT voxels only (full-size spot checks: the slab oracle needs minutes at 512^3).
    Restates the index rule of F.interpolate(mode='trilinear', align_corners=False) -- src = max((dst + 0.5) * in / out
    - 0.5, 0), i0 = floor(src), i1 = min(i0 + 1, in - 1), t = src - i0 -- on the 8 corners of each voxel; pinned to
    ns_composite by tests/test_oracle.py::test_ns_point_oracle_equals_slab_oracle.  voxels (M,3) long -> (C, M) fp32."""
    pn = F.normalize(protos.float(), dim=-1)
    dims = feats.shape[1:]
    idx, wgt = [], []
    for ax in range(3):
        src = ((voxels[:, ax].double() + 0.5) * (dims[ax] / out_shape[ax]) - 0.5).clamp_min(0.0)
        i0 = src.floor().long()
        i1 = (i0 + 1).clamp_max(dims[ax] - 1)
        t = (src - i0).float()
        idx.append((i0, i1))
        wgt.append((1.0 - t, t))
    up = torch.zeros(voxels.size(0), feats.size(0), dtype=torch.float32)
    for bx in range(2):
        for by in range(2):
            for bz in range(2):
                w = wgt[0][bx] * wgt[1][by] * wgt[2][bz]
                up += w[:, None] * feats[:, idx[0][bx], idx[1][by], idx[2][bz]].float().t()
    up = F.normalize(up, dim=1)
    s = (up @ pn.t()).clamp(0, 1) ** exponent                                     # (M, A)
    return torch.stack([s[:, class_offsets[c]:class_offsets[c + 1]].max(dim=1).values
                        for c in range(len(class_offsets) - 1)])


def compose_labels(sims_u8, thresholds):
    """predict_ntf.py:203-215: thresholded running arg-max, strict '>', 0 = background.
    sims_u8 (C,...) uint8; thresholds list of floats in [0,1]."""
    sims = sims_u8.float()
    pred = torch.zeros_like(sims[0])
    best = torch.zeros_like(sims[0])
    for i in range(sims.size(0)):
        m = (sims[i] > int(thresholds[i] * 255)) & (sims[i] > best)
        pred[m] = i + 1
        best[m] = sims[i][m]
    return pred.to(torch.uint8)


def argmax_labels(sims):
    """old/cluster_dino.py:345."""
    return sims.argmax(0)
