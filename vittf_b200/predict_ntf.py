"""Drop-in for the similarity part of /root/reference/predict_ntf.py: ``compute_similarities``
(:24-101) and the label composition of its ``__main__`` (:203-215).  The prototype gather, the
feature x prototype contraction, threshold/pow/mean, the per-class maxima, the bilateral solver and
the label composition run in libvittf_b200.so; torch is used for the tiny glue the reference also
does in torch (trilinear resize of the low-res maps, the uint8 cast with its wrap-around)."""
import torch
import torch.nn.functional as F

from . import ops
from .bilateral_solver3d import crop_pad, solve_many, write_crop_into
from .infer import _cuda_device, make_4d, make_5d
from .similarity import class_offsets, rel_coords, similarity_maps

BLS_GRID = {'sigma_spatial': 7, 'sigma_chroma': 5, 'sigma_luma': 5}      # predict_ntf.py:75-79


def similarity_float(volume_shape, features, annotations):
    """predict_ntf.py:53-72 on the device: per-class float maps at feature resolution, (C,w,h,d) fp32."""
    dev = features.device
    pts = torch.cat(list(annotations.values()))
    rel = rel_coords(pts, volume_shape, dev)
    feats = features if features.dtype in (torch.float16, torch.float32) else features.float()
    feats = feats.contiguous()
    qf = ops.sample_prototypes(feats, rel, 'bilinear')                    # (A, F) fp32
    if features.dtype == torch.float16:
        qf = qf.half().float()                                            # reference keeps qf in the feature dtype
    if len(annotations) == 1 and pts.size(0) > 1024:                      # :62-63, dot with the mean prototype
        qf = qf.mean(dim=0, keepdim=True)
        offs = torch.tensor([0, 1], dtype=torch.int32, device=dev)
    else:
        offs = class_offsets(annotations, dev)
    return similarity_maps(feats, qf.contiguous(), offs, mode="refntf", exponent=2.5, threshold=0.25)


def compute_similarities(volume, features, annotations, bilateral_solver=False):
    """predict_ntf.py:24-101.  volume (W,H,D), features (F,w,h,d), annotations {name: (N,3)} ->
    {name: uint8 (W//2,H//2,D//2)} on the CPU; None without annotations."""
    with torch.no_grad():
        if len(annotations) == 0:
            return
        if sum(int(v.numel()) for v in annotations.values()) == 0:
            return
        features = torch.as_tensor(features)
        dev = _cuda_device(features.device if features.is_cuda else None)
        features = features.to(dev)
        in_dims = tuple(volume.shape[-3:])
        sim_shape = tuple(d // 2 for d in in_dims)
        sims = similarity_float(in_dims, features, annotations)           # (C, w, h, d)
        similarities = {}
        if not bilateral_solver:
            quant = 0.99 * ops.class_max(sims)                            # :98
            for i, k in enumerate(annotations.keys()):
                q = (255.0 / quant[i] * sims[i]).cpu().to(torch.uint8).squeeze()      # :99 verbatim (wraps > 255)
                similarities[k] = F.interpolate(make_5d(q), sim_shape, mode='nearest').squeeze()
            return similarities
        vol = F.interpolate(make_5d(torch.as_tensor(volume).to(dev).float()), sim_shape, mode='trilinear').squeeze()   # :80
        mm = ops.minmax(vol.contiguous())
        vol = (255.0 * ((vol - mm[0]) / (mm[1] - mm[0]))).to(torch.uint8)                                                # :83-84
        for i, k in enumerate(annotations.keys()):
            sim = sims[i:i + 1]
            if tuple(sim.shape[-3:]) != sim_shape:
                sim = F.interpolate(make_5d(sim), sim_shape, mode='trilinear').squeeze(0)                               # :87
            crops, mima = crop_pad([sim, vol], thresh=0.1, pad=2)                                                        # :90
            csim, cvol = crops
            solved, _ = solve_many(make_4d(csim).contiguous(), cvol.contiguous(), None, grid_params=BLS_GRID)           # :92
            sim = write_crop_into(sim.clone(), solved[0], mima)                                                          # :93
            quant = 0.99 * ops.class_max(sim.reshape(1, *sim_shape).contiguous())[0]                                     # :95
            similarities[k] = (255.0 / quant * sim).cpu().to(torch.uint8).squeeze()                                      # :96
        return similarities


def compose_labels(similarities, thresholds):
    """predict_ntf.py:203-215: thresholded running arg-max over the class maps (strict '>', earlier
    class wins ties, 0 = background).  similarities {name: uint8 map} or (C,...) tensor."""
    sims = torch.stack(list(similarities.values())) if isinstance(similarities, dict) else similarities
    dev = _cuda_device(sims.device if sims.is_cuda else None)
    thr = torch.tensor([int(t * 255) for t in thresholds], dtype=torch.int32, device=dev)
    return ops.labels(sims.to(dev).contiguous(), thr, mode=0).cpu()


def argmax_labels(sims):
    """old/cluster_dino.py:345 `pred_sims.argmax(0)` on float class maps."""
    dev = _cuda_device(sims.device if sims.is_cuda else None)
    return ops.labels(sims.to(dev).float().contiguous(), None, mode=1)
