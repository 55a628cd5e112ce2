"""Drop-in for the similarity part of /root/reference/predict_ntf.py: ``compute_similarities``
(:24-101) and the label composition of its ``__main__`` (:203-215).  The prototype gather, the
feature x prototype contraction, threshold/pow/mean, the per-class maxima, the bilateral solver and
the label composition run in libvittf_b200.so; torch is used for the tiny glue the reference also
does in torch (trilinear resize of the low-res maps, the uint8 cast with its wrap-around)."""
import torch
import torch.nn.functional as F

from . import ops
from .bilateral_solver3d import crop_pad, solve_many, write_crop_into
from .compare_feat_sampling import sample_both, sample_surface, sample_uniform
from .infer import _cuda_device, make_4d, make_5d
from .similarity import class_offsets, rel_coords, similarity_maps

BLS_GRID = {'sigma_spatial': 7, 'sigma_chroma': 5, 'sigma_luma': 5}      # predict_ntf.py:75-79

sampling_modes = {                                                      # predict_ntf.py:17-22
    'uniform': sample_uniform,
    'surface': sample_surface,
    'both': sample_both,
    'annotated': lambda *args, **kwargs: None
}
CT_ORG_NAMES = ['liver', 'bladder', 'lung', 'kidney', 'bone']           # :206
CT_ORG_THRESHOLDS = [0.486, 0.264, 0.236, 0.68, 0.291]                   # :207


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


def main(argv=None):
    """predict_ntf.py:104-222: the reference's CLI on a data directory (volume.npy, *features*.npy, labels.npy and/or
    annotations.npy) -> ntf_pred{num_samples}{sampling_mode}{bls}.npy (uint8 label volume at similarity resolution).
    Same flags, file discovery, flips (:141,146: volume and labels along dim -3, NOT features/annotations),
    sampling, per-class thresholds and running arg-max; `--gpu` is accepted for compatibility -- this implementation
    always runs on the GPU (fp16 features with --gpu, fp32 otherwise, like the reference's dtype choice :113-116)."""
    import time
    from argparse import ArgumentParser
    from pathlib import Path

    import numpy as np

    parser = ArgumentParser()
    parser.add_argument('--data', type=str, help='Path to features, annotations, volume etc.')
    parser.add_argument('--bilateral-solver', action='store_true', help='Use bilateral solver')
    parser.add_argument('--load-sims', action='store_true', help='Load similarities from file')
    parser.add_argument('--num-samples', type=float, default=0.0, help='Number of samples to use for each NTF')
    parser.add_argument('--sampling-mode', type=str, choices=['uniform', 'surface', 'both'], default='both', help='Sampling mode')
    parser.add_argument('--gpu', action='store_true', help='Use GPU')
    args = parser.parse_args(argv)

    dev = _cuda_device(None)
    typ = torch.float16 if args.gpu else torch.float32
    dir = Path(args.data)
    if args.num_samples == 0.0:
        args.sampling_mode = 'annotated'
    bls_str = 'bls' if args.bilateral_solver else ''
    out_fn = dir / f'ntf_pred{args.num_samples}{args.sampling_mode}{bls_str}.npy'
    if out_fn.exists():
        print(f'Already inferred NTF preds for {dir} using sampling mode {args.sampling_mode} and {args.num_samples} samples')
        return 0
    print(f'Inferring for {dir} using sampling mode {args.sampling_mode} and {args.num_samples} samples')

    feat_fns = list(filter(lambda p: 'features' in str(p) and 'pred' not in str(p), dir.iterdir()))
    if len(feat_fns) == 0:
        raise ValueError(f'No features found in {dir}')
    elif len(feat_fns) == 1:
        feat_fn = feat_fns[0]
    else:
        feat_fn = sorted(feat_fns, key=lambda p: p.stat().st_size)[-1]
        print(f'Found multiple features in {dir}. Using largest one {feat_fn.name}.')

    volume = np.load(dir / 'volume.npy', allow_pickle=True).astype(np.float32)
    if (dir / 'labels.npy').exists():
        labels = np.load(dir / 'labels.npy', allow_pickle=True)[()]
        labels = np.flip(labels, axis=-3).copy()
    else:
        assert args.num_samples == 0.0, 'Cannot sample labels if they are not provided'
        labels = None
    features = np.load(feat_fn, allow_pickle=True)[()]
    volume = np.flip(volume, axis=-3).copy()
    if isinstance(features, dict):
        features = torch.as_tensor(features['k']).float().squeeze()
    else:
        features = torch.as_tensor(features).float().squeeze()
    draw_samples = sampling_modes[args.sampling_mode]

    if args.num_samples == 0.0:
        annotations = np.load(dir / 'annotations.npy', allow_pickle=True)[()]  # { classname: (N, 3) }
        annotations = {k: torch.as_tensor(v) for k, v in annotations.items()}
    elif args.num_samples > 0.0:
        annotations = {}
        for i in range(1, int(labels.max()) + 1):
            mask = torch.as_tensor(labels == i)
            n_lab = int(mask.sum().item())
            n_samples = min(int(args.num_samples), n_lab) if args.num_samples > 1.0 else int(args.num_samples * n_lab)
            if n_samples > 0:
                annotations[f'ntf{i}'] = draw_samples(mask, n_samples, thin_to_reasonable=True) if args.sampling_mode != 'surface' \
                    else draw_samples(mask, n_samples)
    else:
        raise Exception(f'Invalid value for --num-samples: {args.num_samples}')

    print(f'Computing similarties for {tuple(volume.shape)} with features {tuple(features.shape)}')
    t0 = time.time()
    t1 = t0
    if args.load_sims:
        similarities = {k: torch.as_tensor(v) for k, v in np.load(dir / 'similarities.npy', allow_pickle=True)[()].items()}
        t2 = t1
    else:
        feats_dev = features.to(device=dev, dtype=typ)
        t1 = time.time()
        if torch.cat(list(annotations.values())).size(0) > 10000:                      # :166-169: one class at a time
            similarities = {k: compute_similarities(volume, feats_dev, {k: v}, bilateral_solver=args.bilateral_solver)[k]
                            for k, v in annotations.items()}
        else:
            similarities = compute_similarities(volume, feats_dev, annotations, bilateral_solver=args.bilateral_solver)
        torch.cuda.synchronize()
        t2 = time.time()
        similarities = {k: v.cpu().float() for k, v in similarities.items()}
    print('Similarities:', {k: v.shape for k, v in similarities.items()})
    sims = torch.stack(list(similarities.values()))
    # :203-215 -- zip() stops at the shorter of (5 CT-ORG thresholds, classes)
    n_cls = min(sims.size(0), len(CT_ORG_THRESHOLDS))
    pred = compose_labels(sims[:n_cls], CT_ORG_THRESHOLDS[:n_cls]).numpy().astype(np.uint8)
    np.save(out_fn, pred)
    if tuple(pred.shape[-3:]) != tuple(volume.shape[-3:]):
        pred = F.interpolate(make_5d(torch.as_tensor(pred)), tuple(volume.shape[-3:]), mode='nearest').squeeze().numpy()
    print('Pred:', pred.shape, pred.min(), pred.max())
    print('NTF fit time:', t1 - t0)
    print('NTF predict time:', t2 - t1)
    return 0


if __name__ == '__main__':
    import sys
    sys.exit(main())
