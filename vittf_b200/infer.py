"""Drop-in for /root/reference/infer.py (feature-volume stage): same function names, argument
meaning, CLI flags and saved-file layout; the arithmetic runs in libvittf_b200.so on a B200.

    python -m vittf_b200.infer --data-path VOL [--cache-path C] [--dino-model vits8] [--slice-along all]
                               [--batch-size 1] [--feature-output-size 64] [--overwrite]
                               [--weights CKPT] [--seed 0]           (the last two are additions)

Differences that are deliberate and documented in DESIGN.md:
  * no CPU path (``--cpu`` exits with an error) -- the product never falls back;
  * the ViT parameters come from ``--weights`` (a DINO state dict) or seeded random init, because
    the reference's ``torch.hub.load`` (infer.py:42-43) needs the network;
  * only the K third is produced natively (the only key the reference's callers request,
    infer.py:326,331).
"""
import os
import sys
import time
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

from . import ops
from .vit import engine_for

AXIS_IMAGE_DIMS = {"z": (0, 1), "y": (0, 2), "x": (1, 2)}   # infer.py:138-147
AXIS_SLICE_DIM = {"z": 2, "y": 1, "x": 0}


# ---- tensor helpers, infer.py:10-34 ---------------------------------------------------------------
def make_nd(t, n):
    """Prepends singleton dimensions to `t` until n-dimensional (infer.py:10-18)."""
    if n < t.ndim:
        raise Exception(f'make_nd cannot reduce cardinality. Your Tensor.ndim={t.ndim} > n={n}.')
    return t if n == t.ndim else t[(None,) * (n - t.ndim)]


def make_3d(t): return make_nd(t, 3)
def make_4d(t): return make_nd(t, 4)
def make_5d(t): return make_nd(t, 5)


def norm_minmax(t):
    """(t - min) / (max - min) (infer.py:32-34); the range comes from the native reduction for CUDA input."""
    if t.is_cuda and t.dtype in (torch.uint8, torch.float16, torch.float32) and t.is_contiguous():
        mm = ops.minmax(t)
        out = (t.float() - mm[0]) / (mm[1] - mm[0])
        return out.to(t.dtype) if t.dtype == torch.float16 else out      # the reference keeps a floating input dtype
    mi, ma = t.min(), t.max()
    return (t - mi) / (ma - mi)


def _noop(x, **kwargs): return x


def _cuda_device(dev=None):
    if not torch.cuda.is_available():
        raise RuntimeError("vittf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device(dev) if dev is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("vittf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())


def sample_features3d(feat_vol, rel_coords, mode='nearest'):
    """infer.py:48-72.  feat_vol ([M,] F, W, H, D), rel_coords ([M,] C, A, 3) in [-1,1] (X,Y,Z order)
    -> (M, C, A, F) on feat_vol's device/dtype."""
    if feat_vol.ndim == 4: feat_vol = make_5d(feat_vol)
    if rel_coords.ndim in (2, 3): rel_coords = make_4d(rel_coords)
    if rel_coords.size(0) != feat_vol.size(0):
        rel_coords = rel_coords.expand(feat_vol.size(0), -1, -1, -1)
    if mode not in ('nearest', 'bilinear'):
        raise ValueError(f"sample_features3d: unsupported mode {mode!r}")
    src_dev, src_dtype = feat_vol.device, feat_vol.dtype
    dev = _cuda_device(src_dev if src_dev.type == "cuda" else None)
    M, C_, A = rel_coords.shape[:3]
    outs = []
    for m in range(M):
        fv = feat_vol[m].to(dev)
        if fv.dtype not in (torch.float16, torch.float32):
            fv = fv.float()
        # the reference casts the grid to the feature dtype before grid_sample (infer.py:67)
        rel = rel_coords[m].to(src_dtype).to(dev, torch.float32).reshape(-1, 3).contiguous()
        outs.append(ops.sample_prototypes(fv.contiguous(), rel, mode).view(C_, A, -1))
    return torch.stack(outs).to(src_dev, src_dtype).contiguous()


def resample_topk(feat_vol, sims, K=8, similarity_exponent=2.0, feature_sampling_mode='nearest'):
    """infer.py:75-106.  feat_vol ([M,] F, W, H, D) (normalised) features, sims ([M,] C, A, W, H, D) per-annotation
    similarity maps -> ([M,] C, A, W, H, D): every map is replaced by the mean over its K most similar voxels of
    clamp(<f, f_k>, 0, 1) ** similarity_exponent.  The top-K search (with the reference's tie rule: first K voxels in
    index order with s >= K-th largest), the feature gather and the K-prototype similarity run in libvittf_b200;
    arithmetic is fp32 (the reference uses fp32 for K > 4 and the feature dtype otherwise)."""
    from .similarity import similarity_maps
    if sims.ndim == 5: sims = sims.unsqueeze(0)
    if feat_vol.ndim == 4: feat_vol = make_5d(feat_vol)
    src_dev, src_dtype = feat_vol.device, feat_vol.dtype
    dev = _cuda_device(src_dev if src_dev.type == "cuda" else None)
    M, C_, A = sims.shape[:3]
    W, H, D = sims.shape[-3:]
    if tuple(feat_vol.shape[-3:]) != (W, H, D):
        raise ValueError("resample_topk: similarity maps must have the feature volume's resolution")
    out = []
    for m in range(M):
        maps = sims[m].to(dev, torch.float32).reshape(C_ * A, W * H * D).contiguous()
        flat, _ = ops.topk_voxels(maps, K)                                          # (C*A, K) flat voxel indices
        top = torch.stack([flat // (H * D), (flat // D) % H, flat % D], dim=-1)     # :94  (C*A, K, 3)
        ext = torch.tensor([[[W, H, D]]], dtype=torch.float32, device=dev)
        rel = ((top.float() + 0.5) / ext * 2.0 - 1.0).reshape(-1, 3).contiguous()   # :95
        fv = feat_vol[m].to(dev)
        fv = (fv if fv.dtype in (torch.float16, torch.float32) else fv.float()).contiguous()
        qf2 = ops.sample_prototypes(fv, rel, feature_sampling_mode)                 # :97  (C*A*K, F) fp32
        print('resample_topk() qf2:', torch.Size((M, C_, A, K, qf2.size(-1))))
        offs = torch.arange(0, C_ * A * K + 1, K, dtype=torch.int32, device=dev)    # groups of K prototypes -> mean (:106)
        s2 = similarity_maps(fv, qf2.contiguous(), offs, mode="clamp_mean", exponent=similarity_exponent)
        out.append(s2.view(C_, A, W, H, D))
    res = torch.stack(out)
    print('resample_topk() sims:', torch.Size((M, C_, A, K, W, H, D)))
    return res.to(src_dtype).to(src_dev)


def take_most_dissimilar(features, num_prototypes=35, measure='cosine'):
    """infer.py:108-126: the `num_prototypes` rows of `features` (N, F) with the largest mean distance to all rows
    (1 - mean cosine similarity, or mean Euclidean distance).  Distances are computed natively; the final top-k over
    N scalars is torch glue like in the reference (its `sorted=False` leaves the row ORDER unspecified)."""
    if features.size(0) <= num_prototypes: return features
    if measure not in ('cosine', 'euclidean'):
        raise ValueError(f'Unknown measure: {measure}')
    dev = _cuda_device(features.device if features.is_cuda else None)
    dist = ops.mean_pairwise_distance(features.to(dev, torch.float32).contiguous(), measure)
    largest_dists, selected = torch.topk(dist, num_prototypes, largest=True, sorted=False)
    print(f'Smallest distances (min: {largest_dists.min().item():.4f} avg: {largest_dists.mean().item():.4f}) vs average distance ({dist.mean().item():.4f})')
    return features[selected.to(features.device)]


def image_sizes(vol_shape, patch_size, feature_output_size):
    """infer.py:317-319"""
    ref_fact = sorted(vol_shape[-3:])[1] / feature_output_size
    im_sz = tuple(map(lambda d: int(patch_size * (d // ref_fact)), vol_shape[-3:]))
    return im_sz, tuple(map(lambda d: d // patch_size, im_sz))


def _pool_target(pool_fn, axis, n_slices, f_sz3):
    """Number of output slabs along the slice axis if `pool_fn` can be done natively, else None."""
    if pool_fn is _noop:
        return n_slices
    if isinstance(pool_fn, torch.nn.AdaptiveAvgPool3d):
        size = pool_fn.output_size
        size = (size,) * 3 if isinstance(size, int) else tuple(size)
        r, c = AXIS_IMAGE_DIMS[axis]
        s = AXIS_SLICE_DIM[axis]
        if size[r] in (None, f_sz3[r]) and size[c] in (None, f_sz3[c]):
            return n_slices if size[s] is None else size[s]
    return None


def k_features_axis_device(vol_dev, engine, im_sizes, slice_along, batch_size, n_out, mm=None, rank=0, world=1, compact=False):
    """Device-resident core of compute_qkv: per-axis K features pooled to `n_out` slabs along the slice
    axis, in the reference layout (D, ., ., .).  With world > 1 this rank evaluates only the slices of
    its slab range (vittf_b200/dist.py); the returned buffer is full-size and zero elsewhere, or with
    compact=True holds this rank's slabs only (the block the all-gather sends)."""
    from . import dist
    r, c = AXIS_IMAGE_DIMS[slice_along]
    s_dim = AXIS_SLICE_DIM[slice_along]
    im0, im1 = im_sizes[r], im_sizes[c]
    p = engine.patch
    f0, f1 = im0 // p, im1 // p
    S = vol_dev.shape[s_dim]
    if mm is None:
        mm = ops.minmax(vol_dev)
    o0, o1 = dist.slab_range(n_out, world, rank)
    a, b = dist.slices_for_slabs(S, n_out, o0, o1)
    ext = o1 - o0 if compact else n_out
    shape = {"z": (engine.embed_dim, f0, f1, ext), "y": (engine.embed_dim, f0, ext, f1),
             "x": (engine.embed_dim, ext, f0, f1)}[slice_along]
    alloc = torch.empty if (world == 1 or compact) else torch.zeros
    out = alloc(shape, dtype=torch.float16, device=vol_dev.device)
    if b > a:
        kbuf = torch.empty(b - a, f0 * f1, engine.embed_dim, dtype=torch.float16, device=vol_dev.device)
        for s0 in range(a, b, batch_size):
            s1 = min(b, s0 + batch_size)
            engine.k_features(vol_dev, slice_along, s0, s1, im0, im1, mm, out=kbuf[s0 - a:s1 - a])
        ops.pool_axis(kbuf, f0, f1, slice_along, n_out, out=out, total_slices=S, slice0=a, slabs=(o0, o1), compact=compact)
    return out


def compute_qkv(vol, model, patch_size, im_sizes, pool_fn=_noop, batch_size=1, slice_along='z', return_keys=['q', 'k', 'v'],
                dev=None, typ=torch.float32):
    """infer.py:130-210.  Returns {'k': fp16 CPU tensor (D, ., ., .)} in the reference's layout.

    `model` is any module with the DINO attribute layout; its parameters are read once and its
    forward is never run.  `dev` must be a CUDA device (default: current); `typ` is accepted for
    signature compatibility (compute is bf16 with fp32 accumulation)."""
    if isinstance(return_keys, str): return_keys = [return_keys]
    unsupported = [k for k in return_keys if k != 'k']
    if unsupported:
        raise NotImplementedError(f"vittf_b200.compute_qkv produces only 'k' natively (requested {unsupported}); "
                                  "the reference's callers use return_keys='k' (infer.py:326,331)")
    dev = _cuda_device(dev if (dev is not None and torch.device(dev).type == "cuda") else None)
    engine = engine_for(model, dev, max_batch=max(1, batch_size), max_tokens=_max_tokens(im_sizes, patch_size))
    if engine.patch != patch_size:
        raise ValueError(f"patch_size {patch_size} does not match the model's patch embedding ({engine.patch})")
    v = vol.squeeze()
    if v.ndim != 3:
        raise ValueError(f"compute_qkv expects a 3-D volume, got shape {tuple(vol.shape)}")
    v = v.to(dev)
    if v.dtype not in (torch.uint8, torch.float16, torch.float32):
        v = v.float()                                            # infer.py:137 `vol.float()`
    v = v.contiguous()
    f_sz3 = tuple(d // patch_size for d in im_sizes)
    S = v.shape[AXIS_SLICE_DIM[slice_along]]
    n_out = _pool_target(pool_fn, slice_along, S, f_sz3)
    with torch.no_grad():
        k = k_features_axis_device(v, engine, im_sizes, slice_along, batch_size, S if n_out is None else n_out)
        if n_out is None:                                        # exotic pool_fn: apply it as given
            k = pool_fn(k)
    return {'k': k.cpu()}


def _max_tokens(im_sizes, patch):
    f = sorted(d // patch for d in im_sizes)
    return 1 + f[-1] * f[-2]


def feature_volume(vol, model, patch_size=8, feature_output_size=64, batch_size=8, dev=None, slice_along='all',
                   rank=0, world=1, group=None):
    """The 3-axis loop of infer.py:327-333 with everything kept on the device: returns the merged fp16
    feature volume (D, fX, fY, fZ) as a CUDA tensor (z, then y, then x summed in fp16).  With world > 1
    the slices of every axis are sharded over the ranks; one all-gather per axis (asynchronous: it runs under the
    next axis' ViT compute) assembles the replicated result, un-permuted and summed by a native kernel -- or one
    all-reduce per axis when the pooled slabs do not divide evenly over the ranks."""
    from . import dist
    dev = _cuda_device(dev)
    v = vol.squeeze().to(dev)
    if v.dtype not in (torch.uint8, torch.float16, torch.float32):
        v = v.float()
    v = v.contiguous()
    im_sz, f_sz = image_sizes(tuple(v.shape), patch_size, feature_output_size)
    engine = engine_for(model, dev, max_batch=batch_size, max_tokens=_max_tokens(im_sz, patch_size))
    mm = ops.minmax(v)
    out = None
    axes = ['z', 'y', 'x'] if slice_along == 'all' else [slice_along]
    pending = None                                        # (work, staging, axis) of the exchange in flight

    def finish(pending, out):
        work, staging, ax = pending
        work.wait()                                        # the current stream waits for the collective
        if out is None:
            out = torch.empty((staging.shape[1],) + tuple(full_shape[ax]), dtype=torch.float16, device=dev)
            return ops.accumulate_gathered(out, staging, ax, accumulate=False)
        return ops.accumulate_gathered(out, staging, ax, accumulate=True)

    full_shape = {}
    with torch.no_grad():
        for ax in axes:
            s_dim = AXIS_SLICE_DIM[ax]
            n_out = f_sz[s_dim] if slice_along == 'all' else v.shape[s_dim]
            r_, c_ = AXIS_IMAGE_DIMS[ax]
            fs = [0, 0, 0]
            fs[r_], fs[c_], fs[s_dim] = im_sz[r_] // patch_size, im_sz[c_] // patch_size, n_out
            full_shape[ax] = tuple(fs)
            if dist.even_slabs(n_out, world):
                block = k_features_axis_device(v, engine, im_sz, ax, batch_size, n_out, mm=mm, rank=rank, world=world, compact=True)
                if pending is not None:                   # the previous axis' exchange ran under this axis' ViT compute
                    out = finish(pending, out)
                staging, work = dist.gather_blocks(block, group, async_op=True)
                pending = (work, staging, ax)
                continue
            part = k_features_axis_device(v, engine, im_sz, ax, batch_size, n_out, mm=mm, rank=rank, world=world)
            if pending is not None:
                out = finish(pending, out)
                pending = None
            if world > 1:
                dist.all_reduce_disjoint(part, group)
            out = part if out is None else ops.accumulate_f16(out, part)
        if pending is not None:
            out = finish(pending, out)
    return out


# ---- file handling, infer.py:212-288 ----------------------------------------------------------------
def load_data(data_path):
    data_path = Path(data_path)
    if not data_path.exists():
        print(f'Invalid argument for --data-path (File does not exist): {data_path}')
        sys.exit(1)
    print(f'Attempting to load {data_path}.')
    if data_path.suffix in ['.pt', '.pth']:
        data = torch.load(data_path, weights_only=False)
        vol = data['vol'] if type(data) == dict else data
    elif data_path.suffix == '.npy':
        data = np.load(data_path, allow_pickle=True)
        vol = torch.from_numpy((data[()]['vol'] if data.dtype == "O" else data).astype(np.float32))
    else:
        print(f'Unsupported file extension: {data_path.suffix}')
        sys.exit(1)
    print(f'Loaded volume: {vol.shape} of type {vol.dtype}.')
    assert vol.ndim == 3
    return vol


def load_model(args):
    """infer.py:239-264: DINO v1 (patch 8 / 16) and DINOv2 (patch 14) backbones."""
    if not args.dino_model and not args.dino2_model:
        print('No DINO/DINOv2 model specified, using default: vits8')
        args.dino_model = 'vits8'
    elif args.dino_model and args.dino2_model:
        print('Both --dino-model and --dino2-model were set. Please only set one of them.')
        sys.exit(1)
    from .dino import ARCHS_V2, build_dino
    if args.dino2_model:                                   # infer.py:254-260 (as intended: the reference's branch has a typo, :258)
        if args.dino2_model not in ARCHS_V2:
            print(f'DINOv2 {args.dino2_model} (SwiGLU feed-forward, 1536-d) is not built; available: {sorted(ARCHS_V2)}.')
            sys.exit(1)
        args.dino_model = args.dino2_model
        args.model = args.dino2_model
        return args.dino2_model, build_dino, 14
    args.model = args.dino_model
    return args.dino_model, build_dino, 8 if args.dino_model[-1] == '8' else 16


def handle_output_path(args):
    data_path = Path(args.data_path)
    if not args.cache_path:
        args.cache_path = data_path.parent / f'{data_path.stem}_{args.model.replace("/", "_")}_{args.slice_along}_features{args.feature_output_size}{data_path.suffix}'
    cache_path = Path(args.cache_path)
    if cache_path.exists() and not args.overwrite:
        print(f'Cache file already exists: {cache_path}. Use --overwrite to overwrite.')
        sys.exit(1)
    if not os.access(os.path.dirname(str(args.cache_path)) or os.getcwd(), os.W_OK):
        print(f'Invalid argument for --cache-path (Cannot write to location): {args.cache_path}')
        sys.exit(1)
    return cache_path


def main(argv=None):
    from argparse import ArgumentParser
    dino_archs = ['vits16', 'vits8', 'vitb16', 'vitb8']
    dino2_archs = ['vits14', 'vitb14', 'vitl14', 'vitg14']
    parser = ArgumentParser('Infer DINO features from saved volume')
    parser.add_argument('--data-path', type=str, required=True, help='Path to the saved volume')
    parser.add_argument('--cache-path', type=str, default=None, help='Path to save computed qkv features to.')
    parser.add_argument('--dino-model', type=str, choices=dino_archs, default=None, help='DINO model to use')
    parser.add_argument('--dino2-model', type=str, choices=dino2_archs, default=None, help='DINOv2 model to use')
    parser.add_argument('--slice-along', type=str, choices=['x', 'y', 'z', 'all'], default='all', help='Along which axis to slice volume, as it is fed slice-wise to DINO')
    parser.add_argument('--batch-size', type=int, default=1, help='Feed volume through network in batches')
    parser.add_argument('--feature-output-size', type=int, default=64, help='Produces a features map with aspect ratio of input volume with this value as y resolution. Only if --slice-along ALL')
    parser.add_argument('--cpu', action='store_true', help='Use CPU only')
    parser.add_argument('--overwrite', action='store_true', help='Overwrite existing cache files')
    parser.add_argument('--weights', type=str, default=None, help='[vittf_b200] DINO state dict (.pth); random init if omitted')
    parser.add_argument('--seed', type=int, default=0, help='[vittf_b200] seed of the random-init weights')
    args = parser.parse_args(argv)

    if args.cpu or not torch.cuda.is_available():
        print('vittf_b200 has no CPU path: a CUDA device (B200) is required.')
        sys.exit(1)
    dev = torch.device('cuda', torch.cuda.current_device())
    dino_model, dino_model_fn, patch_size = load_model(args)
    cache_path = handle_output_path(args)

    with torch.no_grad():
        vol = load_data(args.data_path)
        im_sz, feat_out_sz = image_sizes(tuple(vol.shape), patch_size, args.feature_output_size)
        print(f'Input image size: {im_sz}')
        model = dino_model_fn(dino_model, weights=args.weights, seed=args.seed)
        t0 = time.time()
        if args.slice_along in ['x', 'y', 'z']:
            qkv = compute_qkv(vol, model, patch_size, im_sz, batch_size=args.batch_size, return_keys='k', slice_along=args.slice_along, dev=dev)
        else:
            qkv = defaultdict(float)
            qkv['k'] = feature_volume(vol, model, patch_size, args.feature_output_size, args.batch_size, dev).cpu()
            print('k', ':', qkv['k'].shape)
        print(f'Computed qkv along {args.slice_along} in {time.time() - t0}s, saving now to: {cache_path}')
        if cache_path.suffix in ['.pt', '.pth']:
            torch.save(qkv, cache_path)
        elif cache_path.suffix == '.npy':
            np.save(cache_path, {k: v.numpy() for k, v in qkv.items()})
    return 0


if __name__ == '__main__':
    sys.exit(main())
