"""Generate tests/golden/*.npz by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python -m oracle.make_golden

The reference modules are imported unmodified from /root/reference with the
environment shims listed in SURVEY.md §0.4 (missing imports / removed SciPy kwarg /
missing icecream) -- none of which changes arithmetic.  Inputs are regenerated from
seeds by oracle/synth.py inside the tests; only the reference OUTPUTS (and the
inputs that are cheap to store) are committed.
"""
import sys
import types
import warnings
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def import_reference():
    sys.path.insert(0, REF)
    ice = types.ModuleType("icecream")
    ice.ic = lambda *a, **k: None
    sys.modules["icecream"] = ice
    import infer
    import bilateral_solver3d
    from torchvision.transforms.functional import normalize
    from scipy.sparse.linalg import cg as scipy_cg
    infer.normalize = normalize                                   # infer.py:155 vs :293
    bilateral_solver3d.F = torch.nn.functional                    # bilateral_solver3d.py:176-181
    bilateral_solver3d.cg = lambda A, b, x0=None, M=None, maxiter=None, tol=1e-5: scipy_cg(
        A, b, x0=x0, M=M, maxiter=maxiter, rtol=tol)              # SciPy >= 1.14 renamed tol
    import predict_ntf
    return infer, predict_ntf, bilateral_solver3d


def reference_feature_volume(infer, vol, model, patch, fos, batch_size):
    """infer.py:314-333 verbatim control flow (the __main__ body is not importable)."""
    ref_fact = sorted(vol.shape[-3:])[1] / fos
    im_sz = tuple(map(lambda d: int(patch * (d // ref_fact)), vol.shape[-3:]))
    feat_out_sz = tuple(map(lambda d: d // patch, im_sz))
    qkv = defaultdict(float)
    avg_pool = torch.nn.AdaptiveAvgPool3d(output_size=feat_out_sz)
    dev, typ = torch.device("cpu"), torch.float32
    for ax in ["z", "y", "x"]:
        for k, v in infer.compute_qkv(vol, model, patch, im_sz, pool_fn=avg_pool, batch_size=batch_size,
                                      return_keys="k", slice_along=ax, dev=dev, typ=typ).items():
            qkv[k] = (torch.as_tensor(qkv[k]).to(dev) + v.to(dev).squeeze().half()).cpu()
    return qkv["k"], im_sz


def golden_sampling():
    """compare_feat_sampling.py:13-33 under torch.manual_seed: index sets the drop-in must reproduce."""
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import synth
    import_reference()
    import compare_feat_sampling as cfs
    lab = synth.shell_labels((30, 26, 22), 3)
    out = {"shape": np.array(lab.shape)}
    for cls in (0, 1, 2):
        mask = (lab == cls).numpy()
        torch.manual_seed(100 + cls)
        out[f"uniform_{cls}"] = cfs.sample_uniform(mask, 12).numpy()
        torch.manual_seed(200 + cls)
        out[f"surface_{cls}"] = cfs.sample_surface(mask, 10, dist_from_surface=4).numpy()
        torch.manual_seed(300 + cls)
        out[f"both_{cls}"] = cfs.sample_both(mask, 16, dist_from_surface=4).numpy()
        torch.manual_seed(400 + cls)
        out[f"surface_all_{cls}"] = cfs.sample_surface(mask, 10 ** 6, dist_from_surface=2).numpy()   # "< n_samples" branch
    np.savez_compressed(OUT / "sampling.npz", **out)
    print("sampling", {k: v.shape for k, v in out.items()})


def golden_refine():
    """infer.py:75-126 (resample_topk, take_most_dissimilar) on a small class-structured feature volume."""
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import synth
    infer, _, _ = import_reference()
    lr = (12, 10, 8)
    feats, protos = synth.class_features(32, lr, 3, seed=9, dtype=torch.float32)
    g = torch.Generator().manual_seed(4)
    p = torch.nn.functional.normalize(protos.repeat(2, 1) + 0.05 * torch.randn(6, 32, generator=g), dim=-1).view(3, 2, 32)
    sims = (torch.einsum("fwhd,caf->cawhd", feats, p).clamp(0, 1) ** 2.0)
    sims[0, 0] = (sims[0, 0] * 4).round() / 4                   # heavy ties: exercises the `nonzero()[:K]` rule
    out = {"sims_in": sims.numpy()}
    for K in (3, 8):
        out[f"topk{K}"] = infer.resample_topk(feats.clone()[None], sims.clone(), K=K, similarity_exponent=2.0).numpy()[0]
    f2 = torch.randn(40, 16, generator=g)
    f2[:5] *= 3.0
    for measure in ("cosine", "euclidean"):
        out[f"dissim_{measure}"] = infer.take_most_dissimilar(f2.clone(), num_prototypes=9, measure=measure).numpy()
    out["dissim_in"] = f2.numpy()
    np.savez_compressed(OUT / "refine.npz", **out)
    print("refine", {k: v.shape for k, v in out.items()})


def main():
    warnings.filterwarnings("ignore")
    if len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "sampling":
        return golden_sampling()
    if len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "refine":
        return golden_refine()
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import dino_vit, synth
    infer, predict_ntf, bls3d = import_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    golden_sampling()
    golden_refine()

    # ---------------------------------------------------------------- stage 1: feature volume
    cases = {
        # name: (volume shape, fos, depth, batch)
        "feat_cube": ((32, 32, 32), 8, 3, 4),
        "feat_noncubic": ((40, 32, 24), 8, 2, 1),
    }
    for name, (shape, fos, depth, bs) in cases.items():
        vol, _ = synth.ct_volume(shape, n_shells=4, seed=3)
        model = dino_vit.build("vits8", seed=0, depth=depth)
        k, im_sz = reference_feature_volume(infer, vol, model, 8, fos, bs)
        # single-axis, un-pooled variant (infer.py:326)
        kz = infer.compute_qkv(vol, dino_vit.build("vits8", seed=0, depth=depth), 8, im_sz, batch_size=bs,
                               return_keys="k", slice_along="y")["k"]
        np.savez_compressed(OUT / f"{name}.npz", k=k.numpy(), k_y_unpooled=kz.numpy(), im_sz=np.array(im_sz),
                            shape=np.array(shape), fos=fos, depth=depth, seed_vol=3, seed_model=0)
        print(name, tuple(k.shape), k.dtype, "im_sz", im_sz)

    # ---------------------------------------------------------------- stage 2: similarity
    vol_shape = (48, 40, 32)
    feats, _ = synth.class_features(32, (12, 10, 8), 3, seed=5, dtype=torch.float32)
    ann = synth.annotations(vol_shape, 3, 4, seed=5)
    # add border annotations: zero-padding of grid_sample matters there
    ann["ntf1"] = torch.cat([ann["ntf1"], torch.tensor([[0, 0, 0], [47, 39, 31], [0, 20, 31]])])
    sims = predict_ntf.compute_similarities(np.zeros(vol_shape, np.float32), feats.clone(), ann, bilateral_solver=False)
    pts = torch.cat(list(ann.values()))
    rel = (pts.float() + 0.5) / torch.tensor([list(vol_shape)]) * 2.0 - 1.0
    pb = infer.sample_features3d(feats, rel.clone(), mode="bilinear")[0, 0]
    pn = infer.sample_features3d(feats, rel.clone(), mode="nearest")[0, 0]
    np.savez_compressed(OUT / "sim_refntf.npz", feats=feats.numpy(), vol_shape=np.array(vol_shape),
                        ann_names=np.array(list(ann.keys())), ann_sizes=np.array([v.size(0) for v in ann.values()]),
                        ann_pts=pts.numpy(), protos_bilinear=pb.numpy(), protos_nearest=pn.numpy(),
                        **{f"sim_{k}": v.numpy() for k, v in sims.items()})
    print("sim_refntf", {k: tuple(v.shape) for k, v in sims.items()})

    # with the bilateral solver on (predict_ntf.py:73-96): needs a volume with structure
    vol_u8, _ = synth.ct_volume(vol_shape, n_shells=3, seed=5)
    sims_bls = predict_ntf.compute_similarities(vol_u8.float().numpy(), feats.clone(), ann, bilateral_solver=True)
    np.savez_compressed(OUT / "sim_refntf_bls.npz", **{f"sim_{k}": v.numpy() for k, v in sims_bls.items()})
    print("sim_refntf_bls", {k: (tuple(v.shape), int(v.max())) for k, v in sims_bls.items()})

    # ---------------------------------------------------------------- stage 3: bilateral solver
    for name, shape, sig in (("bls_s755", (40, 36, 28), dict(sigma_spatial=7, sigma_chroma=5, sigma_luma=5)),
                             ("bls_s333", (24, 24, 20), dict(sigma_spatial=3, sigma_chroma=3, sigma_luma=3)),
                             ("bls_default", (48, 48, 48), {})):
        r8, lab = synth.ct_volume(shape, n_shells=4, seed=7)
        g = torch.Generator().manual_seed(11)
        t = ((lab == 1).float() * 0.8 + 0.2 * torch.rand(shape, generator=g)).clamp(0, 1)
        out = bls3d.apply_bilateral_solver3d(t[None], r8.expand(3, -1, -1, -1), grid_params=sig)
        cexp = torch.rand((1,) + shape, generator=g)
        out_c = bls3d.apply_bilateral_solver3d(t[None], r8.expand(3, -1, -1, -1), c=cexp, grid_params=sig)
        np.savez_compressed(OUT / f"{name}.npz", out=out.numpy(), out_c=out_c.numpy(), shape=np.array(shape),
                            sig=np.array([sig.get("sigma_spatial", 24), sig.get("sigma_luma", 4), sig.get("sigma_chroma", 4)]))
        print(name, tuple(out.shape), float(out.min()), float(out.max()))

    # crop helpers (tests/test_bls_crop.py property, 3-D)
    g = torch.Generator().manual_seed(2)
    s = torch.rand((1, 9, 8, 7), generator=g) * (synth.shell_labels((9, 8, 7), 2) == 0)
    crops, (mi, ma) = bls3d.crop_pad([s, s[0] * 2], thresh=0.1, pad=2)
    np.savez_compressed(OUT / "crop.npz", s=s.numpy(), c0=crops[0].numpy(), c1=crops[1].numpy(), mi=mi.numpy(), ma=ma.numpy())


def golden_eval():
    """tests/golden/eval_metrics.json = the metrics.json written by the UNMODIFIED /root/reference/evaluate_similarities.py
    (run as a script; icecream stubbed: it only pretty-prints) for the seeded inputs of oracle.evaluate.make_inputs."""
    import json
    import runpy
    import tempfile
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from oracle import evaluate as oev
    ice = types.ModuleType("icecream")
    ice.ic = lambda *a, **k: None
    ice.ic.configureOutput = lambda **k: None

    class _Reg:
        @staticmethod
        def register(_t):
            return lambda f: f
    ice.argumentToString = _Reg
    sys.modules["icecream"] = ice
    with tempfile.TemporaryDirectory() as tmp:
        d, label_fn, names = oev.make_inputs(tmp, seed=0)
        argv = sys.argv
        sys.argv = ["evaluate_similarities.py", "--data", str(d), "--label", str(label_fn), "--labels", *names]
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                runpy.run_path(REF + "/evaluate_similarities.py", run_name="__main__")
        finally:
            sys.argv = argv
        res = json.loads((d / "metrics.json").read_text())
    (OUT / "eval_metrics.json").write_text(json.dumps(res, indent=1))
    print("eval_metrics", {k: round(v["accuracy"], 4) for k, v in res.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "eval":
        golden_eval()
    else:
        main()
