"""Stage 1 on the GPU: streaming kernels, the engine, and the feature volume against the oracle and
the golden vectors produced by the reference's own infer.py (tolerance from north_star:
cosine >= 0.995 per patch vs the fp32 reference)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_library_targets_this_gpu():
    from vittf_b200 import ops
    assert ops.device_arch() == 100, "libvittf_b200 is built for sm_100a (B200) only"


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float16, torch.float32])
def test_minmax(dtype):
    from vittf_b200 import ops
    g = torch.Generator().manual_seed(0)
    v = (torch.rand(37, 41, 43, generator=g) * 200 - (0 if dtype == torch.uint8 else 50)).to(dtype).cuda()
    mm = ops.minmax(v)
    assert mm[0].item() == v.float().min().item() and mm[1].item() == v.float().max().item()


@pytest.mark.parametrize("D", [384, 768])
def test_layernorm(D):
    from vittf_b200 import ops
    x = torch.randn(1001, D, device="cuda") * 3 + 1
    w, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    ref = F.layer_norm(x, (D,), w, b, eps=1e-6)
    assert ((ops.layernorm(x, w, b).float() - ref).abs() / (1.0 + ref.abs())).max().item() < 8e-3   # bf16 output


@pytest.mark.parametrize("axis", ["z", "y", "x"])
@pytest.mark.parametrize("shape,n_out", [((8, 4, 6), 4), ((7, 4, 6), 3), ((6, 4, 6), 6), ((20, 5, 7), 8), ((64, 9, 5), 16),
                                         ((37, 4, 9), 24),
                                         # more pooled slabs than slices: AdaptiveAvgPool3d replicates slices
                                         ((4, 4, 6), 8), ((5, 3, 4), 16), ((7, 4, 4), 24)])
def test_pool_axis_matches_adaptive_avg_pool(axis, shape, n_out):
    from vittf_b200 import ops
    S, f0, f1 = shape
    D = 64
    k = torch.randn(S, f0 * f1, D, device="cuda").half()
    out = ops.pool_axis(k, f0, f1, axis, n_out)
    kk = k.view(S, f0, f1, D)
    perm = {"z": (3, 1, 2, 0), "y": (3, 1, 0, 2), "x": (3, 0, 1, 2)}[axis]
    unpooled = kk.permute(*perm).float()
    tgt = {"z": (f0, f1, n_out), "y": (f0, n_out, f1), "x": (n_out, f0, f1)}[axis]
    ref = F.adaptive_avg_pool3d(unpooled, tgt).half()
    assert torch.equal(out, ref)
    acc = ops.pool_axis(k, f0, f1, axis, n_out, out=out.clone(), accumulate=True)
    assert torch.equal(acc, (ref + ref))
    if n_out >= 16:
        # sharded use: this "rank" holds only the slices of slabs [8, 12) and writes only those (partial 8-slab block)
        lo, hi = (8 * S) // n_out, -((-12 * S) // n_out)
        part = torch.zeros_like(out)
        ops.pool_axis(k[lo:hi].contiguous(), f0, f1, axis, n_out, out=part, total_slices=S, slice0=lo, slabs=(8, 12))
        sl = {"z": (..., slice(8, 12)), "y": (slice(None), slice(None), slice(8, 12)), "x": (slice(None), slice(8, 12))}[axis]
        want = torch.zeros_like(ref)
        want[sl] = ref[sl]
        assert torch.equal(part, want)


@pytest.mark.parametrize("axis", ["z", "y", "x"])
@pytest.mark.parametrize("shape,world", [((16, 8, 16), 2), ((6, 9, 12), 3), ((64, 64, 64), 8), ((5, 7, 8), 1)])
def test_gathered_merge_matches_the_layout_rule(axis, shape, world):
    """Multi-GPU merge: pooling into a rank's compact block + the un-permuting fp16 accumulate of the
    all-gathered blocks == pooling into the full array, bit-exactly (vector and scalar paths, assign and add)."""
    from vittf_b200 import dist, ops
    f0, f1, n_out = shape
    if n_out % world:
        pytest.skip("uneven slabs use the all-reduce path")
    D, S = 64, 2 * n_out
    k = torch.randn(S, f0 * f1, D, device="cuda").half()
    full = ops.pool_axis(k, f0, f1, axis, n_out)
    blocks = []
    for r in range(world):
        o0, o1 = dist.slab_range(n_out, world, r)
        a, b = dist.slices_for_slabs(S, n_out, o0, o1)
        blocks.append(ops.pool_axis(k[a:b].contiguous(), f0, f1, axis, n_out, total_slices=S, slice0=a, slabs=(o0, o1), compact=True))
    staging = torch.stack(blocks)
    assert torch.equal(dist.unpermute_gathered(staging, axis), full)
    out = torch.empty_like(full)
    ops.accumulate_gathered(out, staging, axis, accumulate=False)
    assert torch.equal(out, full)
    base = torch.randn_like(full)
    want = (base.float() + full.float()).half()
    ops.accumulate_gathered(base, staging, axis, accumulate=True)
    assert torch.equal(base, want)


def _oracle_tokens(model, vol, axis, im_sz):
    from oracle import feature_volume as fv
    imgs = fv.slice_images(vol, axis)
    r, c = fv.AXIS_IMAGE_DIMS[axis]
    return model.prepare_tokens(F.interpolate(imgs, size=(im_sz[r], im_sz[c]), mode="nearest"))


@pytest.mark.parametrize("axis", ["z", "y", "x"])
@pytest.mark.parametrize("dtype,shape,fos,tol", [(torch.uint8, (40, 32, 24), 8, 1e-4), (torch.float16, (40, 32, 24), 8, 1e-4),
                                                 (torch.float32, (40, 32, 24), 8, 1e-4), (torch.uint8, (72, 80, 96), 72, 1e-4)])
def test_patch_embed_matches_prepare_tokens(axis, dtype, shape, fos, tol):
    """uint8 / fp16 volumes enter the tensor-core patch embed as exact fp16 operands (1e-4 like the fp32 FMA kernel); fp32
    volumes are rounded to fp16 after min-max normalisation (<= 2^-12 per grey value).  The last case has more than 64
    patches per image row (two 64-patch chunks per item)."""
    from oracle import dino_vit, feature_volume as fv, synth
    from vittf_b200 import ops
    from vittf_b200.vit import fold_patch_embed, interpolate_pos_embed
    vol, _ = synth.ct_volume(shape, n_shells=4, seed=3)
    if dtype != torch.uint8:
        g = torch.Generator().manual_seed(1)
        vol = (vol.float() / 255.0 * 0.9 + 0.1 * torch.rand(shape, generator=g)).to(dtype)
    model = dino_vit.build("vits8", depth=1)
    im_sz, _ = fv.image_sizes(tuple(vol.shape), 8, fos)
    ref = _oracle_tokens(model, vol, axis, im_sz)
    r, c = fv.AXIS_IMAGE_DIMS[axis]
    pw, pb = fold_patch_embed(model.patch_embed.proj.weight, model.patch_embed.proj.bias)
    pos = interpolate_pos_embed(model.pos_embed, model.cls_token, 8, im_sz[r], im_sz[c])
    v = vol.cuda()
    out = ops.patch_embed(v, axis, 0, ref.shape[0], im_sz[r], im_sz[c], 8, ops.minmax(v), pw.cuda(), pb.cuda(), pos.cuda())
    assert (out.cpu() - ref).abs().max().item() < tol


def _cos_min(a, b):
    a = a.float().flatten(1).t()
    b = b.float().flatten(1).t()
    return F.cosine_similarity(a, b, dim=-1).min().item()


@pytest.mark.parametrize("name", ["feat_cube", "feat_noncubic"])
def test_feature_volume_matches_reference_golden(golden, name):
    from oracle import dino_vit, synth
    from vittf_b200 import infer
    g = golden(name)
    shape, fos, depth = tuple(int(v) for v in g["shape"]), int(g["fos"]), int(g["depth"])
    vol, _ = synth.ct_volume(shape, n_shells=4, seed=int(g["seed_vol"]))
    model = dino_vit.build("vits8", seed=int(g["seed_model"]), depth=depth)
    ref = torch.from_numpy(g["k"])
    out = infer.feature_volume(vol, model, 8, fos, batch_size=3).cpu()
    assert out.dtype == torch.float16 and out.shape == ref.shape
    assert _cos_min(out, ref) >= 0.995
    # drop-in signature, single axis, un-pooled (infer.py:326)
    im_sz = tuple(int(v) for v in g["im_sz"])
    ky = infer.compute_qkv(vol, model, 8, im_sz, batch_size=2, return_keys="k", slice_along="y")["k"]
    refy = torch.from_numpy(g["k_y_unpooled"])
    assert ky.shape == refy.shape and ky.dtype == torch.float16 and not ky.is_cuda
    assert _cos_min(ky, refy) >= 0.995
    # 3-axis loop through the drop-in exactly as infer.py:327-333 writes it
    pool = torch.nn.AdaptiveAvgPool3d(tuple(d // 8 for d in im_sz))
    acc = 0.0
    for ax in ["z", "y", "x"]:
        v = infer.compute_qkv(vol, model, 8, im_sz, pool_fn=pool, batch_size=4, return_keys="k", slice_along=ax)["k"]
        acc = torch.as_tensor(acc) + v.squeeze().half()
    assert torch.equal(acc, out)


def test_vit_full_depth_512(golden):
    """ViT-S/8, 12 blocks, 512^2 input (4097 tokens, the benchmark shape): two slices vs the fp32 oracle."""
    from oracle import dino_vit, feature_volume as fv, synth
    from vittf_b200 import ops
    from vittf_b200.vit import engine_for
    model = dino_vit.build("vits8", seed=0)
    vol, _ = synth.ct_volume((128, 128, 2), n_shells=6, seed=1)
    imgs = F.interpolate(fv.slice_images(vol, "z"), size=(512, 512), mode="nearest")
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = fv.hooked_qkv(model, imgs)[:, 1:, 384:768].float()                      # (2, 4096, 384)
    eng = engine_for(model, torch.device("cuda", 0), max_batch=2)
    v = vol.cuda()
    out = eng.k_features(v, "z", 0, 2, 512, 512, ops.minmax(v)).float().cpu()
    cos = F.cosine_similarity(out, ref, dim=-1)
    assert cos.min().item() >= 0.995, cos.min().item()
    assert (out - ref).abs().max().item() < 0.05 * ref.abs().max().item()


def test_feature_volume_smaller_than_feature_output_size():
    """--slice-along all on a volume whose median extent is below --feature-output-size (32^3 with the default 64):
    the reference pools S = 32 slices to 64 slabs (AdaptiveAvgPool3d replicates); same result here (bit-exact pooling of
    the same K features, cosine vs the fp32 oracle)."""
    from oracle import dino_vit, feature_volume as ofv, synth
    from vittf_b200 import infer
    vol, _ = synth.ct_volume((32, 32, 32), n_shells=4, seed=3)
    model = dino_vit.build("vits8", seed=0, depth=2)
    ref = ofv.feature_volume(vol, model, patch=8, fos=64, batch_size=8)
    out = infer.feature_volume(vol, model, 8, 64, batch_size=8).cpu()
    assert out.shape == ref.shape == (384, 64, 64, 64)
    assert _cos_min(out, ref) >= 0.995


def test_minmax_and_norm_minmax_accept_offset_views():
    """A contiguous view with a storage offset (vol[1:]) is not 16-byte aligned; fp16 input keeps its dtype (infer.py:32-34)."""
    from vittf_b200 import infer, ops
    g = torch.Generator().manual_seed(0)
    for dt in (torch.uint8, torch.float16, torch.float32):
        v = (torch.rand(5, 7, 9, generator=g) * 200).to(dt).cuda()
        view = v.flatten()[1:]
        mm = ops.minmax(view)
        assert mm[0].item() == view.float().min().item() and mm[1].item() == view.float().max().item()
        n = infer.norm_minmax(view)
        assert n.dtype == (torch.float16 if dt == torch.float16 else torch.float32)
        ref = (view.float() - view.float().min()) / (view.float().max() - view.float().min())
        assert (n.float() - ref).abs().max().item() < (1e-3 if dt == torch.float16 else 1e-6)


def test_vitb8_full_depth_512():
    """The metric's own backbone (BASELINE.json configs[2]; reference call site infer.py:176-177 with --dino-model vitb8,
    infer.py:303): ViT-B/8, all 12 blocks, 512^2 input (4097 tokens), two slices vs the fp32 oracle.  Tolerance from
    north_star: cosine >= 0.995 per patch."""
    from oracle import dino_vit, feature_volume as fv, synth
    from vittf_b200 import ops
    from vittf_b200.vit import engine_for
    model = dino_vit.build("vitb8", seed=0)
    vol, _ = synth.ct_volume((512, 512, 2), n_shells=16, seed=1)                  # native 512^2 slices: resize factor 1, as at cfg3
    imgs = fv.slice_images(vol, "z")
    ref = fv.hooked_qkv(model, imgs)[:, 1:, 768:1536].float()                     # (2, 4096, 768)
    eng = engine_for(model, torch.device("cuda", 0), max_batch=2)
    v = vol.cuda()
    out = eng.k_features(v, "z", 0, 2, 512, 512, ops.minmax(v)).float().cpu()
    cos = F.cosine_similarity(out, ref, dim=-1)
    assert cos.min().item() >= 0.995, cos.min().item()
    assert (out - ref).abs().max().item() < 0.05 * ref.abs().max().item()


def test_feature_volume_full_size_properties():
    """BASELINE.json configs[1] (256^3 uint8, ViT-S/8, 12 blocks, 768 slice images of 512^2): size-independent properties of
    stage 1 -- the feature volume does not depend on the slice batch size (bit-exact: slices are independent units), nor
    on how the slices are sharded (two emulated ranks write disjoint slabs whose sum is the un-sharded volume, bit-exact),
    every value is finite, and pooled slabs sampled along each axis match the fp32 CPU oracle on those slices (cos >= 0.995)."""
    from oracle import dino_vit, feature_volume as fv, synth
    from vittf_b200 import infer, ops
    from vittf_b200.vit import engine_for
    dev = torch.device("cuda", 0)
    vol, _ = synth.ct_volume((256, 256, 256), n_shells=8, seed=0)
    model = dino_vit.build("vits8", seed=0)
    v = vol.to(dev)
    a = infer.feature_volume(v, model, 8, 64, batch_size=64)
    assert a.shape == (384, 64, 64, 64) and a.dtype == torch.float16 and torch.isfinite(a.float()).all()
    b = infer.feature_volume(v, model, 8, 64, batch_size=24)
    assert torch.equal(a, b)
    # sharding: per axis the ranks' zero-initialised partial volumes have disjoint supports; their sum is the axis volume
    im_sz, f_sz = infer.image_sizes((256, 256, 256), 8, 64)
    eng = engine_for(model, dev, max_batch=32)
    mm = ops.minmax(v)
    acc = None
    for ax in ["z", "y", "x"]:
        parts = [infer.k_features_axis_device(v, eng, im_sz, ax, 32, 64, mm=mm, rank=r, world=2) for r in range(2)]
        assert ((parts[0] != 0) & (parts[1] != 0)).sum().item() == 0
        part = parts[0] + parts[1]
        acc = part if acc is None else ops.accumulate_f16(acc, part)
        if ax == "z":
            # pooled slab 17 of the z pass = mean of slices 68..71: against the fp32 oracle on those four slices
            imgs = F.interpolate(fv.slice_images(vol, "z")[68:72], size=(512, 512), mode="nearest")
            ref = fv.hooked_qkv(model, imgs)[:, 1:, 384:768].float().mean(0)          # (4096, 384)
            got = part[:, :, :, 17].float().reshape(384, -1).t().cpu()
            assert F.cosine_similarity(got, ref, dim=-1).min().item() >= 0.995
    assert torch.equal(acc, a)


@pytest.mark.parametrize("arch,depth", [("vits16", 3), ("vitb16", 2), ("vitb8", 2), ("vits14", 3), ("vitb14", 2), ("vitl14", 2)])
def test_other_backbones_match_oracle(arch, depth):
    """--dino-model vits16 / vitb16 (patch 16: the 224^2 position grid is 14 x 14 and is always
    bicubically resampled), vitb8 (768-d, 12 heads) and the DINOv2 backbones --dino2-model vits14 / vitb14 / vitl14
    (patch 14, 37 x 37 position grid, LayerScale folded into proj / fc2, 1024-d x 16 heads for L) through the same
    engine, against the fp32 oracle."""
    from oracle import dino_vit, feature_volume as ofv, synth
    from vittf_b200 import infer
    patch = int(arch[4:])
    vol, _ = synth.ct_volume((48, 40, 32), n_shells=4, seed=2)
    model = dino_vit.build(arch, seed=1, depth=depth)
    ref = ofv.feature_volume(vol, model, patch=patch, fos=8, batch_size=4)
    out = infer.feature_volume(vol, model, patch, 8, batch_size=5).cpu()
    assert out.shape == ref.shape and out.dtype == torch.float16
    assert _cos_min(out, ref) >= 0.995


@pytest.mark.parametrize("switch,select", [
    ("VITTF_NO_LNFOLD", "reference_golden or full_depth or other_backbones or (test_gemm and not test_gemm_ln)"),
    ("VITTF_GEMM_NO_PAIRS", "reference_golden or full_depth or other_backbones or (test_gemm and not test_gemm_ln)"),
    ("VITTF_ATTN_SAFE_ONLY", "reference_golden or full_depth"),
    ("VITTF_PE_NO_MMA", "reference_golden or patch_embed or full_depth_512"),
    ("VITTF_GEMM_EPI16", "test_gemm and not test_gemm_ln"),
], ids=["VITTF_NO_LNFOLD", "VITTF_GEMM_NO_PAIRS", "VITTF_ATTN_SAFE_ONLY", "VITTF_PE_NO_MMA", "VITTF_GEMM_EPI16"])
def test_engine_switches(switch, select):
    """The engine's A/B switches select code the default path does not run: VITTF_NO_LNFOLD = separate LayerNorm kernel +
    reduce-add residual epilogue for ViT-B / ViT-L too (the default folds norm1 / norm2 into the GEMMs around them),
    VITTF_GEMM_NO_PAIRS = single-CTA GEMM tiles (the default runs N % 256 == 0 as CTA pairs, tcgen05 cta_group::2),
    VITTF_ATTN_SAFE_ONLY = the online-softmax attention kernel alone (default: max-free first pass + safe pass over flagged
    items), VITTF_PE_NO_MMA = the FMA-pipe patch embedding for patch 8, VITTF_GEMM_EPI16 = 16 instead of 8 epilogue warps for the bf16
    epilogues of the CTA-pair GEMM tiles.  The golden feature volumes, the full-depth
    ViT-S/8 and ViT-B/8 checks and the unit tests of the switched kernels must hold under each."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    here = Path(__file__).resolve().parent
    env = dict(os.environ)
    env[switch] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", str(here / "test_gpu_vit.py"), str(here / "test_gpu_gemm.py"), "-x", "-q", "-m", "gpu",
                        "-p", "no:cacheprovider", "-k", f"({select}) and not engine_switches"],
                       env=env, capture_output=True, text=True, cwd=str(here.parent), timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
