"""Stage 2 on the GPU against the golden vectors of the reference's own predict_ntf / infer code and
against the oracle (tolerance from north_star: similarity maps within 2e-3 absolute; labels >= 99.9 %
agreement and exact where the top-2 margin exceeds 1e-2)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _sim_inputs(g):
    feats = torch.from_numpy(g["feats"])
    pts = torch.from_numpy(g["ann_pts"])
    ann, off = {}, 0
    for n, s in zip([str(n) for n in g["ann_names"]], g["ann_sizes"]):
        ann[n] = pts[off:off + int(s)]
        off += int(s)
    return feats, ann, tuple(int(v) for v in g["vol_shape"])


def test_sample_features3d_matches_reference(golden):
    from vittf_b200 import infer
    from oracle import similarity as osim
    g = golden("sim_refntf")
    feats, ann, vs = _sim_inputs(g)
    rel = osim.rel_coords(torch.cat(list(ann.values())), vs)
    for mode, key in (("bilinear", "protos_bilinear"), ("nearest", "protos_nearest")):
        out = infer.sample_features3d(feats.cuda(), rel.clone(), mode=mode)
        assert out.shape == (1, 1, rel.shape[0], feats.shape[0]) and out.is_cuda
        assert (out[0, 0].cpu() - torch.from_numpy(g[key])).abs().max().item() < 1e-6
    # CPU tensors are accepted too (result comes back on the CPU), fp16 features as in `--gpu` mode
    out16 = infer.sample_features3d(feats.half(), rel.clone(), mode="nearest")
    assert not out16.is_cuda and out16.dtype == torch.float16
    assert torch.equal(out16[0, 0], torch.from_numpy(g["protos_nearest"]).half())


def test_compute_similarities_matches_reference(golden):
    from vittf_b200 import predict_ntf
    g = golden("sim_refntf")
    feats, ann, vs = _sim_inputs(g)
    import numpy as np
    out = predict_ntf.compute_similarities(np.zeros(vs, np.float32), feats.cuda(), ann, bilateral_solver=False)
    assert list(out.keys()) == list(ann.keys())
    for n, v in out.items():
        ref = torch.from_numpy(g[f"sim_{n}"])
        assert v.dtype == torch.uint8 and v.shape == ref.shape and not v.is_cuda
        d = (v.int() - ref.int()).abs()
        assert (d > 1).float().mean().item() < 1e-3, n          # one quantisation step from float reassociation
    assert predict_ntf.compute_similarities(np.zeros(vs, np.float32), feats.cuda(), {}) is None


def test_refntf_float_matches_oracle(golden):
    from oracle import similarity as osim
    from vittf_b200 import predict_ntf
    g = golden("sim_refntf")
    feats, ann, vs = _sim_inputs(g)
    ref = torch.stack(list(osim.ref_ntf_float(vs, feats, ann).values()))
    out = predict_ntf.similarity_float(vs, feats.cuda(), ann).cpu()
    assert (out - ref).abs().max().item() < TOL


def test_compute_similarities_with_bilateral_solver(golden):
    import numpy as np
    from oracle import synth
    from vittf_b200 import predict_ntf
    g = golden("sim_refntf")
    gb = golden("sim_refntf_bls")
    feats, ann, vs = _sim_inputs(g)
    vol_u8, _ = synth.ct_volume(vs, n_shells=3, seed=5)
    out = predict_ntf.compute_similarities(vol_u8.float().numpy(), feats.cuda(), ann, bilateral_solver=True)
    for n, v in out.items():
        ref = torch.from_numpy(gb[f"sim_{n}"])
        assert v.shape == ref.shape and v.dtype == torch.uint8
        d = (v.int() - ref.int()).abs()
        d = torch.minimum(d, 256 - d)                           # the reference's uint8 cast wraps above 255
        assert (d > 1).float().mean().item() < 2e-3, (n, d.max().item())


@pytest.mark.parametrize("lr,out_shape,F_,A", [((6, 5, 4), (24, 15, 20), 24, 6), ((16, 16, 16), (64, 64, 64), 96, 8),
                                               ((8, 8, 8), (8, 8, 8), 32, 3), ((12, 10, 8), (31, 29, 17), 40, 20),
                                               ((8, 8, 8), (16, 16, 16), 32, 5), ((8, 8, 8), (64, 64, 64), 48, 9),
                                               ((33, 33, 33), (132, 132, 132), 16, 4), ((6, 5, 4), (24, 20, 16), 24, 7),
                                               ((5, 4, 6), (40, 32, 48), 32, 12), ((32, 32, 32), (128, 128, 128), 64, 32),
                                               # fused tensor-core pass 1 (F % 32 == 0, d % 8 == 0) at awkward extents: odd h,
                                               # d below / not a multiple of the 64-voxel z tile, more than 32 prototypes
                                               ((5, 7, 16), (20, 28, 64), 32, 9), ((6, 5, 24), (24, 20, 96), 64, 40),
                                               ((3, 3, 72), (12, 12, 288), 32, 5), ((9, 6, 8), (72, 48, 64), 96, 33),
                                               # factor 2 on a 32-aligned grid (the configs[0] path) and a non-cubic factor-2 grid
                                               ((32, 32, 32), (64, 64, 64), 32, 6), ((10, 6, 8), (20, 12, 16), 32, 11)])
def test_ns_similarity_matches_oracle(lr, out_shape, F_, A):
    from oracle import similarity as osim, synth
    from vittf_b200.similarity import similarity_maps
    C = min(A, 3)
    feats, protos = synth.class_features(F_, lr, C, seed=2, dtype=torch.float16)
    g = torch.Generator().manual_seed(9)
    offs = [round(i * A / C) for i in range(C + 1)]
    owner = torch.tensor([c for c in range(C) for _ in range(offs[c + 1] - offs[c])])     # prototype -> its class
    p = F.normalize(protos[owner] + 0.05 * torch.randn(A, F_, generator=g), dim=-1)
    ref = osim.ns_composite(feats, p, offs, out_shape, exponent=2.0, slab=5)
    out = similarity_maps(feats.cuda(), p.cuda().contiguous(), torch.tensor(offs, dtype=torch.int32, device="cuda"),
                          out_shape, mode="ns", exponent=2.0).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() < TOL
    # z-slab evaluation (multi-GPU sharding unit) gives the same voxels
    zs = similarity_maps(feats.cuda(), p.cuda().contiguous(), torch.tensor(offs, dtype=torch.int32, device="cuda"),
                         out_shape, mode="ns", exponent=2.0, z_range=(3, out_shape[2] - 2)).cpu()
    assert torch.equal(zs, out[..., 3:out_shape[2] - 2])
    # x-slab evaluation (the slowest axis: pass 1 only evaluates the low-res planes under the slab +- 1)
    if out_shape[0] >= 8:
        xa, xb = out_shape[0] // 4, out_shape[0] - 3
        xs = similarity_maps(feats.cuda(), p.cuda().contiguous(), torch.tensor(offs, dtype=torch.int32, device="cuda"),
                             out_shape, mode="ns", exponent=2.0, x_range=(xa, xb)).cpu()
        assert torch.equal(xs, out[:, xa:xb])
        xz = similarity_maps(feats.cuda(), p.cuda().contiguous(), torch.tensor(offs, dtype=torch.int32, device="cuda"),
                             out_shape, mode="ns", exponent=2.0, x_range=(0, xa + 1), z_range=(3, out_shape[2] - 2)).cpu()
        assert torch.equal(xz, out[:, :xa + 1, :, 3:out_shape[2] - 2])
    if out_shape[2] >= 16:                                        # even slab bounds: the vectorised store path
        zs = similarity_maps(feats.cuda(), p.cuda().contiguous(), torch.tensor(offs, dtype=torch.int32, device="cuda"),
                             out_shape, mode="ns", exponent=2.5, z_range=(4, out_shape[2] - 6)).cpu()
        ref25 = osim.ns_composite(feats, p, offs, out_shape, exponent=2.5, slab=5)
        assert (zs - ref25[..., 4:out_shape[2] - 6]).abs().max().item() < TOL
    # labels: >= 99.9 % agreement, exact where the top-2 margin exceeds 1e-2
    from vittf_b200.predict_ntf import argmax_labels
    lab = argmax_labels(out.cuda()).cpu().long()
    ref_lab = osim.argmax_labels(ref)
    top2 = ref.topk(2, dim=0).values
    margin = top2[0] - top2[1]
    assert (lab == ref_lab).float().mean().item() >= 0.999
    assert torch.equal(lab[margin > 1e-2], ref_lab[margin > 1e-2])


@pytest.mark.parametrize("F_,n,N,A,C", [(384, 64, 128, 32, 4), (384, 64, 256, 32, 8), (768, 64, 512, 64, 16), (384, 128, 512, 24, 8)])
def test_ns_similarity_full_size(F_, n, N, A, C):
    """BASELINE.json shapes (configs[0] -- the factor-2 grid 64^3 -> 128^3 --, configs[1], [2], [3]): spot check of 4096 random + all-corner output voxels against the point
    oracle (2e-3), and the size-independent properties of the stage -- z-slab sharding invariance (bit-exact), invariance
    to the order of a class's prototypes (bit-exact), invariance to a power-of-two scale of the features (the NS
    composition normalises; bit-exact because every product scales exactly), values in [0, 1]."""
    from oracle import similarity as osim, synth
    from vittf_b200.similarity import similarity_maps
    feats, protos = synth.class_features(F_, (n, n, n), C, seed=4, dtype=torch.float16)
    g = torch.Generator().manual_seed(21)
    offs = [round(i * A / C) for i in range(C + 1)]
    owner = torch.tensor([c for c in range(C) for _ in range(offs[c + 1] - offs[c])])
    p = F.normalize(protos[owner] + 0.05 * torch.randn(A, F_, generator=g), dim=-1)
    fc, pc = feats.cuda(), p.cuda().contiguous()
    oc = torch.tensor(offs, dtype=torch.int32, device="cuda")
    out = similarity_maps(fc, pc, oc, (N, N, N), mode="ns", exponent=2.0)
    assert out.shape == (C, N, N, N)
    assert out.min().item() >= 0.0 and out.max().item() <= 1.0 and torch.isfinite(out).all()
    vox = torch.randint(0, N, (4096, 3), generator=g)
    edge = torch.tensor([[x, y, z] for x in (0, 1, N - 2, N - 1) for y in (0, 2, N - 1) for z in (0, 1, 3, N - 1)])
    vox = torch.cat([vox, edge])
    ref = osim.ns_at_voxels(feats, p, offs, (N, N, N), vox)
    got = out[:, vox[:, 0].cuda(), vox[:, 1].cuda(), vox[:, 2].cuda()].cpu()
    assert (got - ref).abs().max().item() < TOL
    # z-slab sharding (the multi-GPU unit): 3 uneven slabs reproduce the full volume bit-exactly
    cuts = [0, N // 4 + 2, N // 2 + 8, N]
    for z0, z1 in zip(cuts[:-1], cuts[1:]):
        zs = similarity_maps(fc, pc, oc, (N, N, N), mode="ns", exponent=2.0, z_range=(z0, z1))
        assert torch.equal(zs, out[..., z0:z1])
    del zs
    # x-slab sharding (similarity-only workloads: pass 1 shards with the maps): 3 uneven slabs, bit-exact
    cuts = [0, N // 4 + 2, N // 2 + 8, N]
    for x0, x1 in zip(cuts[:-1], cuts[1:]):
        xs = similarity_maps(fc, pc, oc, (N, N, N), mode="ns", exponent=2.0, x_range=(x0, x1))
        assert torch.equal(xs, out[:, x0:x1])
    del xs
    # prototype order inside a class
    perm = torch.cat([torch.arange(offs[c], offs[c + 1]).flip(0) for c in range(C)])
    out2 = similarity_maps(fc, pc[perm.cuda()].contiguous(), oc, (N, N, N), mode="ns", exponent=2.0)
    assert torch.equal(out2, out)
    # power-of-two feature scale
    out2 = similarity_maps((fc * 4).contiguous(), pc, oc, (N, N, N), mode="ns", exponent=2.0)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("lr,out_shape", [((8, 8, 8), (32, 32, 32)), ((8, 8, 8), (64, 64, 64)), ((6, 5, 4), (24, 15, 20))])
def test_ns_similarity_with_empty_classes(lr, out_shape):
    """A class without annotations is a zero map (max over nothing -> clamp), in the tensor-core and the generic kernels;
    the other classes are unaffected."""
    from oracle import similarity as osim, synth
    from vittf_b200.similarity import similarity_maps
    feats, protos = synth.class_features(32, lr, 3, seed=2, dtype=torch.float16)
    g = torch.Generator().manual_seed(3)
    p = F.normalize(protos[torch.tensor([0, 0, 0, 2, 2])] + 0.05 * torch.randn(5, 32, generator=g), dim=-1)
    offs = [0, 0, 3, 3, 5, 5]                                           # classes 0, 2 and 4 are empty
    out = similarity_maps(feats.cuda(), p.cuda().contiguous(), torch.tensor(offs, dtype=torch.int32, device="cuda"),
                          out_shape, mode="ns", exponent=2.0).cpu()
    ref = osim.ns_composite(feats, p, [0, 3, 5], out_shape, exponent=2.0, slab=8)
    assert out.shape[0] == 5
    for c in (0, 2, 4):
        assert torch.count_nonzero(out[c]).item() == 0
    assert (out[1] - ref[0]).abs().max().item() < TOL and (out[3] - ref[1]).abs().max().item() < TOL


def test_legacy_similarity_matches_oracle():
    from oracle import similarity as osim, synth
    from vittf_b200 import infer
    from vittf_b200.similarity import class_offsets, rel_coords, similarity_maps
    vs, lr = (48, 40, 32), (12, 10, 8)
    feats, _ = synth.class_features(32, lr, 3, seed=5, dtype=torch.float32)
    ann = synth.annotations(vs, 3, 5, seed=5)
    ref = torch.stack(list(osim.legacy(feats, ann, vs).values()))
    fc = feats.cuda()
    rel = rel_coords(torch.cat(list(ann.values())), vs, fc.device)
    protos = F.normalize(infer.sample_features3d(fc, rel, mode="nearest")[0, 0], dim=-1).contiguous()
    out = similarity_maps(fc, protos, class_offsets(ann, fc.device), mode="legacy", exponent=2.0).cpu()
    assert (out - ref).abs().max().item() < TOL


def test_compose_labels_matches_oracle():
    from oracle import similarity as osim
    from vittf_b200.predict_ntf import compose_labels
    g = torch.Generator().manual_seed(4)
    sims = (torch.rand(5, 20, 18, 16, generator=g) * 255).to(torch.uint8)
    thr = [0.486, 0.264, 0.236, 0.68, 0.291]                   # predict_ntf.py:208
    assert torch.equal(compose_labels(sims, thr), osim.compose_labels(sims, thr))


@pytest.mark.parametrize("shape,z_range", [((24, 20, 16), None), ((21, 17, 13), None), ((32, 32, 32), (8, 20))])
def test_quantized_maps_match_the_reference_expression(shape, z_range):
    """predict_ntf.py:95-100 verbatim in torch -- quant = 0.99 * max, (255 / quant * sim).cpu().to(uint8) with its wrap above
    255, F.interpolate(nearest) to half the grid -- against vittf_quantize_maps_u8, whole volumes
    (even and odd extents) and a z-slab (the multi-GPU unit: global maxima, local planes).  Bit-exact."""
    from vittf_b200 import ops
    g = torch.Generator().manual_seed(6)
    C = 3
    sims = torch.rand((C,) + shape, generator=g) ** 2
    half = tuple(d // 2 for d in shape)
    ref = []
    for c in range(C):
        quant = 0.99 * sims[c].max()
        q = (255.0 / quant * sims[c]).to(torch.uint8)
        ref.append(F.interpolate(q[None, None], half, mode="nearest")[0, 0])
    ref = torch.stack(ref)
    assert (ref < 3).any() and (ref == 255).any()                                  # the wrap and the top bin are exercised
    sc = sims.cuda()
    cmax = ops.class_max(sc)
    if z_range is None:
        out, zo = ops.quantize_maps_u8(sc, cmax, half)
        assert zo == (0, half[2]) and torch.equal(out.cpu(), ref)
    else:
        z0, z1 = z_range
        out, (zo0, zo1) = ops.quantize_maps_u8(sc[..., z0:z1].contiguous(), cmax, half, depth=shape[2], z0=z0)
        assert (zo0, zo1) == (z0 // 2, z1 // 2) and torch.equal(out.cpu(), ref[..., zo0:zo1])


@pytest.mark.parametrize("switch", ["VITTF_SIM_GENERIC", "VITTF_SIM_UP_TC"])
def test_similarity_kernel_switches(switch):
    """The two A/B switches of the stage select kernels the default dispatch only uses for part of the shapes:
    VITTF_SIM_GENERIC = the generic kernels of both passes everywhere, VITTF_SIM_UP_TC = the tcgen05 up-sampling kernel for
    every prototype count (default: <= 8 prototypes or factor 2).  The NS parity sweep and the configs[0] / configs[1]
    full-size checks must hold under both (same oracle, same 2e-3)."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    env = dict(os.environ)
    env[switch] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", str(Path(__file__)), "-x", "-q", "-k",
                        "ns_similarity_matches_oracle or empty_classes or (full_size and (384-64-128 or 384-64-256))"],
                       env=env, capture_output=True, text=True, cwd=str(Path(__file__).resolve().parent.parent))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
