"""Host-side logic of the drop-in (no GPU): sizing rule, weight pre-processing, planning, argument handling."""
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def test_image_sizes_match_reference_rule():
    from oracle import feature_volume as ofv
    from vittf_b200 import infer
    for shape in [(128, 128, 128), (256, 256, 256), (40, 32, 24), (100, 120, 90), (512, 512, 300)]:
        for fos in (8, 64, 96):
            assert infer.image_sizes(shape, 8, fos) == ofv.image_sizes(shape, 8, fos)


def test_patch_embed_folding_identity():
    """conv on 3 identical normalised channels == folded 1-channel taps."""
    from oracle import dino_vit
    from vittf_b200.vit import IMAGENET_MEAN, IMAGENET_STD, fold_patch_embed
    model = dino_vit.build("vits8", depth=1)
    g = torch.rand(2, 1, 32, 40, generator=torch.Generator().manual_seed(0))
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    ref = model.patch_embed.proj((g.expand(-1, 3, -1, -1) - mean) / std)
    w, b = fold_patch_embed(model.patch_embed.proj.weight, model.patch_embed.proj.bias)   # (64, D), (D)
    out = F.conv2d(g, w.t().reshape(-1, 1, 8, 8), b, stride=8)
    assert (out - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("im", [(224, 224), (64, 64), (80, 48), (512, 512)])
def test_pos_embed_interpolation_matches_the_hub_rule(im):
    from oracle import dino_vit
    from vittf_b200.vit import interpolate_pos_embed
    model = dino_vit.build("vits8", depth=1)
    x = torch.zeros(1, 1 + (im[0] // 8) * (im[1] // 8), 384)
    ref = model.interpolate_pos_encoding(x, im[0], im[1])[0].clone()
    ref[0] += model.cls_token[0, 0]
    out = interpolate_pos_embed(model.pos_embed, model.cls_token, 8, im[0], im[1])
    assert torch.equal(out, ref)


def test_pool_target_recognises_the_reference_pool_functions():
    from vittf_b200 import infer
    f3 = (8, 6, 4)
    assert infer._pool_target(infer._noop, "z", 32, f3) == 32
    assert infer._pool_target(torch.nn.AdaptiveAvgPool3d(f3), "z", 32, f3) == 4
    assert infer._pool_target(torch.nn.AdaptiveAvgPool3d(f3), "x", 64, f3) == 8
    assert infer._pool_target(torch.nn.AdaptiveAvgPool3d((4, 6, 4)), "z", 32, f3) is None      # pools in-plane too
    assert infer._pool_target(lambda x: x, "z", 32, f3) is None


def test_compute_qkv_rejects_unsupported_keys_and_cpu():
    from oracle import dino_vit
    from vittf_b200 import infer
    model = dino_vit.build("vits8", depth=1)
    with pytest.raises(NotImplementedError):
        infer.compute_qkv(torch.zeros(8, 8, 8), model, 8, (16, 16, 16), return_keys=["q", "k"])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            infer.compute_qkv(torch.zeros(8, 8, 8), model, 8, (16, 16, 16), return_keys="k")


def test_cli_errors_like_the_reference(tmp_path, capsys):
    from vittf_b200 import infer
    with pytest.raises(SystemExit) as e:          # missing required flag -> argparse error
        infer.main([])
    assert e.value.code == 2
    if not torch.cuda.is_available():
        with pytest.raises(SystemExit) as e:      # no CUDA: explicit refusal instead of a CPU fallback
            infer.main(["--data-path", str(tmp_path / "v.npy")])
        assert e.value.code == 1
        assert "no CPU path" in capsys.readouterr().out


def test_load_data_formats(tmp_path):
    """infer.py:212-237: .pt tensor | {'vol':..}, .npy array | pickled object dict."""
    from vittf_b200 import infer
    v = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    torch.save(v, tmp_path / "a.pt")
    torch.save({"vol": v}, tmp_path / "b.pt")
    np.save(tmp_path / "c.npy", v.numpy())
    np.save(tmp_path / "d.npy", {"vol": v.numpy()}, allow_pickle=True)
    for n in ("a.pt", "b.pt", "c.npy", "d.npy"):
        assert torch.equal(infer.load_data(tmp_path / n).float(), v)
    with pytest.raises(SystemExit):
        infer.load_data(tmp_path / "missing.npy")


def test_dino_weight_container_loads_oracle_state_dict(tmp_path):
    from oracle import dino_vit
    from vittf_b200.dino import build_dino
    ref = dino_vit.build("vits8", seed=3)
    torch.save(ref.state_dict(), tmp_path / "w.pth")
    m = build_dino("vits8", weights=str(tmp_path / "w.pth"))
    assert sum(p.numel() for p in m.parameters()) == 21_670_272
    for (k1, a), (k2, b) in zip(sorted(ref.state_dict().items()), sorted(m.state_dict().items())):
        assert k1 == k2 and torch.equal(a, b)
    assert m.blocks[-1].attn.num_heads == 6 and m.blocks[-1].attn.qkv.in_features == 384


def test_luma_lut_matches_oracle():
    from oracle import bls
    from vittf_b200.bilateral_solver3d import luma_lut
    for s in (3, 4, 5, 7, 2.5):
        assert np.array_equal(luma_lut(s), bls.luma_lut(s))


def test_layernorm_fold_identity():
    """vit.fold_layernorm: LN(x) W^T + b == rstd * (x W'^T - mean * colsum(W')) + b' with W' = W * gamma (bf16), b' = b + W beta,
    for the statistics the GEMM epilogues form from partial (sum, sum of squares) -- evaluated in fp64 on the folded
    (rounded) weights, so that only the fold algebra is under test (gemm.cu header, /root/reference hub Block.forward)."""
    from vittf_b200.vit import fold_layernorm
    g = torch.Generator().manual_seed(3)
    D, N, M = 96, 40, 17
    w, b = torch.randn(N, D, generator=g) * 0.1, torch.randn(N, generator=g) * 0.1
    gamma, beta = 1.0 + 0.3 * torch.randn(D, generator=g), 0.2 * torch.randn(D, generator=g)
    x = torch.randn(M, D, generator=g) * torch.logspace(-1, 1, M)[:, None] + 0.4
    wf, bf, cs = fold_layernorm(w, b, gamma, beta)
    assert wf.dtype == torch.bfloat16 and bf.dtype == torch.float32 and cs.dtype == torch.float32
    assert torch.allclose(cs.double(), wf.double().sum(1), rtol=0, atol=1e-6)
    xd = x.double()
    # partial sums over 3 column slices, as the residual-stream epilogue writes them
    s = sum(xd[:, i:i + 32].sum(1) for i in range(0, D, 32))
    sq = sum((xd[:, i:i + 32] ** 2).sum(1) for i in range(0, D, 32))
    mean = s / D
    rstd = (sq / D - mean ** 2 + 1e-6).rsqrt()
    got = rstd[:, None] * (xd @ wf.double().t() - mean[:, None] * cs.double()[None, :]) + bf.double()
    # the same weights without the fold: gamma is inside wf, so un-fold it for the reference
    ln_unit = F.layer_norm(xd, (D,), None, None, eps=1e-6)                    # (x - mean) * rstd
    ref = ln_unit @ wf.double().t() + bf.double()
    assert torch.allclose(got, ref, rtol=1e-9, atol=1e-9)
    # and against the textbook form within the bf16 rounding of W * gamma
    full = F.layer_norm(xd, (D,), gamma.double(), beta.double(), eps=1e-6) @ w.double().t() + b.double()
    assert (got - full).abs().max().item() < 2e-2 * max(1.0, full.abs().max().item())
