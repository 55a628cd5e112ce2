"""Stage 3 on the GPU: bilateral solver against the golden vectors of the reference's own
apply_bilateral_solver3d (scipy CSR + cg) and the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(g):
    from oracle import synth
    shape = tuple(int(v) for v in g["shape"])
    ss, sl, sc = [int(v) for v in g["sig"]]
    r8, lab = synth.ct_volume(shape, n_shells=4, seed=7)
    gen = torch.Generator().manual_seed(11)
    t = ((lab == 1).float() * 0.8 + 0.2 * torch.rand(shape, generator=gen)).clamp(0, 1)
    cexp = torch.rand((1,) + shape, generator=gen)
    return t, r8, cexp, dict(sigma_spatial=ss, sigma_luma=sl, sigma_chroma=sc)


@pytest.mark.parametrize("name", ["bls_s755", "bls_s333", "bls_default"])
def test_apply_bilateral_solver3d_matches_reference(golden, name):
    from vittf_b200.bilateral_solver3d import apply_bilateral_solver3d
    g = golden(name)
    t, r8, cexp, gp = _inputs(g)
    out = apply_bilateral_solver3d(t[None], r8.expand(3, -1, -1, -1), grid_params=gp)
    assert out.dtype == torch.float32 and not out.is_cuda and out.shape == t.shape
    assert (out - torch.from_numpy(g["out"])).abs().max().item() < 1e-4
    out_c = apply_bilateral_solver3d(t[None].cuda(), r8.expand(3, -1, -1, -1), c=cexp, grid_params=gp)
    assert (out_c - torch.from_numpy(g["out_c"])).abs().max().item() < 1e-4


def test_sobel_confidence_matches_oracle():
    from oracle import bls, synth
    from vittf_b200 import ops
    r8, _ = synth.ct_volume((30, 26, 22), n_shells=4, seed=7)
    ref = torch.from_numpy(bls.sobel_confidence(r8)).float()
    assert (ops.sobel_confidence(r8.cuda()).cpu() - ref).abs().max().item() < 1e-6


def test_multi_rhs_equals_single_and_counts_iterations():
    from oracle import bls, synth
    from vittf_b200.bilateral_solver3d import solve_many
    shape = (40, 36, 28)
    r8, lab = synth.ct_volume(shape, n_shells=4, seed=7)
    t = torch.stack([(lab == c).float() * 0.9 for c in range(3)]).cuda()
    gp = dict(sigma_spatial=7, sigma_luma=5, sigma_chroma=5)
    many, iters = solve_many(t, r8.cuda(), None, gp)
    for c in range(3):
        one, _ = solve_many(t[c:c + 1], r8.cuda(), None, gp)
        assert (one[0] - many[c]).abs().max().item() < 1e-6
        ref, info = bls.solve_dense(t[c:c + 1].cpu(), r8.expand(3, -1, -1, -1), grid_params=gp, return_info=True)
        assert (many[c].cpu() - ref).abs().max().item() < 1e-4
        assert int(iters[c]) == info["iters"]


def test_colour_reference_is_rejected():
    from vittf_b200.bilateral_solver3d import apply_bilateral_solver3d
    r = torch.randint(0, 255, (3, 8, 8, 8), dtype=torch.uint8)
    with pytest.raises(NotImplementedError):
        apply_bilateral_solver3d(torch.rand(1, 8, 8, 8), r)


def test_crop_helpers(golden):
    from vittf_b200.bilateral_solver3d import crop_pad, write_crop_into
    g = golden("crop")
    s = torch.from_numpy(g["s"])
    crops, (lo, hi) = crop_pad([s, s[0] * 2], thresh=0.1, pad=2)
    assert np.array_equal(lo.numpy(), g["mi"]) and np.array_equal(hi.numpy(), g["ma"])
    assert torch.equal(crops[0], torch.from_numpy(g["c0"]))
    full = write_crop_into(torch.zeros_like(s), crops[0], (lo, hi))
    assert torch.equal(full[..., lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]], crops[0])
