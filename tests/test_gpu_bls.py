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


def test_slab_sharded_solver_equals_single_pass():
    """two ranks' z-slabs splatted into the same grid vectors (what the all-reduce(sum) produces) and
    sliced per slab must reproduce the one-pass solve (fp64 atomics reorder the sums: 1e-6)."""
    import ctypes as C
    from vittf_b200 import _lib, ops, synth
    from vittf_b200.bilateral_solver3d import luma_lut, solve_many
    dev = torch.device("cuda")
    r8, lab = synth.ct_volume((40, 36, 44), n_shells=4, seed=5)
    t = torch.stack([(lab == c).float() * 0.8 + 0.1 + 0.05 * torch.rand(lab.shape) for c in range(3)]).to(dev)
    r8 = r8.to(dev)
    gp = dict(sigma_spatial=7, sigma_luma=5, sigma_chroma=5)
    for conf in (None, torch.rand(40, 36, 44, device=dev)):
        ref, it_ref = solve_many(t, r8, conf, gp)
        # emulate world_size = 2 on one device: the "all-reduce" adds the other rank's partial grid vectors
        lut = luma_lut(5)
        lut_dev = torch.from_numpy(lut).to(dev)
        slabs = [(0, 19), (19, 44)]
        lib = _lib.load()
        prm = _lib.BlsParams(40, 36, 44, 7.0, 256.0, 1e-5, 1e-5, 25, int(lut.max()) + 1)
        ncell = lib.vittf_bls_grid_cells(C.byref(prm))
        st = _lib.stream_ptr(dev)
        cmax = torch.zeros(1, device=dev)
        raws = []
        if conf is None:
            for z0, z1 in slabs:
                raw = torch.empty(40, 36, z1 - z0, device=dev)
                _lib.check(lib.vittf_bls_sobel_slab(_lib.ptr(r8), 40, 36, 44, z0, z1, _lib.ptr(raw), _lib.ptr(cmax), st), "sobel")
                raws.append(raw)
        acc = torch.zeros((2 + 3) * ncell, dtype=torch.float64, device=dev)
        for k, (z0, z1) in enumerate(slabs):
            ts = t[..., z0:z1].contiguous()
            cs = raws[k] if conf is None else conf[..., z0:z1].contiguous()
            _lib.check(lib.vittf_bls_splat_slab(C.byref(prm), _lib.ptr(ts), _lib.ptr(r8), _lib.ptr(cs),
                                                _lib.ptr(cmax) if conf is None else None, _lib.ptr(lut_dev), 3, z0, z1,
                                                _lib.ptr(acc), st), "splat")
        outs = []
        for z0, z1 in slabs:
            # every rank solves the (identical) reduced grid problem and slices its own slab
            o, it = ops.bls_solve_sharded(t[..., z0:z1].contiguous(), r8, None if conf is None else conf[..., z0:z1].contiguous(),
                                          lut_dev, 7, 256, 1e-5, 1e-5, 25, int(lut.max()) + 1, z0, z1,
                                          all_reduce_max=lambda x: x.copy_(cmax), all_reduce_sum=lambda x: x.copy_(acc))
            outs.append(o)
            assert torch.equal(it.cpu(), it_ref.cpu())
        got = torch.cat(outs, dim=-1)
        assert (got - ref).abs().max().item() < 1e-6


@pytest.mark.parametrize("n,n_rhs", [(256, 2), (512, 1)])
def test_bls_full_size(n, n_rhs):
    """BASELINE.json configs[4] sizes (256^3 and 512^3): the whole refined volume against the dense CPU oracle (1e-4, the
    solver tolerance of the golden-vector tests), plus the size-independent properties of the solve: a constant target
    is a fixed point, the solution is homogeneous of degree one in the target (PCG is; scale 0.5 is exact in binary)."""
    from oracle import bls, synth
    from vittf_b200.bilateral_solver3d import solve_many
    shape = (n, n, n)
    r8, lab = synth.ct_volume(shape, n_shells=8, seed=7)
    gp = dict(sigma_spatial=7, sigma_luma=5, sigma_chroma=5)
    gen = torch.Generator().manual_seed(3)
    t = torch.stack([((lab == c + 1).float() * 0.8 + 0.2 * torch.rand(shape, generator=gen)).clamp(0, 1) for c in range(n_rhs)])
    rc = r8.cuda()
    out, iters = solve_many(t.cuda(), rc, None, gp)
    for c in range(n_rhs):
        ref, info = bls.solve_dense(t[c:c + 1], r8.expand(3, -1, -1, -1), grid_params=gp, return_info=True)
        assert (out[c].cpu() - ref).abs().max().item() < 1e-4
        assert int(iters[c]) == info["iters"]
    half, _ = solve_many(t[:1].cuda() * 0.5, rc, None, gp)
    assert (half[0] - 0.5 * out[0]).abs().max().item() < 1e-6
    const, _ = solve_many(torch.full((1,) + shape, 0.625, device="cuda"), rc, None, gp)
    assert (const[0] - 0.625).abs().max().item() < 1e-5
