"""Multi-GPU partitioning logic on the CPU: planning functions and a world_size-2
gloo run that shards the slices of every axis, all-reduces the per-axis buffers and must reproduce the
un-sharded oracle bit-exactly."""
import os

import pytest
import torch
import torch.nn.functional as F


def test_slab_and_slice_planning_covers_everything_once():
    from vittf_b200 import dist
    for n_out, S in [(64, 256), (64, 512), (8, 32), (3, 7), (10, 40), (5, 5)]:
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                o0, o1 = dist.slab_range(n_out, world, r)
                a, b = dist.slices_for_slabs(S, n_out, o0, o1)
                seen += list(range(o0, o1))
                for o in range(o0, o1):                     # every AdaptiveAvgPool window lies inside [a, b)
                    w0, w1 = (o * S) // n_out, -((-(o + 1) * S) // n_out)
                    assert a <= w0 and w1 <= b
            assert seen == list(range(n_out))
    assert dist.slab_range(4, 8, 7) == (3, 4) and dist.slab_range(4, 8, 0) == (0, 0)
    zs = [dist.z_range(256, 8, r) for r in range(8)]
    assert zs[0][0] == 0 and zs[-1][1] == 256 and all(zs[i][1] == zs[i + 1][0] for i in range(7))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as tdist
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import dino_vit, feature_volume as ofv
    from vittf_b200 import dist, synth
    torch.set_num_threads(2)
    vol, _ = synth.ct_volume((40, 32, 24), n_shells=4, seed=3)
    model = dino_vit.build("vits8", seed=0, depth=2)
    im_sz, f_sz = ofv.image_sizes(tuple(vol.shape), 8, 8)
    acc = None
    for ax in ("z", "y", "x"):
        s_dim = ofv.AXIS_SLICE_DIM[ax]
        S, n_out = vol.shape[s_dim], f_sz[s_dim]
        o0, o1 = dist.slab_range(n_out, world, rank)
        a, b = dist.slices_for_slabs(S, n_out, o0, o1)
        full = ofv.k_features_axis(vol, model, 8, im_sz, ax, batch_size=4)      # un-pooled (D, ., ., .)
        # this rank's contribution: pool only its own slabs from its own slices, zeros elsewhere
        buf = torch.zeros((full.shape[0],) + tuple(f_sz), dtype=torch.float16)
        if o1 > o0:
            assert S % n_out == 0                                   # windows of different slabs are disjoint
            src = [slice(None)] * 4
            src[1 + s_dim] = slice(a, b)
            tgt = list(f_sz)
            tgt[s_dim] = o1 - o0
            dst = [slice(None)] * 4
            dst[1 + s_dim] = slice(o0, o1)
            buf[tuple(dst)] = F.adaptive_avg_pool3d(full[tuple(src)], tuple(tgt))
        dist.all_reduce_disjoint(buf)
        # the all-gather formulation of the same exchange: compact blocks, rank-major staging, un-permute
        assert dist.even_slabs(n_out, world)
        blk = [slice(None)] * 4
        blk[1 + s_dim] = slice(o0, o1)
        staging, _ = dist.gather_blocks(buf[tuple(blk)].contiguous())
        assert torch.equal(dist.unpermute_gathered(staging, ax), buf)
        acc = buf if acc is None else acc + buf
    if rank == 0:
        torch.save(acc, out)
    tdist.destroy_process_group()


def test_sharded_feature_volume_equals_oracle_gloo(tmp_path):
    import torch.multiprocessing as mp
    from oracle import dino_vit, feature_volume as ofv
    from vittf_b200 import synth
    out = tmp_path / "acc.pt"
    mp.spawn(_worker, args=(2, 29517, str(out)), nprocs=2, join=True)
    vol, _ = synth.ct_volume((40, 32, 24), n_shells=4, seed=3)
    ref = ofv.feature_volume(vol, dino_vit.build("vits8", seed=0, depth=2), patch=8, fos=8, batch_size=4)
    got = torch.load(out)
    assert got.dtype == torch.float16 and torch.equal(got, ref)


def test_x_slab_planning_covers_the_footprint():
    """dist.x_range / similarity.lowres_x_planes (a rank needs the low-res planes under its output slab +- 1):
    the slabs tile the output exactly once, and every low-res plane a slab's trilinear footprint reads (index rule of
    F.interpolate(align_corners=False), predict_ntf.py:87) lies inside the planes pass 1 evaluates for it."""
    from vittf_b200 import dist as vdist
    from vittf_b200.similarity import lowres_x_planes
    for lr_w, u in ((64, 2), (64, 4), (64, 8), (128, 4), (6, 2), (5, 8)):
        out_w = lr_w * u
        for world in (1, 2, 3, 8):
            covered = []
            for rank in range(world):
                x0, x1 = vdist.x_range(out_w, world, rank)
                covered += list(range(x0, x1))
                if x1 <= x0:
                    continue
                xa, xb = lowres_x_planes(lr_w, out_w, (x0, x1))
                assert 0 <= xa < xb <= lr_w
                for x in range(x0, x1):
                    src = max((x + 0.5) / u - 0.5, 0.0)
                    i0 = int(src)
                    i1 = min(i0 + 1, lr_w - 1)
                    assert xa <= i0 < xb, (lr_w, u, world, rank, x)
                    if src - i0 > 0.0:                       # (weight 0 at the clamped border: the plane is not read)
                        assert xa <= i1 < xb, (lr_w, u, world, rank, x)
            assert covered == list(range(out_w))
    assert lowres_x_planes(64, 200, (0, 10)) is None          # non-integer factor: every plane
