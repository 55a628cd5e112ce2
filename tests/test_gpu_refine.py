"""resample_topk / take_most_dissimilar on the GPU vs the reference's golden outputs and the CPU oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "refine.npz")


def _feats():
    from oracle import synth
    return synth.class_features(32, (12, 10, 8), 3, seed=9, dtype=torch.float32)[0]


def test_topk_voxels_tie_rule():
    """thr = K-th largest with multiplicity; selection = first K indices with value >= thr (infer.py:91-92)."""
    from vittf_b200 import ops
    g = torch.Generator().manual_seed(3)
    maps = torch.rand(5, 4000, generator=g)
    maps[1] = (maps[1] * 8).round() / 8                    # many ties
    maps[2] = 0.5                                          # all equal
    for K in (1, 4, 8, 33):
        idx, thr = ops.topk_voxels(maps.cuda(), K)
        for m in range(5):
            t = torch.topk(maps[m], K, sorted=True).values[-1]
            want = (maps[m] >= t).nonzero()[:K, 0]
            assert thr[m].item() == t.item()
            assert torch.equal(idx[m].cpu(), want)


@pytest.mark.parametrize("K", [3, 8])
def test_resample_topk_matches_reference(K):
    from vittf_b200 import infer
    sims = torch.from_numpy(GOLD["sims_in"])
    got = infer.resample_topk(_feats(), sims.clone(), K=K, similarity_exponent=2.0)
    assert got.shape == (1, 3, 2, 12, 10, 8)
    assert np.abs(got[0].numpy() - GOLD[f"topk{K}"]).max() < 2e-5
    # fp16 feature volume on the device (the cache dtype): within fp16 resolution of the fp32 reference
    got16 = infer.resample_topk(_feats().half().cuda(), sims.clone().cuda(), K=K)
    assert got16.dtype == torch.float16 and got16.is_cuda
    assert np.abs(got16[0].float().cpu().numpy() - GOLD[f"topk{K}"]).max() < 5e-3


@pytest.mark.parametrize("measure", ["cosine", "euclidean"])
def test_take_most_dissimilar_matches_reference(measure):
    from oracle import refine
    from vittf_b200 import infer, ops
    f2 = torch.from_numpy(GOLD["dissim_in"])
    d = ops.mean_pairwise_distance(f2.cuda(), measure).cpu()
    ref = refine.mean_distance(f2, measure)                      # torch.cdist may take its matmul path (|x|^2 + |y|^2 - 2xy)
    assert (d - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    got = infer.take_most_dissimilar(f2, 9, measure).numpy()
    assert sorted(map(tuple, got.round(5))) == sorted(map(tuple, GOLD[f"dissim_{measure}"].round(5)))
    assert infer.take_most_dissimilar(f2[:5], 9) is f2[:5] or infer.take_most_dissimilar(f2[:5], 9).shape[0] == 5
    with pytest.raises(ValueError):
        infer.take_most_dissimilar(f2, 9, "manhattan")
