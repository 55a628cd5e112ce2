"""resample_topk / take_most_dissimilar (infer.py:75-126): CPU restatement vs golden outputs of the reference."""
from pathlib import Path

import numpy as np
import torch

GOLD = np.load(Path(__file__).parent / "golden" / "refine.npz")


def _feats():
    from oracle import synth
    return synth.class_features(32, (12, 10, 8), 3, seed=9, dtype=torch.float32)[0]


def test_oracle_resample_topk_matches_reference():
    from oracle import refine
    sims = torch.from_numpy(GOLD["sims_in"])
    for K in (3, 8):
        got = refine.resample_topk(_feats(), sims, K=K)
        assert np.allclose(got.numpy(), GOLD[f"topk{K}"], atol=1e-6)


def test_oracle_take_most_dissimilar_matches_reference():
    from oracle import refine
    f2 = torch.from_numpy(GOLD["dissim_in"])
    for measure in ("cosine", "euclidean"):
        got = refine.take_most_dissimilar(f2, 9, measure).numpy()
        want = GOLD[f"dissim_{measure}"]
        assert sorted(map(tuple, got.round(5))) == sorted(map(tuple, want.round(5)))   # sorted=False: order unspecified
    assert refine.take_most_dissimilar(f2[:5], 9) is not None and refine.take_most_dissimilar(f2[:5], 9).shape[0] == 5
