"""Annotation samplers (compare_feat_sampling.py:13-33): the CPU restatement against golden index sets produced by
the reference functions themselves (oracle/make_golden.py --only sampling)."""
from pathlib import Path

import numpy as np
import torch

GOLD = np.load(Path(__file__).parent / "golden" / "sampling.npz")


def _mask(cls):
    from oracle import synth
    return (synth.shell_labels(tuple(int(v) for v in GOLD["shape"]), 3) == cls).numpy()


def test_oracle_samplers_reproduce_reference_draws():
    from oracle import sampling
    for cls in (0, 1, 2):
        m = _mask(cls)
        torch.manual_seed(100 + cls)
        assert np.array_equal(sampling.sample_uniform(m, 12).numpy(), GOLD[f"uniform_{cls}"])
        torch.manual_seed(200 + cls)
        assert np.array_equal(sampling.sample_surface(m, 10, 4).numpy(), GOLD[f"surface_{cls}"])
        torch.manual_seed(300 + cls)
        assert np.array_equal(sampling.sample_both(m, 16, 4).numpy(), GOLD[f"both_{cls}"])
        torch.manual_seed(400 + cls)
        assert np.array_equal(sampling.sample_surface(m, 10 ** 6, 2).numpy(), GOLD[f"surface_all_{cls}"])


def test_sampled_points_lie_inside_the_mask():
    for cls in (0, 1, 2):
        m = _mask(cls)
        for key in ("uniform", "surface", "both", "surface_all"):
            pts = GOLD[f"{key}_{cls}"]
            assert m[pts[:, 0], pts[:, 1], pts[:, 2]].all()
