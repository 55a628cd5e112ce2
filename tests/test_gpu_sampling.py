"""Samplers + predict_ntf CLI on the GPU: native erosion vs scipy, seeded draws vs the
reference's golden index sets, and the CLI end to end on a synthetic data directory."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "sampling.npz")


def _mask(cls):
    from oracle import synth
    return (synth.shell_labels(tuple(int(v) for v in GOLD["shape"]), 3) == cls).numpy()


@pytest.mark.parametrize("conn", [1, 2, 3, 4])
def test_erosion_matches_scipy(conn):
    from oracle import sampling
    from vittf_b200 import ops
    g = torch.Generator().manual_seed(conn)
    m = (torch.rand(21, 17, 13, generator=g) > 0.15)
    m[:, :, 0] = True                                     # foreground touching the border: border_value = 0 matters
    want = sampling.erode(m.numpy(), conn)
    got = ops.binary_erosion(m.to("cuda", torch.uint8), conn).cpu().numpy().astype(bool)
    assert np.array_equal(got, want)


def test_dropin_samplers_reproduce_reference_draws():
    from vittf_b200 import compare_feat_sampling as cfs
    for cls in (0, 1, 2):
        m = _mask(cls)
        torch.manual_seed(100 + cls)
        assert np.array_equal(cfs.sample_uniform(m, 12).numpy(), GOLD[f"uniform_{cls}"])
        torch.manual_seed(200 + cls)
        assert np.array_equal(cfs.sample_surface(torch.as_tensor(m), 10, dist_from_surface=4).numpy(), GOLD[f"surface_{cls}"])
        torch.manual_seed(300 + cls)
        assert np.array_equal(cfs.sample_both(m, 16, dist_from_surface=4).numpy(), GOLD[f"both_{cls}"])
        torch.manual_seed(400 + cls)
        assert np.array_equal(cfs.sample_surface(m, 10 ** 6, dist_from_surface=2).numpy(), GOLD[f"surface_all_{cls}"])


def test_predict_ntf_cli_end_to_end():
    """volume.npy + labels.npy + *features*.npy -> ntf_pred64.0both.npy; rerun exits early (predict_ntf.py:123-125).
    The directory must not have 'pred' anywhere in its path: the reference's feature-file filter (:129) looks at the
    whole path string, and so does the drop-in."""
    import tempfile
    from oracle import similarity as osim, synth
    from vittf_b200 import predict_ntf
    tmp_path = Path(tempfile.mkdtemp(prefix="ntfcli_"))
    shape, lr = (32, 32, 32), (16, 16, 16)
    lab = synth.shell_labels(shape, 4)                               # classes 1..3 are annotated, 0 is background
    feats, _ = synth.class_features(24, lr, 4, seed=3, dtype=torch.float32)
    vol, _ = synth.ct_volume(shape, n_shells=4, seed=3)
    np.save(tmp_path / "volume.npy", vol.numpy())
    np.save(tmp_path / "labels.npy", lab.numpy().astype(np.uint8))
    np.save(tmp_path / "vol_features16.npy", {"k": feats.numpy()})
    torch.manual_seed(7)
    assert predict_ntf.main(["--data", str(tmp_path), "--num-samples", "64", "--sampling-mode", "both"]) == 0
    out = tmp_path / "ntf_pred64.0both.npy"
    pred = np.load(out)
    assert pred.shape == (16, 16, 16) and pred.dtype == np.uint8 and pred.max() <= 3
    # the same annotations through the CPU oracle of compute_similarities + the label rule
    from oracle import sampling
    labf = np.flip(lab.numpy(), axis=-3).copy()
    torch.manual_seed(7)
    ann = {f"ntf{i}": sampling.sample_both(labf == i, min(64, int((labf == i).sum()))) for i in range(1, 4)}
    volf = np.flip(vol.numpy().astype(np.float32), axis=-3).copy()
    ref = osim.ref_ntf(volf.shape, feats, ann)
    sims = torch.stack([ref[k] for k in ann])
    want = osim.compose_labels(sims, predict_ntf.CT_ORG_THRESHOLDS[:3]).numpy()
    agree = (pred == want).mean()
    assert agree >= 0.999, agree
    assert predict_ntf.main(["--data", str(tmp_path), "--num-samples", "64", "--sampling-mode", "both"]) == 0   # early exit
