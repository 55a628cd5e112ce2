"""The infer.py-compatible CLI end to end on the GPU: flags, cache naming, saved-file layout (infer.py:266-340)."""
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(args, cwd):
    return subprocess.run([sys.executable, "-m", "vittf_b200.infer"] + args, cwd=cwd, capture_output=True, text=True,
                          env={**__import__("os").environ, "PYTHONPATH": str(ROOT)})


def test_cli_writes_the_reference_cache_layout(tmp_path):
    from oracle import dino_vit, feature_volume as ofv
    from vittf_b200 import synth
    vol, _ = synth.ct_volume((32, 32, 32), n_shells=4, seed=3)
    torch.save(vol, tmp_path / "phantom.pt")
    ref_model = dino_vit.build("vits8", seed=5)
    torch.save(ref_model.state_dict(), tmp_path / "w.pth")
    r = _run(["--data-path", str(tmp_path / "phantom.pt"), "--feature-output-size", "8", "--batch-size", "4",
              "--weights", str(tmp_path / "w.pth")], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    cache = tmp_path / "phantom_vits8_all_features8.pt"          # infer.py:279 naming rule
    assert cache.exists()
    data = torch.load(cache, weights_only=False)
    assert isinstance(data, defaultdict) and list(data.keys()) == ["k"]
    k = data["k"]
    assert k.dtype == torch.float16 and tuple(k.shape) == (384, 8, 8, 8) and not k.is_cuda
    ref = ofv.feature_volume(vol, ref_model, patch=8, fos=8, batch_size=4)
    cos = torch.nn.functional.cosine_similarity(k.float().flatten(1).t(), ref.float().flatten(1).t(), dim=-1)
    assert cos.min().item() >= 0.995
    # second run without --overwrite refuses (infer.py:282-284)
    r2 = _run(["--data-path", str(tmp_path / "phantom.pt"), "--feature-output-size", "8"], tmp_path)
    assert r2.returncode == 1 and "already exists" in r2.stdout
    # .npy in -> .npy object-dict out, single axis keeps the un-pooled slice axis (infer.py:326,339-340)
    np.save(tmp_path / "v.npy", vol.numpy())
    r3 = _run(["--data-path", str(tmp_path / "v.npy"), "--feature-output-size", "8", "--slice-along", "y", "--batch-size", "8",
               "--weights", str(tmp_path / "w.pth")], tmp_path)
    assert r3.returncode == 0, r3.stdout + r3.stderr
    out = np.load(tmp_path / "v_vits8_y_features8.npy", allow_pickle=True)[()]
    assert out["k"].dtype == np.float16 and out["k"].shape == (384, 8, 32, 8)


def test_cli_refuses_cpu(tmp_path):
    torch.save(torch.zeros(8, 8, 8), tmp_path / "z.pt")
    r = _run(["--data-path", str(tmp_path / "z.pt"), "--cpu"], tmp_path)
    assert r.returncode == 1 and "no CPU path" in r.stdout
