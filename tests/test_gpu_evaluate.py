"""SURVEY.md 8f row 3 on the GPU: the drop-in evaluate_similarities (native confusion-matrix pass) against the golden
metrics.json of the reference's own script, against the sklearn oracle on multi-class / degenerate inputs, and at the
BASELINE.json volume size (integer work: bit-exact counts)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden" / "eval_metrics.json"


def _close(a, b, tol=1e-12):
    if isinstance(a, dict):
        assert a.keys() == b.keys()
        for k in a:
            _close(a[k], b[k], tol)
    elif isinstance(a, list):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _close(x, y, tol)
    elif isinstance(a, float) or isinstance(b, float):
        assert a == pytest.approx(b, abs=tol)
    else:
        assert a == b


def test_evaluate_matches_the_reference_script(tmp_path):
    from oracle import evaluate as oev
    from vittf_b200 import evaluate_similarities as ev
    d, label_fn, names = oev.make_inputs(tmp_path, seed=0)
    res = ev.evaluate(d, label_fn, names)
    gold = json.loads(GOLDEN.read_text())
    _close(res, gold)
    _close(json.loads((d / "metrics.json").read_text()), gold)                  # the file the CLI leaves behind
    assert ev.main(["--data", str(d), "--label", str(label_fn), "--labels", *names]) == 0


@pytest.mark.parametrize("n,K,seed", [(1, 2, 0), (15, 2, 1), (4099, 6, 2), (100003, 16, 3), (65536, 3, 4)])
def test_label_metrics_match_sklearn(n, K, seed):
    from oracle import evaluate as oev
    from vittf_b200 import evaluate_similarities as ev, ops
    g = torch.Generator().manual_seed(seed)
    t = torch.randint(0, K, (n,), generator=g, dtype=torch.uint8)
    p = torch.where(torch.rand(n, generator=g) < 0.7, t, torch.randint(0, K, (n,), generator=g, dtype=torch.uint8))
    if seed == 4:
        p[p == 1] = 0                                                            # a class that is never predicted
    _close(ev.label_metrics(t, p), oev.metrics(t.numpy(), p.numpy()))
    # unaligned views take the scalar path
    cm = ops.confusion_matrix(t.cuda()[1:], p.cuda()[1:], K) if n > 1 else None
    if cm is not None:
        ref = torch.zeros(K, K, dtype=torch.int64)
        ref.index_put_((t[1:].long(), p[1:].long()), torch.ones(n - 1, dtype=torch.int64), accumulate=True)
        assert torch.equal(cm.cpu(), ref)
    with pytest.raises(ValueError):
        ops.confusion_matrix(torch.full((8,), K, dtype=torch.uint8, device="cuda"), torch.zeros(8, dtype=torch.uint8, device="cuda"), K)


def test_confusion_matrix_full_size():
    """512^3 label volumes (BASELINE.json configs[2] size): exact counts against a histogram of t * K + p."""
    from vittf_b200 import ops, synth
    lab = synth.shell_labels((512, 512, 512), 6).to(torch.uint8).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    pred = torch.where(torch.rand(lab.shape, device="cuda", generator=g) < 0.9, lab, torch.roll(lab, 7, 2))
    cm = ops.confusion_matrix(lab, pred, 6)
    ref = torch.bincount((lab.long() * 6 + pred.long()).flatten(), minlength=36).view(6, 6)
    assert torch.equal(cm, ref) and int(cm.sum()) == 512 ** 3
