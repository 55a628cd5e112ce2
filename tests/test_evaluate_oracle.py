"""SURVEY.md 8f row 3 on the CPU: the restated evaluation loop (oracle/evaluate.py) against the metrics.json the reference's
own evaluate_similarities.py wrote for the same seeded inputs (tests/golden/eval_metrics.json)."""
import json
from pathlib import Path

import pytest

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


def test_oracle_reproduces_the_reference_metrics(tmp_path):
    from oracle import evaluate as oev
    d, label_fn, names = oev.make_inputs(tmp_path, seed=0)
    _close(oev.evaluate(d, label_fn, names), json.loads(GOLDEN.read_text()))
