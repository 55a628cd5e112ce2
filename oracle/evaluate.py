"""CPU restatement of /root/reference/evaluate_similarities.py:58-83.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference file is a script (no importable function), so its
loop body is restated with the same sklearn calls; pinned by tests/golden/eval_metrics.json = the metrics.json the
reference script itself wrote for the seeded inputs of `make_inputs` (oracle/make_golden.py::golden_eval).
"""
import json
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from sklearn.metrics import accuracy_score, confusion_matrix, jaccard_score, precision_recall_fscore_support

label2idx = {'background': 0, 'liver': 1, 'bladder': 2, 'lung': 3, 'kidney': 4, 'bone': 5}      # :27-34


def metrics(labels, pred):
    """:66-78 for flat integer arrays."""
    labels, pred = np.asarray(labels).reshape(-1), np.asarray(pred).reshape(-1)
    prec, rec, f1, _ = precision_recall_fscore_support(labels, pred, average=None, zero_division=0)
    return {
        'accuracy': float(accuracy_score(labels, pred)),
        'precision': prec.tolist(), 'recall': rec.tolist(), 'f1': f1.tolist(),
        'iou': jaccard_score(labels, pred, average=None, zero_division=0).tolist(),
        'confusion_matrix': confusion_matrix(labels, pred).tolist(),
    }


def evaluate(data_dir, label_fn, label_names):
    """:45-83 without the file write."""
    d = Path(data_dir)
    metadata = json.loads((d / 'metadata.json').read_text())
    labels_orig = torch.as_tensor(np.load(label_fn, allow_pickle=True)[()])
    preds = {k: torch.as_tensor(v) for k, v in np.load(d / 'predictions.npy', allow_pickle=True)[()].items()}
    results = {}
    for ln, k in zip(label_names, sorted(preds.keys())):
        p = preds[k]
        labels = F.interpolate((labels_orig == label2idx[ln]).to(torch.uint8)[None, None], p.shape[-3:], mode='nearest').reshape(-1)
        m = metrics(labels.numpy(), p.reshape(-1).numpy())
        m['annotation_time'] = metadata[k]['time']
        m['num_annotations'] = metadata[k]['num_annotations']
        results[ln] = m
    return results


def make_inputs(directory, seed=0, label_shape=(40, 36, 32), pred_shape=(20, 18, 16)):
    """Seeded user-study style inputs: a 6-class label volume, three half-resolution binary predictions (noisy copies of
    the lung / liver / kidney masks), metadata.json.  Returns (data_dir, label_path, label_names)."""
    from . import synth
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    g = torch.Generator().manual_seed(seed)
    labels = synth.shell_labels(label_shape, 6).to(torch.uint8)
    names = ['lung', 'liver', 'kidney']
    preds, meta = {}, {}
    for i, ln in enumerate(names):
        m = F.interpolate((labels == label2idx[ln]).to(torch.uint8)[None, None], pred_shape, mode='nearest')[0, 0]
        flip = torch.rand(pred_shape, generator=g) < 0.07 * (i + 1)
        preds[f'ntf{i + 1}'] = (m ^ flip.to(torch.uint8)).numpy()
        meta[f'ntf{i + 1}'] = {'time': 10.5 * (i + 1), 'num_annotations': 3 + i}
    np.save(d / 'predictions.npy', preds, allow_pickle=True)
    np.save(d / 'labels.npy', labels.numpy())
    (d / 'metadata.json').write_text(json.dumps(meta))
    return d, d / 'labels.npy', names
