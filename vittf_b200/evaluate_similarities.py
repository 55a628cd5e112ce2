"""Drop-in for /root/reference/evaluate_similarities.py: per-class accuracy / precision / recall / F1 /
IoU / confusion matrix of the predicted label volumes against a ground-truth label volume, written to metrics.json.

The only volume-sized work -- the table of (true, predicted) label pairs -- is one native pass over the two uint8 volumes
(vittf_confusion_matrix); the sklearn scores the reference calls (:66-71, average=None) are ratios of that table's entries.
Same CLI (`--data --label --labels`), same input files (predictions.npy, metadata.json, label volume .npy), same
metrics.json keys.  No CPU fallback: the volumes are moved to the CUDA device.
"""
import json
import sys
from argparse import ArgumentParser
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

label2idx = {'background': 0, 'liver': 1, 'bladder': 2, 'lung': 3, 'kidney': 4, 'bone': 5}     # evaluate_similarities.py:27-34
idx2label = ['liver', 'bladder', 'lung', 'kidney', 'bone']


def _safe_div(num, den):
    """sklearn's zero_division='warn' behaviour: 0 where the denominator is 0."""
    num, den = num.double(), den.double()
    return torch.where(den > 0, num / den.clamp_min(1), torch.zeros_like(num))


def label_metrics(labels, pred, dev=None):
    """The metrics block of evaluate_similarities.py:66-78 for two label tensors of equal size (any integer dtype with
    values in [0, 16)).  Classes are the sorted union of the values present, as sklearn's unique_labels."""
    dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else dev
    t = torch.as_tensor(labels).reshape(-1).to(dev).to(torch.uint8)
    p = torch.as_tensor(pred).reshape(-1).to(dev).to(torch.uint8)
    K = int(max(int(t.max().item()), int(p.max().item()))) + 1
    cm_full = ops.confusion_matrix(t, p, K).cpu()
    present = ((cm_full.sum(0) + cm_full.sum(1)) > 0).nonzero().flatten()
    cm = cm_full[present][:, present]
    tp = cm.diag()
    n_true, n_pred = cm.sum(1), cm.sum(0)
    fp, fn = n_pred - tp, n_true - tp
    return {
        'accuracy': float(tp.sum().double() / cm.sum().double()),
        'precision': _safe_div(tp, n_pred).tolist(),
        'recall': _safe_div(tp, n_true).tolist(),
        'f1': _safe_div(2 * tp, 2 * tp + fp + fn).tolist(),
        'iou': _safe_div(tp, tp + fp + fn).tolist(),
        'confusion_matrix': cm.tolist(),
    }


def evaluate(data_dir, label_fn, label_names, dev=None):
    """evaluate_similarities.py:45-83: returns the results dict and writes <data_dir>/metrics.json."""
    d = Path(data_dir)
    label_fn = Path(label_fn)
    assert (d / 'predictions.npy').exists()
    assert label_fn.exists()
    assert (d / 'metadata.json').exists()
    with (d / 'metadata.json').open('r', encoding='UTF-8') as f:
        metadata = json.load(f)
    dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else dev
    labels_orig = torch.as_tensor(np.load(label_fn, allow_pickle=True)[()]).to(dev)
    preds = {k: torch.as_tensor(v) for k, v in np.load(d / 'predictions.npy', allow_pickle=True)[()].items()}
    results = {}
    for ln, k in zip(label_names, sorted(preds.keys())):
        p = preds[k]
        meta = metadata[k]
        labels = F.interpolate((labels_orig == label2idx[ln]).to(torch.uint8)[None, None], p.shape[-3:], mode='nearest').reshape(-1)
        m = label_metrics(labels, p.reshape(-1), dev)
        m['annotation_time'] = meta['time']
        m['num_annotations'] = meta['num_annotations']
        results[ln] = m
    with open(d / 'metrics.json', 'w') as f:
        json.dump(results, f)
    return results


def main(argv=None):
    parser = ArgumentParser()
    parser.add_argument('--data', type=Path, help='Path to features, annotations, volume etc.')
    parser.add_argument('--label', type=Path, default='userstudy/labels-10.npy', help='Path to label volume')
    parser.add_argument('--labels', type=str, nargs='+', default=['lung', 'liver', 'kidney'], help='Labels found in predictions (in order)')
    args = parser.parse_args(argv)
    if not torch.cuda.is_available():
        print("vittf_b200 has no CPU fallback: a CUDA device is required", file=sys.stderr)
        return 1
    from pprint import pprint
    pprint(evaluate(args.data, args.label, args.labels))
    return 0


if __name__ == '__main__':
    sys.exit(main())
