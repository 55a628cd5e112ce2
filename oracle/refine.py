"""TEST INFRASTRUCTURE ONLY -- CPU restatement of /root/reference/infer.py:75-126 (`resample_topk`,
`take_most_dissimilar`), the legacy refinement step between similarity and solver (old/cluster_dino.py:319-320).
Pinned by tests/golden/refine.npz = outputs of the reference functions (oracle/make_golden.py --only refine)."""
import torch
import torch.nn.functional as F

from .similarity import sample_prototypes


@torch.no_grad()
def resample_topk(feat_vol, sims, K=8, similarity_exponent=2.0, mode="nearest"):
    """feat_vol (F,W,H,D) fp32, sims (C,A,W,H,D) -> (C,A,W,H,D)   [infer.py:88-106 with M = 1]"""
    C, A = sims.shape[:2]
    dims = sims.shape[-3:]
    tops = []
    for s in sims.reshape(-1, *dims):
        thr = torch.topk(s.flatten(), K, largest=True, sorted=True).values[-1]      # :91
        tops.append((s >= thr).nonzero()[:K])                                       # :92 (index order, NOT value order)
    top = torch.stack(tops).float()                                                 # (C*A, K, 3)
    rel = (top + 0.5) / torch.tensor([[list(dims)]], dtype=torch.float32) * 2.0 - 1.0
    qf2 = sample_prototypes(feat_vol, rel.reshape(-1, 3), mode).reshape(C, A, K, -1)
    out = torch.einsum("fwhd,cakf->cakwhd", feat_vol, qf2).clamp(0, 1) ** similarity_exponent
    return out.mean(dim=2)


@torch.no_grad()
def mean_distance(features, measure="cosine"):                                      # :118-121
    if measure == "cosine":
        return 1 - F.cosine_similarity(features.unsqueeze(0), features.unsqueeze(1), dim=-1).squeeze(0).mean(0)
    return torch.cdist(features.unsqueeze(0), features.unsqueeze(0)).squeeze(0).mean(0)


@torch.no_grad()
def take_most_dissimilar(features, num_prototypes=35, measure="cosine"):
    if features.size(0) <= num_prototypes:
        return features
    _, sel = torch.topk(mean_distance(features, measure), num_prototypes, largest=True, sorted=False)
    return features[sel]
