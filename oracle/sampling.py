"""TEST INFRASTRUCTURE ONLY -- CPU restatement of /root/reference/compare_feat_sampling.py:13-33 (annotation
samplers) with the reference's own third-party calls (scipy.ndimage.binary_erosion / generate_binary_structure,
torch.multinomial).  Pinned by tests/golden/sampling.npz = outputs of the reference functions themselves under a
fixed seed (oracle/make_golden.py)."""
import numpy as np
import torch
from scipy.ndimage import binary_erosion, generate_binary_structure

ONE = torch.ones(1)


def sample_uniform(vol, n_samples, thin_to_reasonable=False):      # :13-17
    idxs = torch.as_tensor(vol).nonzero()
    while thin_to_reasonable and idxs.size(0) > int(2 ** 24):
        idxs = idxs[::2]
    return idxs[torch.multinomial(ONE.expand(idxs.size(0)), n_samples)]


def erode(vol, connectivity):                                      # :20-23
    return binary_erosion(np.asarray(vol), generate_binary_structure(rank=3, connectivity=connectivity))


def surface_voxels(vol, dist_from_surface=4):                      # :20-25
    outer = erode(vol, dist_from_surface)
    inner = erode(outer, 1)
    return torch.as_tensor(np.logical_xor(inner, outer)).nonzero()


def sample_surface(vol, n_samples, dist_from_surface=4):           # :19-31
    s = surface_voxels(vol, dist_from_surface)
    if s.size(0) > n_samples:
        return s[torch.multinomial(ONE.expand(s.size(0)), n_samples)]
    return s


def sample_both(vol, n_samples, dist_from_surface=4, thin_to_reasonable=False):   # :33-34
    return torch.cat([sample_uniform(vol, n_samples // 2, thin_to_reasonable), sample_surface(vol, n_samples // 2, dist_from_surface)])
