"""Seeded synthetic inputs for tests and benchmarks.

Pure generators (phantom volumes, class-structured feature volumes, annotation points): nothing here
is an algorithm of the hot path.
"""
import torch
import torch.nn.functional as F


def torus_volume(size=128, noise=0.1, seed=0, filled=True):
    """The ``torus_filled`` phantom of /root/reference/create_synthetic_volumes.py:17-25,
    54-69 with the (there unseeded, :40) uniform noise drawn from a seeded generator.
    Returns (volume fp16 in [0,1], label uint8)."""
    g = torch.Generator().manual_seed(seed)
    ls = torch.linspace(-1, 1, size)
    pos = torch.stack(torch.meshgrid(ls, ls, ls, indexing="xy"), dim=-1)
    q = torch.norm(pos[..., :2], dim=-1) - 0.5
    sdf = torch.norm(torch.stack([q, pos[..., 2]], dim=-1), dim=-1) - 0.2
    body = (sdf <= 0).float() if filled else (sdf.abs() < 0.05).float()
    vol = torch.clamp(body + torch.rand(body.shape, generator=g) * noise, 0, 1)
    return vol.to(torch.float16), (body > 0.5).to(torch.uint8)


def shell_labels(size, n_classes):
    """Concentric shells: label = clamp(floor(r * C), C-1), r = distance from the
    centre normalised to the half-diagonal of the inscribed sphere."""
    if isinstance(size, int):
        size = (size, size, size)
    axes = [torch.linspace(-1, 1, s) for s in size]
    gx, gy, gz = torch.meshgrid(*axes, indexing="ij")
    r = torch.sqrt(gx * gx + gy * gy + gz * gz).clamp(max=0.9999)
    return torch.clamp((r * n_classes).floor().long(), max=n_classes - 1)


def ct_volume(size, n_shells=8, seed=0):
    """CT-shaped uint8 phantom: nested shells with intensities linspace(.1,.9) + noise."""
    g = torch.Generator().manual_seed(seed)
    lab = shell_labels(size, n_shells)
    inten = torch.linspace(0.1, 0.9, n_shells)[lab]
    vol = (inten + 0.05 * torch.rand(lab.shape, generator=g)).clamp(0, 1)
    return (vol * 255).to(torch.uint8), lab


def class_features(f_dim, lr_size, n_classes, seed=0, noise=0.03, dtype=torch.float16):
    """Class-structured low-res feature volume (iid noise would give empty maps and
    crash the reference's crop_pad).  Returns (feats (F,f,f,f), class protos (C,F))."""
    g = torch.Generator().manual_seed(seed)
    protos = F.normalize(torch.randn(n_classes, f_dim, generator=g), dim=-1)
    lab = shell_labels(lr_size, n_classes)
    feats = protos[lab].permute(3, 0, 1, 2) + noise * torch.randn((f_dim,) + tuple(lab.shape), generator=g)
    return F.normalize(feats, dim=0).to(dtype).contiguous(), protos


def annotations(vol_size, n_classes, per_class, seed=0):
    """{name: LongTensor (N,3)} voxel coordinates (volume index space, X,Y,Z order)
    drawn uniformly from each class's shell (>= 2 in total)."""
    g = torch.Generator().manual_seed(seed + 1)
    lab = shell_labels(vol_size, n_classes)
    out = {}
    for c in range(n_classes):
        idx = (lab == c).nonzero()
        sel = torch.randperm(idx.size(0), generator=g)[:per_class]
        out[f"ntf{c + 1}"] = idx[sel].long()
    return out
