"""Parameter container with the DINO ViT attribute layout.

The reference obtains its model from ``torch.hub.load('facebookresearch/dino:main', ...)``
(/root/reference/infer.py:42-43), which needs the network.  This module only HOLDS parameters under
the hub model's state-dict names (so real DINO checkpoints load with ``load_state_dict``) and exposes
the attributes the reference touches (``blocks[-1].attn.qkv``, ``.attn.num_heads``); the forward pass
is not implemented here -- it runs in the native engine (vittf_b200/vit.py).
"""
import torch
import torch.nn as nn

ARCHS = {"vits16": (384, 12, 6, 16), "vits8": (384, 12, 6, 8), "vitb16": (768, 12, 12, 16), "vitb8": (768, 12, 12, 8)}
# DINOv2 (infer.py:45-46,254-260; hub 'facebookresearch/dinov2' dinov2_vit{s,b,l}14): patch 14, 518^2 training size (37 x 37
# position grid), LayerScale after attention and MLP (blocks[i].ls1.gamma / ls2.gamma).  vitg14 (SwiGLU FFN, 1536-d) is not built.
ARCHS_V2 = {"vits14": (384, 12, 6, 14), "vitb14": (768, 12, 12, 14), "vitl14": (1024, 24, 16, 14)}


class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _LayerScale(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))


class _Block(nn.Module):
    def __init__(self, dim, heads, layer_scale=False):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, 4 * dim)
        if layer_scale:
            self.ls1 = _LayerScale(dim)
            self.ls2 = _LayerScale(dim)


class _PatchEmbed(nn.Module):
    def __init__(self, patch, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)


class DinoWeights(nn.Module):
    def __init__(self, name):
        super().__init__()
        v2 = name in ARCHS_V2
        dim, depth, heads, patch = (ARCHS_V2 if v2 else ARCHS)[name]
        self.embed_dim = dim
        self.patch_embed = _PatchEmbed(patch, dim)
        grid = (518 if v2 else 224) // patch
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, grid * grid + 1, dim))
        self.blocks = nn.ModuleList([_Block(dim, heads, layer_scale=v2) for _ in range(depth)])
        if v2:
            self.mask_token = nn.Parameter(torch.zeros(1, dim))
        self.norm = nn.LayerNorm(dim, eps=1e-6)

    def forward(self, *a, **k):
        raise RuntimeError("DinoWeights only holds parameters; the forward pass runs in the native engine "
                           "(vittf_b200.infer.compute_qkv)")


def unwrap_checkpoint(sd):
    """Accepts the backbone-only state dicts torch.hub serves as well as the official FULL DINO checkpoints
    ({'teacher': ..., 'student': ...} with 'module.' / 'backbone.' prefixes and projection-head entries)."""
    if isinstance(sd, dict):
        for key in ("teacher", "state_dict", "model"):
            if key in sd and isinstance(sd[key], dict):
                sd = sd[key]
                break
    out = {}
    for k, v in sd.items():
        for pre in ("module.", "backbone."):
            if k.startswith(pre):
                k = k[len(pre):]
        if not k.startswith("head."):
            out[k] = v
    return out


def build_dino(name, weights=None, seed=0):
    """DINO-style random init under `seed` (trunc-normal 0.02 for Linear / pos / cls, LayerNorm 1/0,
    Conv2d default), or the parameters of a DINO checkpoint."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = DinoWeights(name)
    nn.init.trunc_normal_(m.pos_embed, std=0.02)
    nn.init.trunc_normal_(m.cls_token, std=0.02)
    for mod in m.modules():
        if isinstance(mod, nn.Linear):
            nn.init.trunc_normal_(mod.weight, std=0.02)
            nn.init.zeros_(mod.bias)
        elif isinstance(mod, nn.LayerNorm):
            nn.init.ones_(mod.weight)
            nn.init.zeros_(mod.bias)
        elif isinstance(mod, _LayerScale):
            nn.init.uniform_(mod.gamma, 0.5, 1.5)            # (hub init is 1.0; random so that the folding is visible in tests)
    torch.random.set_rng_state(gen_state)
    if weights is not None:
        sd = torch.load(weights, map_location="cpu", weights_only=False)
        sd = unwrap_checkpoint(sd)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        if missing:
            raise RuntimeError(f"checkpoint {weights} lacks parameters: {missing[:5]}...")
    for p in m.parameters():
        p.requires_grad_(False)
    return m.eval()
