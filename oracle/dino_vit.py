"""Restatement of the DINO ViT the reference pulls from torch.hub.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference does ``torch.hub.load('facebookresearch/dino:main', 'dino_vits8')``
(/root/reference/infer.py:42-43); that repository is NOT vendored, the branch is
unpinned and there is no network, so the architecture is restated here from its
published definition.  The attribute names the reference
touches -- ``model._modules["blocks"][-1]._modules["attn"]._modules["qkv"]``
(infer.py:135) and ``.attn.num_heads`` (infer.py:180) -- are preserved, as are
the state-dict key names of the hub model, so that real DINO checkpoints load.

Parity status: UNPINNED (no reference test or golden vector exists for the ViT;
our pins are the parameter counts 21 670 272 (S/8) / 85 807 872 (B/8) and the
dead-code identity checked in tests/test_oracle.py).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

ARCHS = {
    # name: (embed_dim, depth, heads, patch)
    "vits16": (384, 12, 6, 16),
    "vits8": (384, 12, 6, 8),
    "vitb16": (768, 12, 12, 16),
    "vitb8": (768, 12, 12, 8),
}


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, c // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(b, n, c)
        return self.proj(x), attn


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        y, _ = self.attn(self.norm1(x))
        x = x + y
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, embed_dim):
        super().__init__()
        self.num_patches = (img_size // patch_size) ** 2
        self.patch_size = patch_size
        self.proj = nn.Conv2d(3, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    """DINO ViT, num_classes=0 (head is Identity), img_size 224."""

    def __init__(self, patch_size=8, embed_dim=384, depth=12, num_heads=6, img_size=224):
        super().__init__()
        self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, embed_dim)
        g2 = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, g2 + 1, embed_dim))
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def interpolate_pos_encoding(self, x, w, h):
        npatch = x.shape[1] - 1
        n = self.pos_embed.shape[1] - 1
        if npatch == n and w == h:
            return self.pos_embed
        cls_pos = self.pos_embed[:, 0]
        patch_pos = self.pos_embed[:, 1:]
        dim = x.shape[-1]
        p = self.patch_embed.patch_size
        w0, h0 = w // p + 0.1, h // p + 0.1
        g = int(math.sqrt(n))
        patch_pos = F.interpolate(
            patch_pos.reshape(1, g, g, dim).permute(0, 3, 1, 2),
            scale_factor=(w0 / g, h0 / g),
            mode="bicubic",
        )
        assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
        patch_pos = patch_pos.permute(0, 2, 3, 1).view(1, -1, dim)
        return torch.cat((cls_pos.unsqueeze(0), patch_pos), dim=1)

    def prepare_tokens(self, x):
        b, _, w, h = x.shape
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(b, -1, -1), x), dim=1)
        return x + self.interpolate_pos_encoding(x, w, h)

    def forward(self, x):
        x = self.prepare_tokens(x)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)[:, 0]


# ---------------------------------------------------------------------------------------------
# DINOv2 (/root/reference/infer.py:45-46: torch.hub.load('facebookresearch/dinov2', 'dinov2_vits14'), patch 14,
# infer.py:254-260 -- the branch carries a typo (`dinoo_model`, :258) and never ran; this restates the hub model it meant
# to load from the published dinov2/models/vision_transformer.py: img_size 518 (37 x 37 position grid), LayerScale after
# the attention and the MLP (init_values 1.0), interpolate_offset 0.1 (the same bicubic rule as DINO v1), no register
# tokens, mask_token unused at inference.  vitg14 (SwiGLU FFN, 1536-d) is not restated.  Parity: UNPINNED like v1; pin =
# the parameter count of dinov2_vits14, 22 056 576.
# ---------------------------------------------------------------------------------------------
ARCHS_V2 = {
    "vits14": (384, 12, 6, 14),
    "vitb14": (768, 12, 12, 14),
    "vitl14": (1024, 24, 16, 14),
}


class LayerScale(nn.Module):
    def __init__(self, dim, init_values=1.0):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class BlockV2(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = LayerScale(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = LayerScale(dim)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x))[0])
        return x + self.ls2(self.mlp(self.norm2(x)))


class DinoV2VisionTransformer(VisionTransformer):
    def __init__(self, patch_size=14, embed_dim=384, depth=12, num_heads=6, img_size=518):
        super().__init__(patch_size=patch_size, embed_dim=embed_dim, depth=0, num_heads=num_heads, img_size=img_size)
        self.blocks = nn.ModuleList([BlockV2(embed_dim, num_heads) for _ in range(depth)])
        self.mask_token = nn.Parameter(torch.zeros(1, embed_dim))
        self.blocks.apply(self._init_weights)
        # LayerScale is 1.0 in the hub constructor and learned afterwards: random values so that a test sees its effect
        for blk in self.blocks:
            nn.init.uniform_(blk.ls1.gamma, 0.5, 1.5)
            nn.init.uniform_(blk.ls2.gamma, 0.5, 1.5)

    def prepare_tokens_with_masks(self, x, masks=None):
        return self.prepare_tokens(x)


def build(name="vits8", seed=0, depth=None):
    """Random-init model under a fixed seed (north_star: random-init weights)."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    if name in ARCHS_V2:
        d, l, h, p = ARCHS_V2[name]
        model = DinoV2VisionTransformer(patch_size=p, embed_dim=d, depth=depth or l, num_heads=h).eval()
        torch.random.set_rng_state(gen_state)
        for prm in model.parameters():
            prm.requires_grad_(False)
        return model
    d, l, h, p = ARCHS[name]
    model = VisionTransformer(patch_size=p, embed_dim=d, depth=depth or l, num_heads=h).eval()
    torch.random.set_rng_state(gen_state)
    for prm in model.parameters():
        prm.requires_grad_(False)
    return model
