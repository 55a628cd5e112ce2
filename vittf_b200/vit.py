"""Host side of the ViT K-feature engine.

Reads the parameters of any module that has the DINO attribute layout the reference relies on
(``model.blocks[-1].attn.qkv`` / ``.attn.num_heads``, /root/reference/infer.py:135,180), pre-processes
them once (bf16 copies, patch-embed folding, pos-embed interpolation) and drives the native engine
(vittf_vit_k_features).  ``model.forward`` is never called.
"""
import ctypes as C
import math
import os
import weakref

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import BlockWeights, VitConfig, check, load, ptr, stream_ptr
from .ops import AXIS_INDEX

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # infer.py:39
IMAGENET_STD = (0.229, 0.224, 0.225)   # infer.py:40


def fold_patch_embed(weight, bias):
    """Conv2d(3, D, p, p) applied to three identical grey channels normalised with the ImageNet
    mean/std == one-channel conv with W' = sum_c W_c/std_c and b' = b - sum W_c mean_c/std_c
    (SURVEY.md App. D2).  Returns (tap-major (p*p, D) fp32, (D) fp32)."""
    w = weight.detach().double().cpu()
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float64).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float64).view(1, 3, 1, 1)
    w1 = (w / std).sum(dim=1)                                   # (D, p, p)
    b1 = bias.detach().double().cpu() - (w * mean / std).sum(dim=(1, 2, 3))
    d = w1.shape[0]
    return w1.reshape(d, -1).t().contiguous().float(), b1.float()


def fold_layernorm(weight, bias, ln_w, ln_b):
    """LayerNorm folded into the Linear that follows it (gemm.cu, vittf_gemm_bf16_ln):
        LN(x) W^T + b = rstd * (x W'^T - mean * colsum(W')) + b'     with W' = W * gamma, b' = b + W beta.
    Returns (W' as bf16, b' fp32, colsum fp32 = row sums of the ROUNDED W', so that the mean term cancels exactly what the
    tensor cores accumulate)."""
    w = weight.detach().double().cpu()
    g, be = ln_w.detach().double().cpu(), ln_b.detach().double().cpu()
    w_folded = (w * g[None, :]).float().to(torch.bfloat16)
    b_folded = (bias.detach().double().cpu() + w @ be).float()
    colsum = w_folded.double().sum(dim=1).float()
    return w_folded, b_folded, colsum


def interpolate_pos_embed(pos_embed, cls_token, patch, im0, im1):
    """hub `interpolate_pos_encoding` + the cls-token add of `prepare_tokens`, evaluated once per
    image size on the host.  Returns fp32 (1 + f0*f1, D), row 0 = cls_token + pos[0]."""
    pos = pos_embed.detach().float().cpu()
    n = pos.shape[1] - 1
    f0, f1 = im0 // patch, im1 // patch
    dim = pos.shape[-1]
    if f0 * f1 == n and im0 == im1:
        patch_pos = pos[0, 1:]
    else:
        g = int(math.sqrt(n))
        w0, h0 = f0 + 0.1, f1 + 0.1
        pp = F.interpolate(pos[:, 1:].reshape(1, g, g, dim).permute(0, 3, 1, 2), scale_factor=(w0 / g, h0 / g),
                           mode="bicubic")
        assert int(w0) == pp.shape[-2] and int(h0) == pp.shape[-1]
        patch_pos = pp.permute(0, 2, 3, 1).reshape(-1, dim)
    row0 = cls_token.detach().float().cpu().view(1, dim) + pos[0, :1]
    return torch.cat([row0, patch_pos], dim=0).contiguous()


class VitEngine:
    """Native engine bound to one module's parameters on one CUDA device."""

    def __init__(self, model, device, max_batch=8, max_tokens=4097 + 128):
        self.device = torch.device(device)
        blocks = model._modules["blocks"]
        attn = blocks[-1]._modules["attn"]
        self.num_heads = attn.num_heads
        self.embed_dim = attn._modules["qkv"].in_features
        proj = model.patch_embed.proj
        self.patch = proj.kernel_size[0]
        self.depth = len(blocks)
        self.mlp_hidden = blocks[0].mlp.fc1.out_features
        self.max_batch = max_batch
        self.max_tokens = max_tokens
        dev = self.device
        self._keep = []      # device tensors referenced by raw pointers inside the engine
        # norm1 / norm2 folded into qkv / fc1 (no LayerNorm pass; VITTF_NO_LNFOLD=1 keeps the separate LayerNorm kernel: A/B)
        # -- for embed dims the 256-wide CTA-pair tiles cover (ViT-B, ViT-L); at D = 384 the residual-stream epilogue would run on
        # 128-wide single-CTA tiles and loses what the fold saves (ViT-S/8 step: 457.8 ms folded, 456.2 ms with the kernel)
        self.ln_fold = (os.environ.get("VITTF_NO_LNFOLD") is None and self.embed_dim % 256 == 0
                        and self.mlp_hidden >= 2 * self.embed_dim
                        and 0 < load().vittf_gemm_ln_slots(self.embed_dim) <= _lib.LN_SLOTS)

        def f32(t):
            t = t.detach().to(dev, torch.float32).contiguous()
            self._keep.append(t)
            return t

        def bf16(t):
            t = t.detach().to(dev, torch.bfloat16).contiguous()
            self._keep.append(t)
            return t

        pw, pb = fold_patch_embed(proj.weight, proj.bias)
        self.patch_w, self.patch_b = f32(pw), f32(pb)
        arr = (BlockWeights * self.depth)()
        # the engine's attention (vittf_attention_prescaled) takes q pre-scaled by hd^-0.5 * log2(e), so that a score is
        # directly the base-2 exponent: fold the factor into the Q rows of the qkv projection once, in fp32, before the
        # bf16 copy (K and V rows -- and with them the hooked K features -- are untouched)
        head_dim = self.embed_dim // self.num_heads
        q_scale = torch.ones(3 * self.embed_dim, dtype=torch.float32)
        q_scale[:self.embed_dim] = head_dim ** -0.5 * math.log2(math.e)
        for i, blk in enumerate(blocks):
            qkv_w = blk.attn.qkv.weight.detach().float().cpu() * q_scale[:, None]
            qkv_b = blk.attn.qkv.bias.detach().float().cpu() * q_scale
            proj_w, proj_b = blk.attn.proj.weight.detach().float().cpu(), blk.attn.proj.bias.detach().float().cpu()
            fc2_w, fc2_b = blk.mlp.fc2.weight.detach().float().cpu(), blk.mlp.fc2.bias.detach().float().cpu()
            # DINOv2 LayerScale (x + gamma * f(x)): a per-output-channel factor of the projection that precedes it -- folded
            # into its weight rows and bias in fp32 before the bf16 copy
            mods = blk._modules
            if "ls1" in mods and hasattr(mods["ls1"], "gamma"):
                g1 = mods["ls1"].gamma.detach().float().cpu()
                proj_w, proj_b = proj_w * g1[:, None], proj_b * g1
            if "ls2" in mods and hasattr(mods["ls2"], "gamma"):
                g2 = mods["ls2"].gamma.detach().float().cpu()
                fc2_w, fc2_b = fc2_w * g2[:, None], fc2_b * g2
            fc1_w, fc1_b = blk.mlp.fc1.weight.detach().float().cpu(), blk.mlp.fc1.bias.detach().float().cpu()
            fields = dict(ln1_w=f32(blk.norm1.weight), ln1_b=f32(blk.norm1.bias),
                          proj_w=bf16(proj_w), proj_b=f32(proj_b),
                          ln2_w=f32(blk.norm2.weight), ln2_b=f32(blk.norm2.bias),
                          fc2_w=bf16(fc2_w), fc2_b=f32(fc2_b))
            if self.ln_fold:
                qw, qb, qs = fold_layernorm(qkv_w, qkv_b, blk.norm1.weight, blk.norm1.bias)
                fw, fb, fs = fold_layernorm(fc1_w, fc1_b, blk.norm2.weight, blk.norm2.bias)
                fields.update(qkv_w=bf16(qw), qkv_b=f32(qb), qkv_colsum=f32(qs), fc1_w=bf16(fw), fc1_b=f32(fb), fc1_colsum=f32(fs))
            else:
                fields.update(qkv_w=bf16(qkv_w), qkv_b=f32(qkv_b), fc1_w=bf16(fc1_w), fc1_b=f32(fc1_b))
            for k, v in fields.items():
                setattr(arr[i], k, v.data_ptr())
        self._blocks = arr
        self._pos_src = (model.pos_embed.detach().cpu().clone(), model.cls_token.detach().cpu().clone())
        self._pos_cache = {}
        cfg = VitConfig(self.embed_dim, self.depth, self.num_heads, self.patch, self.mlp_hidden)
        handle = C.c_void_p()
        with torch.cuda.device(dev):
            check(load().vittf_vit_create(C.byref(handle), C.byref(cfg), arr, ptr(self.patch_w), ptr(self.patch_b),
                                          max_batch, max_tokens), "vittf_vit_create")
        self._handle = handle
        self._ws = None
        self._finalizer = weakref.finalize(self, load().vittf_vit_destroy, handle)

    def timing(self, enable):
        check(load().vittf_vit_timing_enable(self._handle, int(enable)), "vittf_vit_timing_enable")

    def read_timing(self):
        """{'attention': (ms, launches), 'gemm': (ms, launches)} since the last read (synchronises)."""
        ms = (C.c_double * 2)()
        n = (C.c_int64 * 2)()
        check(load().vittf_vit_timing_read(self._handle, ms, n), "vittf_vit_timing_read")
        return {"attention": (ms[0], n[0]), "gemm": (ms[1], n[1])}

    def pos_for(self, im0, im1):
        key = (im0, im1)
        if key not in self._pos_cache:
            self._pos_cache[key] = interpolate_pos_embed(self._pos_src[0], self._pos_src[1], self.patch, im0, im1).to(self.device)
        return self._pos_cache[key]

    def _workspace(self, batch, tokens):
        need = load().vittf_vit_workspace_bytes(self._handle, batch, tokens)
        if need < 0:
            raise _lib.VittfError("vittf_vit_workspace_bytes: bad arguments")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def k_features(self, vol, axis, s0, s1, im0, im1, mm, out=None):
        """K features of the patch tokens of slices [s0, s1): fp16 (s1-s0, f0*f1, D)."""
        _lib.require_cuda(vol, mm, out)
        f0, f1 = im0 // self.patch, im1 // self.patch
        tokens = 1 + f0 * f1
        B = s1 - s0
        if out is None:
            out = torch.empty(B, f0 * f1, self.embed_dim, dtype=torch.float16, device=self.device)
        ws = self._workspace(B, tokens)
        X, Y, Z = vol.shape
        with torch.cuda.device(self.device):          # the engine's device need not be the current one
            check(load().vittf_vit_k_features(self._handle, ptr(vol), _lib.DTYPE_CODE[vol.dtype], X, Y, Z, AXIS_INDEX[axis],
                                              s0, s1, im0, im1, ptr(mm), ptr(self.pos_for(im0, im1)), ptr(out), ptr(ws),
                                              ws.numel(), stream_ptr(self.device)), "vittf_vit_k_features")
        return out


_ENGINES = weakref.WeakKeyDictionary()


def engine_for(model, device, max_batch=8, max_tokens=4097 + 128):
    """One engine per (module, device); parameters are read once (cache keyed on the module)."""
    per_model = _ENGINES.setdefault(model, {})
    key = (str(device), max_batch, max_tokens)
    if key not in per_model:
        per_model[key] = VitEngine(model, device, max_batch, max_tokens)
    return per_model[key]
