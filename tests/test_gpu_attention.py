"""tcgen05 flash attention (vittf_attention) against torch softmax(QK^T/8)V in fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,tokens,heads,scale", [(1, 17, 6, 1.0), (1, 65, 6, 1.0), (2, 128, 6, 1.0), (2, 129, 6, 1.0), (1, 144, 6, 2.0), (1, 256, 6, 1.0),
                                                   (1, 383, 6, 1.0), (2, 300, 6, 3.0), (1, 1025, 12, 1.0),
                                                   (2, 4097, 6, 1.0), (1, 4097, 6, 6.0)])
def test_attention_matches_torch(B, tokens, heads, scale):
    from vittf_b200 import ops
    D = heads * 64
    g = torch.Generator(device="cuda").manual_seed(tokens + heads)
    qkv = (torch.randn(B, tokens, 3, heads, 64, device="cuda", generator=g) * scale).bfloat16()
    qk = qkv[:, :, :2].reshape(B * tokens, 2 * D).contiguous()
    tok_pad = ops.tok_pad_of(tokens)
    vt = torch.zeros(B, D, tok_pad, dtype=torch.bfloat16, device="cuda")
    vt[:, :, :tokens] = qkv[:, :, 2].reshape(B, tokens, D).permute(0, 2, 1)
    out = ops.attention(qk, vt.view(B * D, tok_pad), B, tokens, heads, tok_pad)
    q, k, v = [qkv[:, :, i].permute(0, 2, 1, 3).float() for i in range(3)]
    att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v                  # (B, h, N, 64)
    ref = att.permute(0, 2, 1, 3).reshape(B * tokens, D)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err
    cos = torch.nn.functional.cosine_similarity(out.float(), ref, dim=-1).min().item()
    assert cos > 0.999, cos


@pytest.mark.parametrize("B,tokens,heads,scale,expect_flags", [(1, 17, 6, 1.0, False), (2, 129, 6, 1.0, False), (1, 383, 6, 2.0, False),
                                                              (2, 4097, 6, 1.0, False), (1, 1025, 12, 6.0, True),
                                                              (1, 4097, 6, 6.0, True)])
def test_attention_prescaled_two_pass(B, tokens, heads, scale, expect_flags):
    """The engine's entry point: q pre-scaled, max-free first pass, safe second pass over flagged CTAs.  Moderate scores
    never raise a flag; scores of several hundred (scale 6: |s| up to ~400 in log2 units) overflow 2^s, are flagged and
    must come out of the second pass exactly as from the safe kernel."""
    from vittf_b200 import ops
    D = heads * 64
    g = torch.Generator(device="cuda").manual_seed(tokens + heads)
    qkv = (torch.randn(B, tokens, 3, heads, 64, device="cuda", generator=g) * scale).bfloat16()
    qk = qkv[:, :, :2].clone()
    qk[:, :, 0] = (qk[:, :, 0].float() * (0.125 * 1.4426950408889634)).bfloat16()
    qk = qk.reshape(B * tokens, 2 * D).contiguous()
    tok_pad = ops.tok_pad_of(tokens)
    vt = torch.zeros(B, D, tok_pad, dtype=torch.bfloat16, device="cuda")
    vt[:, :, :tokens] = qkv[:, :, 2].reshape(B, tokens, D).permute(0, 2, 1)
    out, flags = ops.attention_prescaled(qk, vt.view(B * D, tok_pad), B, tokens, heads, tok_pad, return_flags=True)
    k, v = [qkv[:, :, i].permute(0, 2, 1, 3).float() for i in (1, 2)]
    q2 = qk.view(B, tokens, 2, heads, 64)[:, :, 0].permute(0, 2, 1, 3).float()        # the bf16 values the kernel sees
    att = torch.softmax(q2 @ k.transpose(-1, -2) * 0.6931471805599453, dim=-1) @ v     # 2^s = e^(s ln 2)
    ref = att.permute(0, 2, 1, 3).reshape(B * tokens, D)
    torch.cuda.synchronize()
    assert bool(flags.any().item()) == expect_flags
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err
    cos = torch.nn.functional.cosine_similarity(out.float(), ref, dim=-1).min().item()
    assert cos > 0.999, cos
