"""tcgen05 GEMM epilogues (vittf_gemm_bf16) against torch on the same bf16-rounded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, N, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    return a, w, b


def _ref(a, w, b):
    return a.float() @ w.float().t() + b


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 384, 384), (4097 * 2, 1536, 384), (1000, 384, 1536), (131, 768, 768)])
def test_gemm_bias_bf16(M, N, K):
    from vittf_b200 import _lib, ops
    a, w, b = _mk(M, N, K)
    out = ops.gemm_bf16(a, w, b, _lib.EPI_BIAS_BF16)
    ref = _ref(a, w, b)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), err


def test_gemm_gelu():
    from vittf_b200 import _lib, ops
    a, w, b = _mk(513, 1536, 384, seed=1)
    out = ops.gemm_bf16(a, w, b, _lib.EPI_BIAS_GELU_BF16)
    ref = torch.nn.functional.gelu(_ref(a, w, b))
    assert (out.float() - ref).abs().max().item() < 2e-2


def test_gemm_residual_fp32():
    from vittf_b200 import _lib, ops
    a, w, b = _mk(777, 384, 1536, seed=2)
    x = torch.randn(777, 384, device="cuda")
    ref = x + _ref(a, w, b)
    ops.gemm_bf16(a, w, b, _lib.EPI_BIAS_RESID_F32, out=x)
    assert (x - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("tokens,B", [(65, 3), (4097, 2)])
def test_gemm_qkv_split(tokens, B):
    from vittf_b200 import _lib, ops
    D, heads = 384, 6
    a, w, b = _mk(B * tokens, 3 * D, D, seed=3)
    tok_pad = ops.tok_pad_of(tokens)
    qk, vt = ops.gemm_bf16(a, w, b, _lib.EPI_QKV_SPLIT, tokens=tokens, tok_pad=tok_pad)
    ref = _ref(a, w, b)
    assert (qk.float() - ref[:, :2 * D]).abs().max().item() < 2e-2
    v_ref = ref[:, 2 * D:].view(B, tokens, heads * 64).permute(0, 2, 1)          # (B, heads*64, tokens)
    vt = vt.view(B, heads * 64, tok_pad)
    assert (vt[:, :, :tokens].float() - v_ref).abs().max().item() < 2e-2
    assert vt[:, :, tokens:].abs().max().item() == 0


def test_gemm_kfeat_drops_cls():
    from vittf_b200 import _lib, ops
    tokens, B, D = 65, 4, 384
    a, w, b = _mk(B * tokens, D, D, seed=4)
    out = ops.gemm_bf16(a, w, b, _lib.EPI_KFEAT_F16, tokens=tokens)
    ref = _ref(a, w, b).view(B, tokens, D)[:, 1:].reshape(-1, D)
    assert out.dtype == torch.float16 and out.shape == ref.shape
    assert (out.float() - ref).abs().max().item() < 5e-3 * max(1.0, ref.abs().max().item())


def test_gemm_vitb_shapes():
    """The GEMM shapes of the metric's backbone (ViT-B/8, two 4097-token images): fc2 (K = 3072), the QKV split at
    N = 2304 and fc1 + GELU at N = 3072."""
    from vittf_b200 import _lib, ops
    M, D, tokens = 2 * 4097, 768, 4097
    a, w, b = _mk(M, D, 4 * D, seed=11)                                            # fc2: (8194, 768, 3072)
    x = torch.randn(M, D, device="cuda")
    want = x + _ref(a, w, b)
    ops.gemm_bf16(a, w, b, _lib.EPI_BIAS_RESID_F32, out=x)
    assert (x - want).abs().max().item() < 4e-3 * max(1.0, want.abs().max().item())
    a, w, b = _mk(M, 3 * D, D, seed=12)                                            # qkv: (8194, 2304, 768)
    tok_pad = ops.tok_pad_of(tokens)
    qk, vt = ops.gemm_bf16(a, w, b, _lib.EPI_QKV_SPLIT, tokens=tokens, tok_pad=tok_pad)
    ref = _ref(a, w, b)
    assert (qk.float() - ref[:, :2 * D]).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    v_ref = ref[:, 2 * D:].view(2, tokens, D).permute(0, 2, 1)
    vt = vt.view(2, D, tok_pad)
    assert (vt[:, :, :tokens].float() - v_ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    assert vt[:, :, tokens:].abs().max().item() == 0
    a, w, b = _mk(M, 4 * D, D, seed=13)                                            # fc1 + GELU: (8194, 3072, 768)
    g = ops.gemm_bf16(a, w, b, _lib.EPI_BIAS_GELU_BF16)
    refg = torch.nn.functional.gelu(_ref(a, w, b))
    assert (g.float() - refg).abs().max().item() < 2e-2 * max(1.0, refg.abs().max().item())


@pytest.mark.parametrize("N", [384, 1152, 1536])
def test_gemm_many_row_blocks(N):
    """More row blocks than SMs (the persistent tile loop wraps) at K = 384, every epilogue."""
    from vittf_b200 import _lib, ops
    M, K = 148 * 128 + 77, 384
    a, w, b = _mk(M, N, K, seed=7)
    ref = _ref(a, w, b)
    out = ops.gemm_bf16(a, w, b, _lib.EPI_BIAS_BF16)
    assert (out.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    if N == 1536:
        g = ops.gemm_bf16(a, w, b, _lib.EPI_BIAS_GELU_BF16)
        assert (g.float() - torch.nn.functional.gelu(ref)).abs().max().item() < 2e-2
    if N == 384:
        x = torch.randn(M, N, device="cuda")
        want = x + ref
        ops.gemm_bf16(a, w, b, _lib.EPI_BIAS_RESID_F32, out=x)
        assert (x - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())
    if N == 1152:
        tokens = 37
        Mq = (M // tokens) * tokens
        qk, vt = ops.gemm_bf16(a[:Mq].contiguous(), w, b, _lib.EPI_QKV_SPLIT, tokens=tokens, tok_pad=128)
        assert (qk.float() - ref[:Mq, :768]).abs().max().item() < 2e-2
        v_ref = ref[:Mq, 768:].view(Mq // tokens, tokens, 384).permute(0, 2, 1)
        assert (vt.view(Mq // tokens, 384, 128)[:, :, :tokens].float() - v_ref).abs().max().item() < 2e-2


def test_gemm_rejects_bad_shapes():
    from vittf_b200 import _lib, ops
    a, w, b = _mk(64, 100, 64)
    with pytest.raises(_lib.VittfError):
        ops.gemm_bf16(a, w, b, _lib.EPI_BIAS_BF16)
