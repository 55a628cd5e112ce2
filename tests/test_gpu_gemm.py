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


# ---------------------------------------------------------------------------------------------------------------------
# LayerNorm folded into the GEMMs around it (vittf_gemm_bf16_ln): the chain of one pre-LN transformer block
#   x -> [ln_prepare] -> qkv consumer / fc1 consumer (LN applied in the epilogue) -> residual-stream producer -> consumer ...
# against torch (fp32 LayerNorm + Linear on the same bf16-rounded operands).
# ---------------------------------------------------------------------------------------------------------------------
def _ln_case(M, D, seed, offset=0.3, spread=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(M, D, device="cuda", generator=g)
    if spread:                                       # per-row scales over two decades and a few large channels ("massive
        x = x * torch.logspace(-1, 1, M, device="cuda")[:, None]      # activations"), plus a non-zero row mean
        x[:, 7] *= 20.0
    x = x + offset
    gamma = 1.0 + 0.3 * torch.randn(D, device="cuda", generator=g)
    beta = 0.2 * torch.randn(D, device="cuda", generator=g)
    return x, gamma, beta


def _fold(w, b, gamma, beta):
    from vittf_b200.vit import fold_layernorm
    wf, bf, cs = fold_layernorm(w.cpu(), b.cpu(), gamma.cpu(), beta.cpu())
    return wf.cuda(), bf.cuda(), cs.cuda()


@pytest.mark.parametrize("M,D", [(300, 384), (2 * 4097, 768), (148 * 128 + 77, 384), (131, 1024)])
def test_ln_prepare_layout_and_stats(M, D):
    from vittf_b200 import _lib, ops
    x, _, _ = _ln_case(M, D, seed=20)
    xt, xb, stats = ops.ln_prepare(x)
    assert torch.equal(ops.xt_to_rows(xt, M, D), x)
    assert torch.equal(xb, x.bfloat16())
    assert stats.shape == (ops.m_pad_of(M), _lib.LN_SLOTS, 2)
    s = stats.sum(1)[:M].double()
    assert torch.allclose(s[:, 0], x.double().sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(s[:, 1], (x.double() ** 2).sum(1), rtol=1e-5, atol=1e-3)
    assert stats[:M, 1:].abs().max().item() == 0


@pytest.mark.parametrize("M,D,N", [(513, 384, 1536), (2 * 4097, 768, 3072), (148 * 128 + 77, 384, 1536)])
def test_gemm_ln_consumer_gelu(M, D, N):
    """fc1: GELU(LN2(x) W^T + b) with LN2 applied as two row scalars in the epilogue."""
    from vittf_b200 import _lib, ops
    x, gamma, beta = _ln_case(M, D, seed=21)
    g = torch.Generator(device="cuda").manual_seed(5)
    w = torch.randn(N, D, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    wf, bf, cs = _fold(w, b, gamma, beta)
    xt, xb, stats = ops.ln_prepare(x)
    out = ops.gemm_bf16_ln(xb, wf, bf, _lib.EPI_BIAS_GELU_BF16, colsum=cs, stats=stats)
    ln = torch.nn.functional.layer_norm(x, (D,), gamma, beta, eps=1e-6)
    ref = torch.nn.functional.gelu(ln @ w.t() + b)
    # what the unfused path computes: bf16(LN(x)) against bf16(W) -- the fold must be no further from fp32 than that
    unfused = torch.nn.functional.gelu(ln.bfloat16().float() @ w.bfloat16().float().t() + b).bfloat16().float()
    err, err_unfused = (out.float() - ref).abs().max().item(), (unfused - ref).abs().max().item()
    assert err < 3e-2 * max(1.0, ref.abs().max().item()), (err, err_unfused)
    assert err < 2.0 * err_unfused + 1e-2, (err, err_unfused)


@pytest.mark.parametrize("tokens,B,D", [(65, 3, 384), (4097, 2, 768)])
def test_gemm_ln_consumer_qkv_and_kfeat(tokens, B, D):
    from vittf_b200 import _lib, ops
    M, heads = B * tokens, D // 64
    x, gamma, beta = _ln_case(M, D, seed=22)
    g = torch.Generator(device="cuda").manual_seed(6)
    w = torch.randn(3 * D, D, device="cuda", generator=g) * 0.05
    b = torch.randn(3 * D, device="cuda", generator=g) * 0.1
    wf, bf, cs = _fold(w, b, gamma, beta)
    xt, xb, stats = ops.ln_prepare(x)
    tok_pad = ops.tok_pad_of(tokens)
    qk, vt = ops.gemm_bf16_ln(xb, wf, bf, _lib.EPI_QKV_SPLIT, colsum=cs, stats=stats, tokens=tokens, tok_pad=tok_pad)
    ref = torch.nn.functional.layer_norm(x, (D,), gamma, beta, eps=1e-6) @ w.t() + b
    tol = 3e-2 * max(1.0, ref.abs().max().item())
    assert (qk.float() - ref[:, :2 * D]).abs().max().item() < tol
    v_ref = ref[:, 2 * D:].view(B, tokens, heads * 64).permute(0, 2, 1)
    vt = vt.view(B, heads * 64, tok_pad)
    assert (vt[:, :, :tokens].float() - v_ref).abs().max().item() < tol
    assert vt[:, :, tokens:].abs().max().item() == 0
    # K features: rows [D, 2D) of the folded weight, CLS rows dropped, fp16
    kf = ops.gemm_bf16_ln(xb, wf[D:2 * D].contiguous(), bf[D:2 * D].contiguous(), _lib.EPI_KFEAT_F16, colsum=cs[D:2 * D].contiguous(),
                          stats=stats, tokens=tokens)
    k_ref = ref[:, D:2 * D].view(B, tokens, D)[:, 1:].reshape(-1, D)
    assert kf.dtype == torch.float16 and kf.shape == k_ref.shape
    assert (kf.float() - k_ref).abs().max().item() < tol


@pytest.mark.parametrize("M,D,K", [(777, 384, 1536), (2 * 4097, 768, 3072), (2 * 4097, 768, 768), (148 * 128 + 77, 384, 384), (131, 1024, 1024)])
def test_gemm_ln_producer(M, D, K):
    """proj / fc2: the residual stream (row-tiled fp32) += A W^T + b, its bf16 copy and the per-row partial sums."""
    from vittf_b200 import _lib, ops
    a, w, b = _mk(M, D, K, seed=23)
    x, _, _ = _ln_case(M, D, seed=24, spread=False)
    xt, _, _ = ops.ln_prepare(x)
    want = x + _ref(a, w, b)
    xb, stats = ops.gemm_bf16_ln(a, w, b, _lib.EPI_BIAS_RESID_LN, xt=xt)
    got = ops.xt_to_rows(xt, M, D)
    assert (got - want).abs().max().item() < 4e-3 * max(1.0, want.abs().max().item())
    assert torch.equal(xb, got.bfloat16())                        # the copy is the rounding of the stream it wrote
    slots = ops.ln_slots(D)
    assert 0 < slots <= _lib.LN_SLOTS and stats[:, slots:].abs().sum().item() == 0
    s = stats.sum(1)[:M].double()
    assert torch.allclose(s[:, 0], got.double().sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(s[:, 1], (got.double() ** 2).sum(1), rtol=1e-5, atol=1e-3)
    # the per-slot partials cover disjoint column slices
    width = D // slots
    assert torch.allclose(stats[:M, 1, 0].double(), got[:, width:2 * width].double().sum(1), rtol=1e-5, atol=1e-3)


def test_gemm_ln_block_chain():
    """One pre-LN block's MLP half through the folded chain: x2 = x + fc2(GELU(fc1(LN2(x)))), then the statistics the
    producer wrote feed the next consumer (LN1 of the following block) -- against fp32 torch."""
    from vittf_b200 import _lib, ops
    M, D = 1000, 384
    x, gamma, beta = _ln_case(M, D, seed=25, spread=False)
    g = torch.Generator(device="cuda").manual_seed(9)
    w1, b1 = torch.randn(4 * D, D, device="cuda", generator=g) * 0.05, torch.randn(4 * D, device="cuda", generator=g) * 0.1
    w2, b2 = torch.randn(D, 4 * D, device="cuda", generator=g) * 0.03, torch.randn(D, device="cuda", generator=g) * 0.1
    w3, b3 = torch.randn(3 * D, D, device="cuda", generator=g) * 0.05, torch.randn(3 * D, device="cuda", generator=g) * 0.1
    g3, be3 = 1.0 + 0.3 * torch.randn(D, device="cuda", generator=g), 0.2 * torch.randn(D, device="cuda", generator=g)
    w1f, b1f, cs1 = _fold(w1, b1, gamma, beta)
    w3f, b3f, cs3 = _fold(w3, b3, g3, be3)
    xt, xb, stats = ops.ln_prepare(x)
    hid = ops.gemm_bf16_ln(xb, w1f, b1f, _lib.EPI_BIAS_GELU_BF16, colsum=cs1, stats=stats)
    xb2, stats2 = ops.gemm_bf16_ln(hid, w2.bfloat16(), b2, _lib.EPI_BIAS_RESID_LN, xt=xt)
    out = ops.gemm_bf16_ln(xb2, w3f, b3f, _lib.EPI_BIAS_BF16, colsum=cs3, stats=stats2)
    F = torch.nn.functional
    h_ref = F.gelu(F.layer_norm(x, (D,), gamma, beta, eps=1e-6) @ w1.t() + b1)
    x2_ref = x + h_ref @ w2.t() + b2
    ref = F.layer_norm(x2_ref, (D,), g3, be3, eps=1e-6) @ w3.t() + b3
    assert (ops.xt_to_rows(xt, M, D) - x2_ref).abs().max().item() < 3e-2 * max(1.0, x2_ref.abs().max().item())
    assert (out.float() - ref).abs().max().item() < 4e-2 * max(1.0, ref.abs().max().item())
